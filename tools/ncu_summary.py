"""Condense an ncu report (.ncu-rep, read with `ncu -i ... --page raw --csv`) into the few numbers the roofline
discussion needs, as a markdown table.  Runs on the build box (no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--flops F] [--bytes B] > profiles/<name>.md
"""
import argparse
import csv
import io
import subprocess

PICK = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock during the capture"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_size", "cluster size"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (of elapsed)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "DRAM read % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "  of which TMA loads"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2 -> SM rate"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--flops", type=float, default=0.0, help="algorithmic FLOPs of one launch")
    ap.add_argument("--bytes", type=float, default=0.0, help="algorithmic HBM bytes of one launch")
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"### `{r[col['Kernel Name']].split('(')[0]}`  (launch id {r[col['ID']]})\n")
        print("| quantity | value |")
        print("|---|---|")
        vals = {}
        for key, label in PICK:
            if key in col and r[col[key]] != "":
                vals[key] = (r[col[key]], units[col[key]])
                print(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
        dur = vals.get("gpu__time_duration.sum")
        if dur:
            scale = {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}[dur[1]]
            t = float(dur[0]) * scale
            if args.flops:
                print(f"| algorithmic FLOPs / duration | {args.flops / t / 1e12:.0f} TFLOP/s (cold cache, profiler-serialised) |")
            if args.bytes:
                print(f"| algorithmic bytes / duration | {args.bytes / t / 1e9:.0f} GB/s |")
        dr = vals.get("dram__bytes_read.sum")
        dw = vals.get("dram__bytes_write.sum")
        if dr and dw:
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            traffic = float(dr[0]) * mult[dr[1]] + float(dw[0]) * mult[dw[1]]
            print(f"| **traffic** = DRAM read + write | {traffic:.4g} bytes"
                  + (f" = {traffic / args.bytes:.2f} x algorithmic |" if args.bytes else " |"))
        print()


if __name__ == "__main__":
    main()
