"""Per-source-line stall samples of one profiled kernel, on the build box (no GPU): joins the SASS page of an ncu report
(`ncu -i rep --page source --csv`, instruction order) with the line table nvdisasm prints for the same function in the
in-tree libb200rag.so (compiled with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <mangled-name-substring> [--top 25]
"""
import argparse
import csv
import glob
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("func", help="substring of the mangled kernel name, e.g. dense_finish_kernelILi1")
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--kernel", default=None, help="regex selecting one kernel of a multi-kernel report (ncu -k regex:...)")
ap.add_argument("--lib", default=os.path.join(ROOT, "advanced-rag-milvus_b200", "b200rag", "libb200rag.so"))
args = ap.parse_args()

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", args.lib], cwd=tmp, capture_output=True)
lines = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    m = re.search(r"^\.text\.(\S*%s\S*):$" % re.escape(args.func), dis, re.M)
    if not m:
        continue
    body = dis[m.end():]
    end = re.search(r"^//-{10,} \.", body, re.M)
    body = body[: end.start()] if end else body
    lines, cur = [], ("?", 0)
    for ln in body.splitlines():
        mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    break
if lines is None:
    raise SystemExit(f"function matching {args.func!r} not found in {args.lib}")

cmd = ["ncu", "-i", args.report, "--page", "source", "--csv"] + (["-k", "regex:" + args.kernel] if args.kernel else [])
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
if len(data) != len(lines):
    print(f"warning: {len(data)} profiled instructions vs {len(lines)} disassembled (library rebuilt since the capture?)")
agg = {}
tot_s = tot_i = 0
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else ("?", 0)
    s, n = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
    a = agg.setdefault(key, [0, 0])
    a[0] += s
    a[1] += n
    tot_s += s
    tot_i += n
print(f"{tot_s} samples, {tot_i} warp instructions")
src_cache = {}
for (f, l), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[: args.top]:
    path = os.path.join(ROOT, "advanced-rag-milvus_b200", "csrc", f)
    if path not in src_cache:
        src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
    text = src_cache[path][l - 1].strip()[:95] if 0 < l <= len(src_cache[path]) else ""
    print(f"{100 * s / max(tot_s, 1):5.1f}% samples {100 * n / max(tot_i, 1):5.1f}% instr  {f}:{l:<4d} {text}")
