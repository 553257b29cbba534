set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
timeout 300 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_scan3 --launch-skip 9 --launch-count 1 -o gpurun_out/r2_scan3 $B > gpurun_out/r2_scan3.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_finish2 --launch-skip 4 --launch-count 1 -o gpurun_out/r2_finish2 $B > gpurun_out/r2_finish2.out 2>&1
timeout 300 python tools/mmr_one.py 3 256
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmr_select_inv --launch-skip 3 --launch-count 1 -o gpurun_out/r2_mmr python tools/mmr_one.py 3 256 > gpurun_out/r2_mmr.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_query --launch-skip 3 --launch-count 1 -o gpurun_out/r2_sparse python tools/sparse_ab.py > gpurun_out/r2_sparse.out 2>&1
ls -la gpurun_out/*.ncu-rep
