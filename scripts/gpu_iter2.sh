#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dwconv" -p no:cacheprovider 2>&1 | tail -4
timeout -s KILL 600 python scripts/bench_ops.py dwconv 2>&1 | tee gpurun_out/bench_ops_dwconv.txt
GA_DW_TMA=0 timeout -s KILL 600 python scripts/bench_ops.py dwconv 2>&1 | tee gpurun_out/bench_ops_dwconv_old.txt
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv5x5_tma" -s 3 -c 3 -o gpurun_out/prof_dw_tma -f python scripts/bench_ops.py dwconv > gpurun_out/ncu_dw_tma.log 2>&1; ls -la gpurun_out/prof_dw_tma.ncu-rep
timeout -s KILL 600 python bench.py --workload pgd --pgd-steps 1 --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/p.json 2> gpurun_out/p.err && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_pgd.csv \
    python bench.py --workload pgd --pgd-steps 1 --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_pgd_run.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_pgd.csv 30
