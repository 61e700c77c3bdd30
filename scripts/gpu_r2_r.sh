#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"; timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider -k "conv or halo or persistent" > gpurun_out/r2r_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2r_tests.log
echo "== k1 (persistent)"; timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tail -4
echo "== k1 (per-tap)"; GA_TC_P1X1=0 timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tail -4
echo "== bench"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -1 gpurun_out/r2r_bench.err
echo "== bench (per-tap 1x1)"; GA_TC_P1X1=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/r2r_bench0.json 2> gpurun_out/r2r_bench0.err; tail -1 gpurun_out/r2r_bench0.err
echo "== pgd"; timeout -s KILL 600 python bench.py --workload pgd --steps 1 --warmup 1 --batch 512 --no-cpu-baseline > gpurun_out/r2r_pgd.json 2> gpurun_out/r2r_pgd.err; tail -1 gpurun_out/r2r_pgd.err
