#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"; timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_stylegan_paths_gpu.py -q -m gpu -x -p no:cacheprovider -k "conv or halo" > gpurun_out/r2q_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2q_tests.log
echo "== k1 (lean)"; timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tail -6
echo "== k1 (generic)"; GA_TC_LEAN_1X1=0 timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tail -6
echo "== bench"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -1 gpurun_out/r2q_bench.err
echo "== bench (generic 1x1)"; GA_TC_LEAN_1X1=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/r2q_bench0.json 2> gpurun_out/r2q_bench0.err; tail -1 gpurun_out/r2q_bench0.err
echo "== pgd"; timeout -s KILL 600 python bench.py --workload pgd --steps 1 --warmup 1 --batch 512 --no-cpu-baseline > gpurun_out/r2q_pgd.json 2> gpurun_out/r2q_pgd.err; tail -1 gpurun_out/r2q_pgd.err
echo "== pgd (generic 1x1)"; GA_TC_LEAN_1X1=0 timeout -s KILL 600 python bench.py --workload pgd --steps 1 --warmup 1 --batch 512 --no-cpu-baseline > gpurun_out/r2q_pgd0.json 2> gpurun_out/r2q_pgd0.err; tail -1 gpurun_out/r2q_pgd0.err
