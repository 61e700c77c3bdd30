#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"; timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider -k "conv or halo or persistent" > gpurun_out/r2s_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s_tests.log
echo "== kernel tests (P1X1=2)"; GA_TC_P1X1=2 timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider -k "persistent" > gpurun_out/r2s_tests2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s_tests2.log
for i in 1 2; do
echo "== bench"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -1 gpurun_out/r2s_bench.err
echo "== bench (per-tap 1x1)"; GA_TC_P1X1=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/r2s_bench0.json 2> gpurun_out/r2s_bench0.err; tail -1 gpurun_out/r2s_bench0.err
done
echo "== pgd"; timeout -s KILL 600 python bench.py --workload pgd --steps 1 --warmup 1 --batch 512 --no-cpu-baseline > gpurun_out/r2s_pgd.json 2> gpurun_out/r2s_pgd.err; tail -1 gpurun_out/r2s_pgd.err
echo "== pgd (per-tap 1x1)"; GA_TC_P1X1=0 timeout -s KILL 600 python bench.py --workload pgd --steps 1 --warmup 1 --batch 512 --no-cpu-baseline > gpurun_out/r2s_pgd0.json 2> gpurun_out/r2s_pgd0.err; tail -1 gpurun_out/r2s_pgd0.err
