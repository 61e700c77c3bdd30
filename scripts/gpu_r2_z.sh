#!/bin/bash
mkdir -p gpurun_out
echo "== mbconv tests"; timeout -s KILL 900 python -m pytest tests/test_mbconv_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
echo "== nvae tests"; timeout -s KILL 900 python -m pytest tests/test_nvae_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
echo "== mbconv"; timeout -s KILL 300 python scripts/bench_ops.py mbconv 2>&1 | tail -3
echo "== trace"; timeout -s KILL 300 python scripts/trace_mbconv.py > gpurun_out/r2z_trace_mbconv.txt 2>&1; tail -42 gpurun_out/r2z_trace_mbconv.txt
echo "== bench"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline 2>&1 >gpurun_out/r2z_bench.json | tail -1
