#!/bin/bash
mkdir -p gpurun_out
echo "== mbconv tests"; timeout -s KILL 600 python -m pytest tests/test_mbconv_gpu.py tests/test_nvae_gpu.py tests/test_edge_cases_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
echo "== mbconv"; timeout -s KILL 300 python scripts/bench_ops.py mbconv 2>&1 | tail -3
for i in 1 2; do timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline 2>&1 >gpurun_out/r2ae_bench.json | tail -1; done
for i in 1 2; do GA_FUSE_CSUM=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline 2>&1 >/dev/null | tail -1; done
