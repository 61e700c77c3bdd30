#!/bin/bash
mkdir -p gpurun_out
echo "== pipes"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench_pipes.cu && timeout -s KILL 120 /tmp/ubench | tee gpurun_out/r2x_ubench.txt | grep MUFU
for st in 0 1200 2000 2800 3600; do
echo "== GA_MB_STAGGER=$st"; GA_MB_STAGGER=$st timeout -s KILL 300 python scripts/bench_ops.py mbconv 2>&1 | tail -3
done
echo "== mbconv tests, stagger 2000"; GA_MB_STAGGER=2000 timeout -s KILL 600 python -m pytest tests/test_mbconv_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
