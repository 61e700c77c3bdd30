#!/bin/bash
mkdir -p gpurun_out
echo "== pipes"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench_pipes.cu && timeout -s KILL 120 /tmp/ubench | tee gpurun_out/r2u_ubench.txt
for f in 0 1; do
echo "== GA_MB_F16=$f"
export GA_MB_F16=$f
timeout -s KILL 600 python -m pytest tests/test_mbconv_gpu.py tests/test_nvae_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -E "mbconv n=|purified max|passed|failed|bf16\]" | tail -30
timeout -s KILL 300 python scripts/bench_ops.py mbconv 2>&1 | tail -4
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline 2>&1 >/dev/null | tail -1
done
unset GA_MB_F16
echo "== two streams"; timeout -s KILL 600 python scripts/two_stream.py 10 2>&1 | tail -6
