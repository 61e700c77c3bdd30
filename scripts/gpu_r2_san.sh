#!/bin/bash
mkdir -p gpurun_out
echo "== batch sweep"
for b in 512 768 1024; do timeout -s KILL 600 python bench.py --steps 6 --warmup 3 --extras 0 --no-cpu-baseline --batch $b 2>&1 >gpurun_out/r2san_bench_b$b.json | tail -1; done
echo "== compute-sanitizer memcheck: kernels added or rewritten this round"
timeout -s KILL 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_mbconv_gpu.py tests/test_attacks_gpu.py tests/test_classifier_grad_gpu.py -q -m gpu -x -p no:cacheprovider \
  -k "halo or fused_channel or preprocess or apgd_step or maxpool3x3 or persistent or mbconv" > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r2_sanitize_memcheck.log | tail -8
