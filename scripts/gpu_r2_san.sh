#!/bin/bash
mkdir -p gpurun_out
echo "== compute-sanitizer memcheck: new kernels"
timeout -s KILL 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_attacks_gpu.py tests/test_classifier_grad_gpu.py -q -m gpu -x -p no:cacheprovider \
  -k "halo or fused_channel or preprocess or apgd_step or maxpool3x3" > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r2_sanitize_memcheck.log | tail -8
