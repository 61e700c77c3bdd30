#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py tests/test_nvae_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
for v in 512 128; do
echo "== GA_SE_APPLY_PPB=$v GA_SE_BWD_APPLY_PPB=$v"
GA_SE_APPLY_PPB=$v GA_SE_BWD_APPLY_PPB=$v timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline 2>&1 >/dev/null | tail -1
GA_SE_APPLY_PPB=$v GA_SE_BWD_APPLY_PPB=$v timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline 2>&1 >/dev/null | tail -1
GA_SE_APPLY_PPB=$v GA_SE_BWD_APPLY_PPB=$v timeout -s KILL 900 python bench.py --workload pgd --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | tail -1
done
