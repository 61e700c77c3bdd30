#!/bin/bash
mkdir -p gpurun_out
echo "== backward tests"; timeout -s KILL 1200 python -m pytest tests/test_backward_gpu.py tests/test_mbconv_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for fb in 1 2; do echo "== pgd GA_FUSE_BWD=$fb"; GA_FUSE_BWD=$fb timeout -s KILL 900 python bench.py --workload pgd --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >gpurun_out/r2ai_pgd_fb$fb.json | tail -1; done
