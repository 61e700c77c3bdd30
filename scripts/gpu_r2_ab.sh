#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout -s KILL 1200 python -m pytest tests/test_backward_gpu.py tests/test_nvae_gpu.py tests/test_attacks_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for st in 1 2; do
echo "== pgd streams=$st"; timeout -s KILL 900 python bench.py --workload pgd --steps 1 --warmup 1 --no-cpu-baseline --streams $st 2>&1 >gpurun_out/r2ab_pgd_s$st.json | tail -1
done
