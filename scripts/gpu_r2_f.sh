#!/bin/bash
mkdir -p gpurun_out
for cfg in "GA_TC_HALO=1 GA_TC_LEAN=1" "GA_TC_HALO=0 GA_TC_LEAN=1" "GA_TC_HALO=1 GA_TC_LEAN=0"; do
  echo "== $cfg"; env $cfg timeout -s KILL 600 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -k "purifier_gradient or input_gradient_through" -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed"
done
