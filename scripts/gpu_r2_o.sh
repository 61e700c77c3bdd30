#!/bin/bash
mkdir -p gpurun_out
echo "== trace mbconv (after elect)"; timeout -s KILL 300 python scripts/trace_mbconv.py 512 > gpurun_out/r2o_trace.txt 2>&1; head -40 gpurun_out/r2o_trace.txt
for s in 0 20; do echo "== sleep_ns $s"; GA_MB_SLEEP_NS=$s timeout -s KILL 300 python scripts/trace_mbconv.py 512 2>&1 | head -1; done
echo "== act_hi"; GA_MB_ACT_HI=1 timeout -s KILL 300 python scripts/trace_mbconv.py 512 2>&1 | head -1
echo "== 64-image counts + classifier grads"; timeout -s KILL 900 python -m pytest tests/test_stylegan_paths_gpu.py tests/test_classifier_grad_gpu.py -q -m gpu -s -k "clean_counts or classifier_input" -p no:cacheprovider 2>&1 | grep -E "passed|failed|arg-max|rel-L2|Error" | tail -12
