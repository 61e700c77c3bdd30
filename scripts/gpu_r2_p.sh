#!/bin/bash
mkdir -p gpurun_out
echo "== mbconv tests (dw16)"; GA_MB_DW16=1 timeout -s KILL 600 python -m pytest tests/test_mbconv_gpu.py -q -m gpu -x -s -p no:cacheprovider 2>&1 | grep -E "mbconv n=|passed|failed|Error" | tail -12
echo "== trace dw16"; GA_MB_DW16=1 timeout -s KILL 300 python scripts/trace_mbconv.py 512 > gpurun_out/r2p_trace16.txt 2>&1; head -44 gpurun_out/r2p_trace16.txt
echo "== bench ops mbconv"; GA_MB_DW16=1 timeout -s KILL 300 python scripts/bench_ops.py mbconv 2>&1 | tail -8
echo "== (8-warp reference)"; timeout -s KILL 300 python scripts/bench_ops.py mbconv 2>&1 | tail -8
