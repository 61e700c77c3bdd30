#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout -s KILL 600 python -m pytest tests/test_mbconv_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
timeout -s KILL 600 python scripts/bench_ops.py mbconv 2>&1 | grep "attack path"
timeout -s KILL 300 python scripts/trace_mbconv_bwd.py > gpurun_out/r2aj_trace_bwd.txt 2>&1; grep -A7 "dw warp 5" gpurun_out/r2aj_trace_bwd.txt; grep -A5 "act warp 1" gpurun_out/r2aj_trace_bwd.txt
