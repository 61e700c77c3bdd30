#!/bin/bash
mkdir -p gpurun_out
echo "== trace mbconv"; timeout -s KILL 300 python scripts/trace_mbconv.py 512 > gpurun_out/r2b_trace.txt 2>&1; cat gpurun_out/r2b_trace.txt
echo "== pytest backward+rest"; timeout -s KILL 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error|50-step|arg-max|oracle:" gpurun_out/r2b_tests.log | tail -30
