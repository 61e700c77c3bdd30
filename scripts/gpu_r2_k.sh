#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_classifier_grad_gpu.py -q -m gpu -x -s -p no:cacheprovider > gpurun_out/r2k_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error|rel-L2|assert" gpurun_out/r2k_tests.log | tail -20
