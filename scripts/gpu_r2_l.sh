#!/bin/bash
mkdir -p gpurun_out
GA_TC_HALO=0 timeout -s KILL 120 python scripts/trace_conv3x3.py 256 8 512 2>&1 | head -1
GA_TC_HALO=0 GA_TC_BLOCK_N=256 timeout -s KILL 120 python scripts/trace_conv3x3.py 256 8 512 2>&1 | head -1
GA_TC_HALO=0 GA_TC_BLOCK_N=64 timeout -s KILL 120 python scripts/trace_conv3x3.py 256 8 512 2>&1 | head -1
GA_TC_HALO=0 GA_TC_SHORT_KB=100 timeout -s KILL 120 python scripts/trace_conv3x3.py 256 8 512 2>&1 | head -1
timeout -s KILL 300 python -m pytest tests/test_classifier_grad_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
