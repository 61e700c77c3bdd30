#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "conv_tc" 2>&1 | tail -3
echo "== default"; timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tee gpurun_out/bench_ops_k1_v2.txt
echo "== GA_TC_BLOCK_N=64"; GA_TC_BLOCK_N=64 timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tee gpurun_out/bench_ops_k1_n64.txt
echo "== GA_TC_BLOCK_N=256"; GA_TC_BLOCK_N=256 timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tee gpurun_out/bench_ops_k1_n256.txt
