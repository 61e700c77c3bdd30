#!/bin/bash
# One gpurun call: kernel unit tests (SIMT and tensor-core in separate processes), engine parity, smoke, short bench.
# Usage: gpurun --timeout 1500 -- bash scripts/gpu_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels (non-TC)" ; timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not conv_tc" -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/kernels_simt.log
echo "== kernels (TC)" ; timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_tc" -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/kernels_tc.log
echo "== nvae parity" ; timeout -s KILL 900 python -m pytest tests/test_nvae_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/nvae.log
echo "== smoke" ; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -8 | tee gpurun_out/smoke.log
echo "== bench" ; timeout -s KILL 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
