#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"; timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_nvae_gpu.py tests/test_edge_cases_gpu.py tests/test_ablations_gpu.py -q -m gpu -x -p no:cacheprovider > gpurun_out/r2j_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2j_tests.log
echo "== hbm ops"; timeout -s KILL 600 python scripts/bench_ops.py hbm 2>&1 | tee gpurun_out/r2j_ops_hbm.txt | head -12
