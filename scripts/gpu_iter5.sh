#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_backward_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "conv or extras or grad" 2>&1 | tail -4
timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tee gpurun_out/bench_ops_k1.txt
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 16 -c 5 -o gpurun_out/prof_k1 -f python scripts/bench_ops.py k1 > gpurun_out/ncu_k1.log 2>&1; ls -la gpurun_out/prof_k1.ncu-rep
