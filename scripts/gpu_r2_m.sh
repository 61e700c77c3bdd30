#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline > gpurun_out/r2m_plain.json 2> gpurun_out/r2m_plain.err && \
timeout -s KILL 1500 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc_kernel|mbconv_fused_kernel|se_residual_kernel" -s 40 -c 12 \
    -o gpurun_out/r2m_prof -f python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline > gpurun_out/r2m_ncu.log 2>&1
ls -la gpurun_out/r2m_prof.ncu-rep; tail -3 gpurun_out/r2m_ncu.log
