#!/bin/bash
# ncu --set full of ONE launch of the fused decoder-cell kernel at 32x32 (launch 31 of the first forward).  Keep reports small
# (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_plain.json 2> gpurun_out/bench_ncu_plain.err || exit 1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"mbconv_fused" -s ${1:-30} -c 1 \
    -o gpurun_out/prof_mbconv_b -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mbconv_b.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_mbconv_b.log | cut -c1-200
