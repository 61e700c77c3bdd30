#!/bin/bash
# ncu --set full of a few launches of the three kernel classes of the purify step (small reports: gpurun_out/ is capped at 64 MiB):
#   conv_tc (first encoder-tower launches), se_residual, mbconv_fused (one per scale)
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_plain.json 2> gpurun_out/bench_ncu_plain.err || exit 1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"conv_tc_kernel" -s 10 -c 8 -o gpurun_out/prof_conv_tc -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"se_residual_kernel" -s 10 -c 3 -o gpurun_out/prof_se -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"mbconv_fused" -s 13 -c 2 -o gpurun_out/prof_mbconv_a -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"mbconv_fused" -s 30 -c 1 -o gpurun_out/prof_mbconv_b -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"dwconv5x5_tma" -s 4 -c 3 -o gpurun_out/prof_dwconv_tma -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep
