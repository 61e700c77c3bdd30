#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline --cuda-graph 0 --streams 1 --batch 512"
echo "== plain"; timeout -s KILL 600 $CMD > /dev/null 2> gpurun_out/r2prof2_plain.err; echo "rc=$?"
echo "== ncu full"; timeout -s KILL 1500 ncu --set full --clock-control none -k regex:"mbconv_fused_kernel" -s 12 -c 22 -o /tmp/r2prof2_full -f $CMD > gpurun_out/r2prof2_ncu.log 2>&1; echo "rc=$?"; ls -la /tmp/r2prof2_full.ncu-rep
ncu -i /tmp/r2prof2_full.ncu-rep --page raw --csv > gpurun_out/r2prof2_full_raw.csv 2>/dev/null; ls -la gpurun_out/r2prof2_full_raw.csv
