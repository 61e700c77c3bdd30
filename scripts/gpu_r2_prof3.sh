#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload gender --steps 1 --warmup 1 --no-cpu-baseline"
echo "== plain"; timeout -s KILL 600 $CMD > /dev/null 2> gpurun_out/r2prof3_plain.err; echo "rc=$?"; tail -1 gpurun_out/r2prof3_plain.err
echo "== ncu full"; timeout -s KILL 1500 ncu --set full --clock-control none -k regex:"styled_bias_act_vec8_kernel|torgb_fused_kernel|conv3x3_tc_kernel" -s 30 -c 36 -o /tmp/r2prof3_full -f $CMD > gpurun_out/r2prof3_ncu.log 2>&1; echo "rc=$?"
ncu -i /tmp/r2prof3_full.ncu-rep --page raw --csv > gpurun_out/r2prof3_full_raw.csv 2>/dev/null; ls -la gpurun_out/r2prof3_full_raw.csv
