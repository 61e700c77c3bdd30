#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline"
echo "== plain"; timeout -s KILL 600 $CMD > gpurun_out/r2prof_plain.json 2> gpurun_out/r2prof_plain.err; echo "rc=$?"
echo "== launch list"; timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2prof_launches.csv $CMD > gpurun_out/r2prof_ncu_list.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2prof_launches.csv
echo "== reference arm"; timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2prof_reference.json 2> gpurun_out/r2prof_reference.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2prof_reference.json
