#!/bin/bash
mkdir -p gpurun_out
echo "== pgd breakdown"; timeout -s KILL 600 python bench.py --workload pgd --steps 1 --warmup 1 --batch 512 --no-cpu-baseline --breakdown > gpurun_out/r2i_pgd.json 2> gpurun_out/r2i_pgd_breakdown.txt; head -64 gpurun_out/r2i_pgd_breakdown.txt
echo "== ncu launch list"
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline > gpurun_out/r2i_plain.json 2> gpurun_out/r2i_plain.err && \
timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2i_launches.csv \
    python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline > gpurun_out/r2i_ncu_run.log 2>&1
wc -l gpurun_out/r2i_launches.csv
