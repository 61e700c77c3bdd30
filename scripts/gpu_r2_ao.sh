#!/bin/bash
mkdir -p gpurun_out
for wl in gender cars; do timeout -s KILL 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --breakdown > /dev/null 2> gpurun_out/r2ao_breakdown_$wl.txt; grep -v "^\[" gpurun_out/r2ao_breakdown_$wl.txt | head -32; done
