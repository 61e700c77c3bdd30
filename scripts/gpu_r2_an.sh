#!/bin/bash
for wl in gender cars; do for st in 1 2; do for gr in 0 1; do
echo "== $wl streams=$st graph=$gr"; timeout -s KILL 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --streams $st --cuda-graph $gr 2>&1 >/dev/null | tail -1
done; done; done
