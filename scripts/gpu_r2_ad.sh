#!/bin/bash
echo "== SE 256"; timeout -s KILL 300 python scripts/overlap_probe.py 2>&1 | tail -3
echo "== SE 128"; GA_SE_THREADS=128 timeout -s KILL 300 python scripts/overlap_probe.py 2>&1 | tail -3
for i in 1 2; do GA_SE_THREADS=128 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline 2>&1 >/dev/null | tail -1; done
for i in 1 2; do timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline 2>&1 >/dev/null | tail -1; done
