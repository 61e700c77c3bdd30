#!/bin/bash
mkdir -p gpurun_out
echo "== mbconv tests"; timeout -s KILL 600 python -m pytest tests/test_mbconv_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
echo "== mbconv"; timeout -s KILL 300 python scripts/bench_ops.py mbconv 2>&1 | tail -3
for b in 512 1024; do for i in 1 2; do timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline --batch $b 2>&1 >gpurun_out/r2ac_bench_b$b.json | tail -1; done; done
echo "== streams 1"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline --streams 1 2>&1 >/dev/null | tail -1
