#!/bin/bash
mkdir -p gpurun_out
echo "== se_residual at <= 48 registers (co-resident with the fused cell?)"
for b in 512 1024; do timeout -s KILL 600 python bench.py --steps 6 --warmup 3 --extras 0 --no-cpu-baseline --batch $b 2>&1 >gpurun_out/r2aa_bench_b$b.json | tail -1; done
timeout -s KILL 600 python bench.py --steps 6 --warmup 3 --extras 0 --no-cpu-baseline --batch 2048 2>&1 >gpurun_out/r2aa_bench_b2048.json | tail -1
