#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --workload pgd --pgd-steps 3 --steps 1 --warmup 1 --no-cpu-baseline --breakdown --cuda-graph 0 > gpurun_out/bench_pgd_bd.json 2> gpurun_out/breakdown_pgd.txt; head -48 gpurun_out/breakdown_pgd.txt; tail -2 gpurun_out/breakdown_pgd.txt
