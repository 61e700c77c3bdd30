#!/bin/bash
mkdir -p gpurun_out
echo "== pipes"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench_pipes.cu && timeout -s KILL 120 /tmp/ubench | tee gpurun_out/r2w_ubench.txt | tail -8
echo "== plain run"; timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline --cuda-graph 0 --streams 1 > /dev/null 2> gpurun_out/r2w_plain.err; echo "rc=$?"
echo "== ncu mbconv"; timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:mbconv_fused_kernel -s 34 -c 2 -o gpurun_out/r2w_mbconv -f python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline --cuda-graph 0 --streams 1 > gpurun_out/r2w_ncu.log 2>&1; echo "rc=$?"; ls -la gpurun_out/r2w*.ncu-rep
