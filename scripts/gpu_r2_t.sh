#!/bin/bash
mkdir -p gpurun_out
echo "== persistent 1x1 tests"; timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider -k "persistent" 2>&1 | tail -2
echo "== (P1X1=2)"; GA_TC_P1X1=2 timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider -k "persistent" 2>&1 | tail -2
echo "== two streams"; timeout -s KILL 600 python scripts/two_stream.py 10 2>&1 | tail -6
echo "== pipes"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench_pipes.cu && timeout -s KILL 120 /tmp/ubench | tee gpurun_out/r2t_ubench.txt
echo "== plain run"; timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline > /dev/null 2> gpurun_out/r2t_plain.err; echo "rc=$?"
echo "== ncu mbconv"; timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:mbconv_fused_kernel -s 46 -c 3 -o gpurun_out/r2t_mbconv -f python bench.py --steps 1 --warmup 1 --extras 0 --no-cpu-baseline > gpurun_out/r2t_ncu.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
