#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout -s KILL 900 python -m pytest tests/test_nvae_gpu.py tests/test_mbconv_gpu.py tests/test_backward_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -4
for st in 1 2 3; do for gr in 0 1; do
echo "== bench streams=$st graph=$gr"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline --streams $st --cuda-graph $gr 2>&1 >gpurun_out/r2v_bench_s${st}g${gr}.json | tail -1
done; done
echo "== trace"; timeout -s KILL 300 python scripts/trace_mbconv.py > gpurun_out/r2v_trace_mbconv_f16.txt 2>&1; tail -30 gpurun_out/r2v_trace_mbconv_f16.txt
