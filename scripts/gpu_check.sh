#!/bin/bash
# One gpurun call: kernel unit tests (SIMT and tensor-core in separate processes), engine parity, smoke, short bench,
# then (only if the bench exited 0) the ncu launch list of the same bench command.
# Usage: gpurun --timeout 1500 -- bash scripts/gpu_check.sh [ncu]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels (non-TC)" ; timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not conv_tc" -p no:cacheprovider > gpurun_out/kernels_simt.log 2>&1; tail -5 gpurun_out/kernels_simt.log
echo "== kernels (TC)" ; timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_tc" -p no:cacheprovider > gpurun_out/kernels_tc.log 2>&1; tail -5 gpurun_out/kernels_tc.log
echo "== nvae parity" ; timeout -s KILL 900 python -m pytest tests/test_nvae_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/nvae.log 2>&1; grep -E "err|passed|failed|FAILED|Error" gpurun_out/nvae.log | tail -40
echo "== smoke" ; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "== bench" ; timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; rc=$?; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" = "ncu" ] && [ $rc -eq 0 ]; then
  echo "== ncu launch list"
  timeout -s KILL 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_plain.json 2> gpurun_out/bench_ncu_plain.err && \
  timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
  tail -3 gpurun_out/ncu_run.log; wc -l gpurun_out/launches.csv
fi
