#!/bin/bash
# One gpurun call: kernel unit tests (SIMT and tensor-core in separate processes), engine parity, smoke, short bench.
# Optional extra stages (only if the plain bench exited 0):  ncu = launch list;  full = ncu --set full of the top kernels;
# breakdown = per-op CUDA-event table.
# Usage: gpurun --timeout 1800 -- bash scripts/gpu_check.sh [pgd] [sg] [ncu] [full] [breakdown]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels (non-TC)" ; timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not conv_tc" -p no:cacheprovider > gpurun_out/kernels_simt.log 2>&1; tail -5 gpurun_out/kernels_simt.log
echo "== kernels (TC)" ; timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_tc" -p no:cacheprovider > gpurun_out/kernels_tc.log 2>&1; tail -5 gpurun_out/kernels_tc.log
echo "== fused decoder cell" ; timeout -s KILL 300 python -m pytest tests/test_mbconv_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/mbconv.log 2>&1; grep -E "mbconv n=|passed|failed|FAILED|Error|error" gpurun_out/mbconv.log | tail -12
echo "== nvae parity" ; timeout -s KILL 900 python -m pytest tests/test_nvae_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/nvae.log 2>&1; grep -E "err|passed|failed|FAILED|Error" gpurun_out/nvae.log | grep -v "^tap\|fp32\] tiny" | tail -30
echo "== backward / pgd" ; timeout -s KILL 900 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/backward.log 2>&1; grep -E "grad|PGD|passed|failed|FAILED|Error|error" gpurun_out/backward.log | tail -30
echo "== stylegan" ; timeout -s KILL 600 python -m pytest tests/test_stylegan_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/stylegan.log 2>&1; grep -E "generator@|passed|failed|FAILED|Error|error" gpurun_out/stylegan.log | tail -12
echo "== stylegan paths (configs 3/4)" ; timeout -s KILL 900 python -m pytest tests/test_stylegan_paths_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/stylegan_paths.log 2>&1; grep -E "purified max-abs|passed|failed|FAILED|Error|error|assert" gpurun_out/stylegan_paths.log | tail -25
echo "== ablations / alpha search" ; timeout -s KILL 300 python -m pytest tests/test_ablations_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "== smoke" ; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "== bench" ; timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; rc=$?; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));r=d['roofline'];print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e',d['e2e']['value'],'tc',r['achieved'],r['frac'],r['share_of_step']);[print(x) for x in r['by_shape']];print(d['cpu_baseline'])"; tail -5 gpurun_out/bench.err
for stage in "$@"; do
  if [ "$stage" = "pgd" ]; then
    echo "== pgd bench"
    timeout -s KILL 900 python bench.py --workload pgd --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pgd.json 2> gpurun_out/bench_pgd.err; python -c "
import json;d=json.load(open('gpurun_out/bench_pgd.json'));print({k:d[k] for k in ('metric','value','ms_per_step','gpu_launches')}, d['e2e'], d['counters'], d['roofline']['achieved'])"; tail -5 gpurun_out/bench_pgd.err
  fi
  if [ "$stage" = "sg" ]; then
    for wl in gender cars; do
      echo "== $wl bench"
      timeout -s KILL 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$wl.json'));print({k:d[k] for k in ('metric','value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_$wl.err
    done
  fi
  if [ $rc -ne 0 ]; then break; fi
  if [ "$stage" = "breakdown" ]; then
    echo "== per-op breakdown"
    timeout -s KILL 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --breakdown > gpurun_out/bench_breakdown.json 2> gpurun_out/breakdown.txt; head -45 gpurun_out/breakdown.txt
  fi
  if [ "$stage" = "ncu" ]; then
    echo "== ncu launch list"
    timeout -s KILL 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_plain.json 2> gpurun_out/bench_ncu_plain.err && \
    timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
    wc -l gpurun_out/launches.csv
  fi
  if [ "$stage" = "full" ]; then
    echo "== ncu --set full (top kernels)"
    timeout -s KILL 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_plain.json 2> gpurun_out/bench_ncu_plain.err && \
    timeout -s KILL 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|dwconv5x5_tiled|se_residual" -s 300 -c 24 \
        -o gpurun_out/prof_top -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
    ls -la gpurun_out/prof_top.ncu-rep; tail -2 gpurun_out/ncu_full.log
  fi
done
