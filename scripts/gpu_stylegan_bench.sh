#!/bin/bash
# One gpurun call: StyleGAN path parity tests + benches of BASELINE configs[2] (gender) and configs[3] (cars) with per-op breakdown.
# Usage: gpurun --timeout 1500 -- bash scripts/gpu_stylegan_bench.sh [purify]
mkdir -p gpurun_out
echo "== stylegan paths (configs 3/4)" ; timeout -s KILL 900 python -m pytest tests/test_stylegan_paths_gpu.py tests/test_stylegan_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/stylegan_paths.log 2>&1; grep -E "purified max-abs|generator@|passed|failed|FAILED|Error|error|assert" gpurun_out/stylegan_paths.log | tail -25
for wl in "$@"; do
  echo "== bench $wl"
  extra=""; [ "$wl" = "purify" ] || extra="--breakdown"
  timeout -s KILL 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline $extra > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$wl.json'));r=d['roofline'];print({k:d[k] for k in ('value','ms_per_step','gpu_launches','tflops_algorithmic')}, 'e2e',d['e2e']['value'],'tc',r['achieved'],r['frac'],r['share_of_step']);[print(x) for x in r['by_shape']]"; head -42 gpurun_out/bench_$wl.err
done
