#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "tf32 or conv_tc" 2>&1 | tail -6
timeout -s KILL 900 python -m pytest tests/test_stylegan_paths_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/stylegan_paths.log 2>&1; grep -E "purified max-abs|passed|failed|FAILED|Error|error|assert" gpurun_out/stylegan_paths.log | tail -12
timeout -s KILL 600 python scripts/diag_e4e_bf16.py 2>&1 | tail -6
timeout -s KILL 900 python bench.py --workload gender --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_gender.json 2> gpurun_out/bench_gender.err; python -c "
import json;d=json.load(open('gpurun_out/bench_gender.json'));print('gender tf32 backbone', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/bench_gender.err
GA_E4E_TF32_BACKBONE=0 timeout -s KILL 900 python bench.py --workload gender --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_gender_bf16bb.json 2> gpurun_out/bench_gender_bf16bb.err; python -c "
import json;d=json.load(open('gpurun_out/bench_gender_bf16bb.json'));print('gender bf16 backbone', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/bench_gender_bf16bb.err
