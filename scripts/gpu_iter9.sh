#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 1" "1 1"; do set -- $cfg
echo "== GA_E4E_TF32_BACKBONE=$1 GA_E4E_TF32_HEADS=$2"
GA_E4E_TF32_BACKBONE=$1 GA_E4E_TF32_HEADS=$2 timeout -s KILL 600 python scripts/diag_e4e_bf16.py 2>&1 | tail -5
GA_E4E_TF32_BACKBONE=$1 GA_E4E_TF32_HEADS=$2 timeout -s KILL 900 python bench.py --workload gender --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_gender_$1$2.json 2> gpurun_out/bench_gender_$1$2.err; python -c "
import json;d=json.load(open('gpurun_out/bench_gender_$1$2.json'));print('gender', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/bench_gender_$1$2.err
done
