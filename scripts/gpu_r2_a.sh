#!/bin/bash
# round 2, call A: whole GPU suite, default bench line (with extras + incumbent-GPU bar), reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout -s KILL 1500 python -m pytest tests -q -m gpu -x -s -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error|PGD|arg-max|tiny .*bf16|oracle:" gpurun_out/r2a_tests.log | grep -v "fp32\] tiny" | tail -40
echo "== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench default"; timeout -s KILL 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2a_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
    r=d['roofline']; print('tc', r['achieved'], r['frac'], r['share_of_step'])
    for x in r['by_shape']: print('  ', x)
    for k in d['roofline_other_kernels']: print(' other', k['kernel'][:40], round(k['achieved'],1), k['unit'], round(k['frac'],3), round(k['share_of_step'],3))
    print('cpu', d['cpu_baseline'])
    for k in ('strong_scaling','pgd','gender','cars'):
        e=d.get(k,{}); print(k, {kk:e.get(kk) for kk in ('value','ms_per_step','error','hbm_peak_gb')}, (e.get('e2e') or {}).get('value'))
    print('incumbent', json.dumps(d.get('incumbent_gpu'), indent=1))
except Exception as ex:
    print('parse failed', ex)
PY
echo "== bench reference arm"; timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "rc=$?"; cat gpurun_out/r2a_bench_ref.json | cut -c1-600; tail -2 gpurun_out/r2a_bench_ref.err
