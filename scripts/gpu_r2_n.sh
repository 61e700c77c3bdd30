#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"; timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "halo or conv_tc or fused_channel" -p no:cacheprovider > gpurun_out/r2n_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2n_tests.log
timeout -s KILL 120 python scripts/trace_conv3x3.py 64 32 512 2>&1 | head -12
timeout -s KILL 120 python scripts/trace_conv3x3.py 128 16 512 2>&1 | head -3
echo "== nvae + backward tests"; timeout -s KILL 900 python -m pytest tests/test_nvae_gpu.py tests/test_backward_gpu.py -q -m gpu -x -s -p no:cacheprovider > gpurun_out/r2n_tests2.log 2>&1; echo "rc=$?"; grep -E "passed|failed|bf16\].*(purified|gradient)" gpurun_out/r2n_tests2.log | tail -12
echo "== bench (no extras)"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "rc=$?"; tail -2 gpurun_out/r2n_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2n_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
    r=d['roofline']; print('tc', r['achieved'], r['frac'], r['share_of_step'])
    for x in r['by_shape'][:6]: print('  ', x)
    for k in d['roofline_other_kernels'][:4]: print(' other', k['kernel'][:40], round(k['achieved'],1), k['unit'], round(k['frac'],3), round(k['share_of_step'],3))
except Exception as ex:
    print('parse failed', ex)
PY
