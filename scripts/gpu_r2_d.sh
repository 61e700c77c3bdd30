#!/bin/bash
mkdir -p gpurun_out
echo "== halo kernel tests"; timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "halo or conv_tc" -p no:cacheprovider > gpurun_out/r2c_halo.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2c_halo.log
timeout -s KILL 120 python scripts/trace_conv3x3.py 64 32 512 2>&1 | tee gpurun_out/r2d_trace_c64.txt
timeout -s KILL 120 python scripts/trace_conv3x3.py 128 16 512 2>&1 | tee gpurun_out/r2d_trace_c128.txt
GA_TC_HALO=0 timeout -s KILL 120 python scripts/trace_conv3x3.py 64 32 512 2>&1 | head -1
GA_TC_HALO=0 timeout -s KILL 120 python scripts/trace_conv3x3.py 128 16 512 2>&1 | head -1
GA_TC_HALO=0 timeout -s KILL 120 python scripts/trace_conv3x3.py 256 8 512 2>&1 | head -1
echo "== bench (no extras)"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
    r=d['roofline']; print('tc', r['achieved'], r['frac'], r['share_of_step'])
    for x in r['by_shape']: print('  ', x)
except Exception as ex:
    print('parse failed', ex)
PY
