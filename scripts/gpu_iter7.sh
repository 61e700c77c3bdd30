#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -k "conv or extras or grad" 2>&1 | tail -3
timeout -s KILL 300 python scripts/bench_ops.py k1 2>&1 | tee gpurun_out/bench_ops_k1_v3.txt
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_purify.json 2> gpurun_out/bench_purify.err; python -c "
import json;d=json.load(open('gpurun_out/bench_purify.json'));print('purify', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/bench_purify.err
timeout -s KILL 900 python bench.py --workload pgd --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pgd.json 2> gpurun_out/bench_pgd.err; python -c "
import json;d=json.load(open('gpurun_out/bench_pgd.json'));print('pgd', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['counters'])"; tail -3 gpurun_out/bench_pgd.err
