#!/bin/bash
# quick iteration: dwconv / backward / nvae parity tests, then purify + pgd benches and the pgd per-op breakdown
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dwconv" -p no:cacheprovider 2>&1 | tail -8
timeout -s KILL 600 python -m pytest tests/test_backward_gpu.py tests/test_nvae_gpu.py tests/test_mbconv_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_purify.json 2> gpurun_out/bench_purify.err; python -c "
import json;d=json.load(open('gpurun_out/bench_purify.json'));print('purify', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/bench_purify.err
timeout -s KILL 900 python bench.py --workload pgd --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pgd.json 2> gpurun_out/bench_pgd.err; python -c "
import json;d=json.load(open('gpurun_out/bench_pgd.json'));print('pgd', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['counters'])"; tail -3 gpurun_out/bench_pgd.err
timeout -s KILL 600 python bench.py --workload pgd --pgd-steps 5 --steps 1 --warmup 1 --no-cpu-baseline --breakdown --cuda-graph 0 > gpurun_out/bench_pgd_bd.json 2> gpurun_out/breakdown_pgd.txt; head -40 gpurun_out/breakdown_pgd.txt; tail -2 gpurun_out/breakdown_pgd.txt
timeout -s KILL 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --breakdown > gpurun_out/bench_breakdown.json 2> gpurun_out/breakdown.txt; head -30 gpurun_out/breakdown.txt; tail -2 gpurun_out/breakdown.txt
