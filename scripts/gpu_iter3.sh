#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_backward_gpu.py tests/test_mbconv_gpu.py tests/test_nvae_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -6
for hi in 1 0; do for ns in 100 20; do echo "== GA_MB_ACT_HI=$hi GA_MB_SLEEP_NS=$ns"; GA_MB_ACT_HI=$hi GA_MB_SLEEP_NS=$ns timeout -s KILL 300 python scripts/bench_ops.py mbconv 2>&1 | tail -3; done; done | tee gpurun_out/bench_ops_mbconv.txt
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_purify.json 2> gpurun_out/bench_purify.err; python -c "
import json;d=json.load(open('gpurun_out/bench_purify.json'));print('purify', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/bench_purify.err
for b in 128 256 512; do
timeout -s KILL 900 python bench.py --workload pgd --batch $b --pgd-steps 10 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pgd_b$b.json 2> gpurun_out/bench_pgd_b$b.err; python -c "
import json;d=json.load(open('gpurun_out/bench_pgd_b$b.json'));print('pgd batch $b (10 steps)', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['counters'])"; tail -3 gpurun_out/bench_pgd_b$b.err
done
nvidia-smi --query-gpu=memory.used --format=csv
