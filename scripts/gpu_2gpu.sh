#!/bin/bash
# 2-GPU weak-scaling check of both NVAE workloads, launched the way the driver does
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; python -c "
import json;d=json.load(open('gpurun_out/bench_2gpu.json'));print('purify x2', {k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['counters'])"; tail -3 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload pgd --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pgd_2gpu.json 2> gpurun_out/bench_pgd_2gpu.err; python -c "
import json;d=json.load(open('gpurun_out/bench_pgd_2gpu.json'));print('pgd x2', {k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['counters'])"; tail -3 gpurun_out/bench_pgd_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>&1 | tail -2 | cut -c1-400
