#!/bin/bash
# N-GPU weak-scaling check of both NVAE workloads, launched the way the driver does: bash scripts/gpu_ngpu.sh N  (gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; python -c "
import json;d=json.load(open('gpurun_out/bench_${N}gpu.json'));print('purify x$N', {k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['counters'])"; tail -3 gpurun_out/bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload pgd --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pgd_${N}gpu.json 2> gpurun_out/bench_pgd_${N}gpu.err; python -c "
import json;d=json.load(open('gpurun_out/bench_pgd_${N}gpu.json'));print('pgd x$N', {k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['counters'])"; tail -3 gpurun_out/bench_pgd_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --impl reference --steps 1 --warmup 0 2>&1 | tail -2 | cut -c1-400
