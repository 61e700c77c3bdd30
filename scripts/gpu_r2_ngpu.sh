#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_r2_ngpu.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "rc=$?"
grep "\[bench\]" gpurun_out/r2_bench_${N}gpu.err | head; tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, 'e2e', d['e2e']['value'], d['counters'])
for k in ('strong_scaling','pgd'):
    e=d.get(k,{}); print(k, {kk:e.get(kk) for kk in ('value','ms_per_step','batch_per_gpu','global_batch','error')})
PY
if [ "$N" -le 2 ]; then echo "== reference arm under torchrun"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-300; fi
