#!/bin/bash
# what the driver runs at round end, in its order: whole GPU suite in ONE process, smoke(), reference arm, default bench
mkdir -p gpurun_out
( time timeout -s KILL 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -6
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout -s KILL 900 python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -4 gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err; python -c "
import json;d=json.load(open('gpurun_out/bench_default.json'));print({k:d[k] for k in ('metric','value','steps','warmup','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e'], 'cpu', d['cpu_baseline'])"
wc -l gpurun_out/bench_default.json
