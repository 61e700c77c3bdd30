#!/bin/bash
# pipe micro-benchmark + graph test + ncu --set full of the depthwise kernel on the PGD path (batch 128)
mkdir -p gpurun_out
scripts/_bin/ubench_pipes > gpurun_out/ubench_pipes.txt 2>&1; cat gpurun_out/ubench_pipes.txt
timeout -s KILL 600 python -m pytest tests/test_backward_gpu.py -q -m gpu -k cuda_graph -p no:cacheprovider 2>&1 | tail -5
timeout -s KILL 600 python bench.py --workload pgd --pgd-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/p.json 2> gpurun_out/p.err && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv5x5_tiled" -s 60 -c 6 -o gpurun_out/prof_dw_pgd -f \
   python bench.py --workload pgd --pgd-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_dw_pgd.log 2>&1
ls -la gpurun_out/prof_dw_pgd.ncu-rep; tail -3 gpurun_out/ncu_dw_pgd.log | cut -c1-300
