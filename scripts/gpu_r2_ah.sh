#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python scripts/bench_ops.py mbconv 2>&1 | tee gpurun_out/r2ah_ops_mbconv.txt | tail -8
