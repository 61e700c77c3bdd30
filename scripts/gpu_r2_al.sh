#!/bin/bash
for v in 16 32 64 128 256 512; do echo "== GA_SE_APPLY_PPB=$v"; GA_SE_APPLY_PPB=$v timeout -s KILL 300 python scripts/bench_ops.py hbm 2>&1 | grep -i "se_residual " ; done
