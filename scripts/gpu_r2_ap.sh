#!/bin/bash
mkdir -p gpurun_out
echo "== stylegan tests"; timeout -s KILL 1500 python -m pytest tests/test_stylegan_gpu.py tests/test_stylegan_paths_gpu.py tests/test_edge_cases_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -4
