#!/bin/bash
echo "== kernel tests"; timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -4
