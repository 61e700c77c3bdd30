#!/bin/bash
echo "== backward tests"; timeout -s KILL 1200 python -m pytest tests/test_backward_gpu.py tests/test_classifier_grad_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
timeout -s KILL 900 python bench.py --workload pgd --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | tail -1
