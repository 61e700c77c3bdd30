#!/bin/bash
echo "== stylegan tests"; timeout -s KILL 1500 python -m pytest tests/test_stylegan_gpu.py tests/test_stylegan_paths_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for f in 1 0; do for wl in gender; do echo "== $wl GA_SG_SUPERPIXEL=$f"; GA_SG_SUPERPIXEL=$f timeout -s KILL 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | tail -1; done; done
