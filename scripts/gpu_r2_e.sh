#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout -s KILL 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|bf16\].*(err|purified)|smoke" gpurun_out/r2e_tests.log | tail -40
echo "== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -3
