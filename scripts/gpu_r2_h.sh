#!/bin/bash
mkdir -p gpurun_out
echo "== attacks + NF/C64 tests"; timeout -s KILL 900 python -m pytest tests/test_attacks_gpu.py tests/test_nvae_gpu.py -q -m gpu -s -p no:cacheprovider -k "attacks or apgd or fgsm or normalizing or c64" > gpurun_out/r2h_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error|APGD|FGSM|NF checkpoint|C64" gpurun_out/r2h_tests.log | tail -40
