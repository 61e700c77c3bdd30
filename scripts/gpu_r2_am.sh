#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -v "^\.*$" | grep -i -E "cos|rel|err|match|flag|count|passed|failed" | head -40
