#!/bin/bash
mkdir -p gpurun_out
echo "== full GPU suite"; timeout -s KILL 2400 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2full_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2full_tests.log
echo "== smoke"; timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== ops"; timeout -s KILL 600 python scripts/bench_ops.py mbconv > gpurun_out/r2full_ops_mbconv.txt 2>&1; tail -6 gpurun_out/r2full_ops_mbconv.txt
echo "== default bench"; timeout -s KILL 1500 python bench.py > gpurun_out/r2full_bench_default.json 2> gpurun_out/r2full_bench_default.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/r2full_bench_default.err
echo "== pgd bench"; timeout -s KILL 900 python bench.py --workload pgd --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2full_bench_pgd.json 2> gpurun_out/r2full_bench_pgd.err; grep "\[bench\]" gpurun_out/r2full_bench_pgd.err
