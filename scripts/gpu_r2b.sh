#!/bin/bash
# round 2: whole GPU suite, smoke, 1-GPU bench with the streaming extra
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench n1 rc=$?"; cat gpurun_out/r2b_bench_n1.json; tail -3 gpurun_out/r2b_bench_n1.err
