#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
PAREBEN_LIB=pareben_b200/libpareben_timing.so timeout 300 python scripts/phase_timing.py gaussian > gpurun_out/r3f_pt_gauss.log 2>&1; head -5 gpurun_out/r3f_pt_gauss.log
timeout 600 python scripts/gram_check.py 1 > gpurun_out/r3f_gram.log 2>&1; tail -4 gpurun_out/r3f_gram.log
PAREBEN_SCHED_TIME=1 timeout 600 python scripts/gram_check.py 1 2>&1 | tail -3
timeout 900 python bench.py --no-cpu --steps 3 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; cat gpurun_out/r3f_bench.json
