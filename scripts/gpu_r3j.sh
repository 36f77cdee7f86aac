#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python scripts/epis_scale.py 200 10 16 > gpurun_out/r3j_epis.log 2>&1; tail -2 gpurun_out/r3j_epis.log
timeout 900 python scripts/config3_scale.py 8 > gpurun_out/r3j_config3.log 2>&1; tail -5 gpurun_out/r3j_config3.log
timeout 900 python bench.py --no-cpu --steps 3 --no-stream > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; cat gpurun_out/r3j_bench.json
