#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q -m gpu 2>&1 | tail -15
PAREBEN_TIMING=1 timeout 600 python scripts/config5_stream.py 2000 10 10 > gpurun_out/s2_k2000.log 2>&1; grep "stream round" gpurun_out/s2_k2000.log | awk '{print $5}' | tr '\n' ' ' | cut -c1-1500; echo; tail -1 gpurun_out/s2_k2000.log
timeout 600 python scripts/config5_stream.py 5000 10 10 > gpurun_out/s2_k5000.log 2>&1; tail -1 gpurun_out/s2_k5000.log
PAREBEN_TIMING=1 timeout 300 python scripts/config5_stream.py 600 10 10 1000 main > gpurun_out/s2_k600_main.log 2>&1; grep "stream round" gpurun_out/s2_k600_main.log | awk '{print $5}' | tr '\n' ' ' | cut -c1-1500; echo; tail -1 gpurun_out/s2_k600_main.log
