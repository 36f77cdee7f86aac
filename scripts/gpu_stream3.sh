#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python scripts/config5_stream.py 2000 1 10 > gpurun_out/s3_k2000_full.log 2>&1; tail -1 gpurun_out/s3_k2000_full.log
PAREBEN_TIMING=1 timeout 900 python scripts/config5_stream.py 20000 40 10 > gpurun_out/s3_k20000.log 2>&1; grep "stream round" gpurun_out/s3_k20000.log; tail -1 gpurun_out/s3_k20000.log
