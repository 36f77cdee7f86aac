#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python scripts/config5_stream.py 2000 1 10 2>&1 | tail -1
timeout 600 python scripts/config5_stream.py 5000 1 10 2>&1 | tail -1
