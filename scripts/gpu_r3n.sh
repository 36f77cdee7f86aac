#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gram.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/profile_case.py gaussian 4000 2>&1 | tail -1
timeout 300 python scripts/profile_case.py binomial 2000 2>&1 | tail -1
