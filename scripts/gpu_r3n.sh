#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/profile_case.py gaussian 4000 2>&1 | tail -1
timeout 300 python scripts/profile_case.py gaussian 1000 2>&1 | tail -1
