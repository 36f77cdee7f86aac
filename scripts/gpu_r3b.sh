#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gram_check.py 1 > gpurun_out/r3b_gram.log 2>&1; cat gpurun_out/r3b_gram.log
PAREBEN_GRAM=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
