#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/imma_check.py > gpurun_out/r3i_imma.log 2>&1; cat gpurun_out/r3i_imma.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
PAREBEN_LIB=pareben_b200/libpareben_timing.so timeout 300 python scripts/phase_timing.py binomial > gpurun_out/r3i_pt_binom.log 2>&1; cat gpurun_out/r3i_pt_binom.log
