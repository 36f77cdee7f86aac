#!/bin/bash
# round-2 session-3 baseline: gpu tests, phase timing (both priors)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PAREBEN_LIB=pareben_b200/libpareben_timing.so timeout 300 python scripts/phase_timing.py gaussian > gpurun_out/r3a_pt_gauss.log 2>&1; cat gpurun_out/r3a_pt_gauss.log
PAREBEN_LIB=pareben_b200/libpareben_timing.so timeout 300 python scripts/phase_timing.py binomial > gpurun_out/r3a_pt_binom.log 2>&1; cat gpurun_out/r3a_pt_binom.log
