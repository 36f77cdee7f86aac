#!/bin/bash
# experiment: the library built with -fmad=true (python __graft_entry__.py variant fmad -fmad=true) against the golden vectors, and its speed
mkdir -p gpurun_out
export PAREBEN_LIB=pareben_b200/libpareben_fmad.so
timeout 300 python scripts/parity_report.py 2>&1 | tail -8
timeout 200 python scripts/profile_case.py binomial 2000 2>&1 | tail -1
timeout 200 python scripts/profile_case.py gaussian 4000 2>&1 | tail -1
