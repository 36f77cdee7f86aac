#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed\|Fatal\|Segmentation\|^tests/\|Error" gpurun_out/r2c_pytest.log | head -20; tail -5 gpurun_out/r2c_pytest.log
