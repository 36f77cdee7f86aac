#!/bin/bash
mkdir -p gpurun_out
PAREBEN_TIMING=1 timeout 600 python scripts/config5_stream.py 5000 1 10 > gpurun_out/s4_k5000_wide.log 2>&1; grep "stream round" gpurun_out/s4_k5000_wide.log | awk '{print $NF, $(NF-1)}' | tr '\n' ' '; echo; tail -1 gpurun_out/s4_k5000_wide.log
PAREBEN_STREAM_WIDE=0 PAREBEN_TIMING=1 timeout 600 python scripts/config5_stream.py 5000 1 10 > gpurun_out/s4_k5000_narrow.log 2>&1; grep "stream round" gpurun_out/s4_k5000_narrow.log | awk '{print $NF, $(NF-1)}' | tr '\n' ' '; echo; tail -1 gpurun_out/s4_k5000_narrow.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_scan -c 1 -o gpurun_out/s4_scan_full -f python scripts/config5_stream.py 2000 1 10 > gpurun_out/s4_ncu.log 2>&1; tail -2 gpurun_out/s4_ncu.log
