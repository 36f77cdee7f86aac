#!/bin/bash
# quick GPU check: gpu tests, parity report, short bench (no CPU leg)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/parity_report.py 2>&1 | tail -6
timeout 600 python bench.py --no-cpu --steps 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms', d['ms_per_step'],'e2e', d['e2e']['value'],'frac', d['roofline']['frac'])"
timeout 300 python scripts/gpu_probe.py 2>&1 | tail -6
