#!/bin/bash
# short bench + gpu tests with an alternative build of the library: $1 = path of the .so
export PAREBEN_LIB=$PWD/$1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --no-cpu --steps 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms', d['ms_per_step'],'e2e', d['e2e']['value'],'frac', d['roofline']['frac'])"
timeout 300 python scripts/gpu_probe.py 2>&1 | tail -2
