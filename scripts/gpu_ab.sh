#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== RC16"; timeout 600 python bench.py --no-cpu --steps 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
echo "== RC8"; PAREBEN_LIB=$PWD/pareben_b200/libpareben_rc8.so timeout 600 python bench.py --no-cpu --steps 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
timeout 300 python scripts/gpu_probe.py 2>&1 | tail -7
