#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
for t in 64 128 256; do
echo "== threads $t"; PAREBEN_THREADS=$t timeout 600 python bench.py --no-cpu --steps 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
PAREBEN_THREADS=$t timeout 300 python scripts/gpu_probe.py 2>&1 | tail -5 | grep -v "fits=1200\|fits=54" | tail -1
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
