#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"; cat gpurun_out/r2_bench_n$N.json; tail -2 gpurun_out/r2_bench_n$N.err
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "two_devices or sl_prefilter_reproduces" 2>&1 | tail -3; timeout 300 python scripts/worst_binomial_fit.py; fi
