#!/bin/bash
# round 2 (session 3), N-GPU box: strong-scaling bench line + one-call fan-out
mkdir -p gpurun_out
N=$1
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/r3_bench_n$N.json 2> gpurun_out/r3_bench_n$N.err; echo "bench n$N rc=$?"; cat gpurun_out/r3_bench_n$N.json; tail -2 gpurun_out/r3_bench_n$N.err
timeout 300 python scripts/one_call_devices.py $N > gpurun_out/r3_one_call_n$N.log 2>&1; cat gpurun_out/r3_one_call_n$N.log
