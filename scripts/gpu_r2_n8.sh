#!/bin/bash
# round 2, 8-GPU box: strong-scaling bench line (config 2 + Gaussian extra + config 5 at full size through the streaming kernels)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"; cat gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err
timeout 300 python scripts/one_call_devices.py 8 > gpurun_out/r2_one_call_n8.log 2>&1; cat gpurun_out/r2_one_call_n8.log
