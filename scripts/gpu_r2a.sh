#!/bin/bash
# round 2, first check (2-GPU box): tests, 1- and 2-GPU bench (strong scaling), scheduler trace
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench n1 rc=$?"; cat gpurun_out/r2a_bench_n1.json; tail -3 gpurun_out/r2a_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/r2a_bench_n2.json; tail -3 gpurun_out/r2a_bench_n2.err
python __graft_entry__.py timing > gpurun_out/r2a_timing_build.log 2>&1
PAREBEN_LIB=pareben_b200/libpareben_timing.so timeout 600 python scripts/phase_timing.py gaussian > gpurun_out/r2a_pt_gauss.log 2>&1; cat gpurun_out/r2a_pt_gauss.log
PAREBEN_LIB=pareben_b200/libpareben_timing.so timeout 600 python scripts/phase_timing.py binomial > gpurun_out/r2a_pt_binom.log 2>&1; head -5 gpurun_out/r2a_pt_binom.log
