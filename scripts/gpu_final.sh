#!/bin/bash
# Round-end style run: tests, smoke, parity report, bench (+ reference arm), ncu launch list, full-size ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/parity_report.py > gpurun_out/parity_stdout.log 2>&1; tail -9 gpurun_out/parity_stdout.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --stream-k 2000 > gpurun_out/plain_bench.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --stream-k 2000 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 300 python scripts/profile_case.py binomial 2000 > gpurun_out/plain_profile_full.log 2>&1 &&
timeout 1700 ncu --set full --clock-control none --import-source on -k regex:eben_fit -s 1 -c 1 -o gpurun_out/prof_binom_full -f python scripts/profile_case.py binomial 2000 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log; cat gpurun_out/plain_profile_full.log
timeout 300 python scripts/config5_stream.py 2000 1 10 > gpurun_out/plain_stream_k2000.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:stream_scan -c 1 -o gpurun_out/prof_stream_scan -f python scripts/config5_stream.py 2000 1 10 > gpurun_out/ncu_stream.log 2>&1
echo "stream capture rc=$?"; tail -1 gpurun_out/plain_stream_k2000.log
