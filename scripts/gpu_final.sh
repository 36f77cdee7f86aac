#!/bin/bash
# Round-end style run, in parts so that each call's gpurun_out/ stays under the 64 MiB that travel back:
#   bash scripts/gpu_final.sh tests      gpu tests, smoke
#   bash scripts/gpu_final.sh bench      parity report, bench line (+ launch list under ncu)
#   bash scripts/gpu_final.sh binom      ncu --set full of the fit kernel on config 2 (2,000 fits) and of the streaming scan
#   bash scripts/gpu_final.sh gauss      ncu --set full of the fit kernel on the bundled Gaussian design (1,000 fits)
# then here: python scripts/make_profiles.py r02 ; python scripts/make_profiles.py r02 gpurun_out/prof_gauss_full.ncu-rep gaussian "<workload>"
mkdir -p gpurun_out
part=${1:-all}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
if [ "$part" = tests ] || [ "$part" = all ]; then
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
if [ "$part" = bench ] || [ "$part" = all ]; then
timeout 300 python scripts/parity_report.py > gpurun_out/parity_stdout.log 2>&1; tail -9 gpurun_out/parity_stdout.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
# (reference arm: profiles/r02_bench_reference_n1.json, CPU code unchanged since)
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --stream-k 2000 > gpurun_out/plain_bench.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --stream-k 2000 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
fi
if [ "$part" = binom ] || [ "$part" = all ]; then
timeout 300 python scripts/profile_case.py binomial 2000 > gpurun_out/plain_profile_full.log 2>&1 &&
timeout 1700 ncu --set full --clock-control none --import-source on -k regex:eben_fit -s 1 -c 1 -o gpurun_out/prof_binom_full -f python scripts/profile_case.py binomial 2000 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log; cat gpurun_out/plain_profile_full.log
timeout 300 python scripts/config5_stream.py 2000 1 10 > gpurun_out/plain_stream_k2000.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:stream_scan -c 1 -o gpurun_out/prof_stream_scan -f python scripts/config5_stream.py 2000 1 10 > gpurun_out/ncu_stream.log 2>&1
echo "stream capture rc=$?"; tail -1 gpurun_out/plain_stream_k2000.log
fi
if [ "$part" = gauss ] || [ "$part" = all ]; then
timeout 300 python scripts/profile_case.py gaussian 1000 > gpurun_out/plain_profile_gauss.log 2>&1 &&
timeout 1700 ncu --set full --clock-control none --import-source on -k regex:eben_fit -s 1 -c 1 -o gpurun_out/prof_gauss_full -f python scripts/profile_case.py gaussian 1000 > gpurun_out/ncu_gauss.log 2>&1
echo "gaussian capture rc=$?"; tail -1 gpurun_out/ncu_gauss.log; cat gpurun_out/plain_profile_gauss.log
fi
ls -la gpurun_out | head -30; du -sm gpurun_out
