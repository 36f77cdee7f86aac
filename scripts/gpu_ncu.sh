#!/bin/bash
# ncu full capture of the fit kernel on the small profile case: $1 = prior, $2 = fits, $3 = output name
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python scripts/profile_case.py $1 $2 > gpurun_out/plain_profile_$3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:eben_fit -s 1 -c 1 -o gpurun_out/$3 python scripts/profile_case.py $1 $2 > gpurun_out/ncu_full_$3.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full_$3.log; cat gpurun_out/plain_profile_$3.log
