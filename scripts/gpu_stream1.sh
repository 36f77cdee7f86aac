#!/bin/bash
# streaming mode, first contact: each test under its own timeout; memcheck of the smallest case if anything fails
mkdir -p gpurun_out
for t in test_stream_config1_slice_matches_oracle_and_cached test_stream_epis_gaussian_pairs test_stream_main_effects_bundled_rows test_stream_real_valued_design test_stream_many_fits_many_tiles; do
  timeout 300 python -m pytest tests/test_gpu_stream.py -x -q -m gpu -k $t > gpurun_out/s1_$t.log 2>&1
  rc=$?
  echo "== $t rc=$rc"; tail -25 gpurun_out/s1_$t.log
  if [ $rc -ne 0 ] && [ -z "$SAN_DONE" ]; then
    SAN_DONE=1
    timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_stream.py -x -q -m gpu -k test_stream_config1 > gpurun_out/s1_memcheck.log 2>&1
    echo "== memcheck rc=$?"; grep -v "^$" gpurun_out/s1_memcheck.log | head -80
  fi
done
