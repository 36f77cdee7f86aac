#!/bin/bash
# DRAM traffic / L2 hit rate of the fit kernel at 1 and 2 blocks per SM (config 2, all 2000 fits)
mkdir -p gpurun_out
for b in 2 1; do
  echo "== blocks per SM $b"
  PAREBEN_BLOCKS_PER_SM=$b timeout 300 python scripts/profile_case.py binomial 2000 2>&1 | tail -1
  PAREBEN_BLOCKS_PER_SM=$b timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:eben_fit -s 1 -c 1 python scripts/profile_case.py binomial 2000 2>&1 | grep -E "dram__|lts__|l1tex__|gpu__time"
done
