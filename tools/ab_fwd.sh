#!/bin/bash
# A/B of the row-per-lane forward main pass at different register caps (NZ_RL_FWD=1 forces the row-per-lane forward)
mkdir -p gpurun_out
{
for shape in "12 128 65536" "12 128 262144" "12 256 65536" "12 1024 4096"; do
  echo "== $shape"
  echo "warpscan  $(NZ_RL_FWD=0 NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)"
  echo "rl16      $(NZ_RL_FWD=1 NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)"
  for v in "$@"; do
    echo "$v $(NZ_RL_FWD=1 NNUZOO_B200_LIB=tune_variants/$v/libnnuzoo_b200.so NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)"
  done
done
} 2>&1 | tee gpurun_out/ab_fwd.log
