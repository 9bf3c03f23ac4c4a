#!/bin/bash
# old (warp-scan, chained) vs row-per-lane (forced for every size) kernels over the M2Net scan shapes at batch 12
mkdir -p gpurun_out
{
for shape in "12 128 262144" "12 128 65536" "12 128 16384" "12 128 4096" "12 128 1024" "12 256 65536" "12 256 16384" "12 256 4096" "12 256 1024" "12 512 16384" "12 512 4096" "12 512 1024" "12 512 256" "12 1024 4096" "12 1024 1024" "12 1024 256" "2 768 4096" "1 128 262144" "2 64 2097152"; do
  n=$(NZ_RL_MIN_ELTS=0 NZ_RL_FWD=1 NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)
  o=$(NZ_NO_RL=1 NZ_PROF_FINE=0 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)
  echo "RL  $n"; echo "OLD $o"
done
} 2>&1 | tee gpurun_out/rl_table.log
