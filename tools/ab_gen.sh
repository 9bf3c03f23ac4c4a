#!/bin/bash
# generic A/B: bash tools/ab_gen.sh "ENV=.. ENV=.." variant...   over the main shapes; prints fwd / bwd times
envs="$1"; shift
mkdir -p gpurun_out
{
for shape in "12 128 65536" "12 128 262144" "12 256 65536" "12 1024 4096"; do
  echo "== $shape"
  echo "default   $(env $envs NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)"
  for v in "$@"; do
    echo "$v $(env $envs NNUZOO_B200_LIB=tune_variants/$v/libnnuzoo_b200.so NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)"
  done
done
} 2>&1 | tee -a gpurun_out/ab_gen.log
