#!/bin/bash
# A/B of the backward main pass: slab (NZ_RL_BWD2=0) vs butterfly (default), plus register-cap variants of the butterfly
mkdir -p gpurun_out
{
for shape in "12 128 65536" "12 128 262144" "12 256 65536" "12 1024 4096"; do
  echo "== $shape"
  echo "slab      $(NZ_RL_BWD2=0 NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)"
  echo "bfly12    $(NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)"
  for v in b2_minb8 b2_minb14 b2_minb16; do
    echo "$v $(NNUZOO_B200_LIB=tune_variants/$v/libnnuzoo_b200.so NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)"
  done
done
} 2>&1 | tee gpurun_out/ab_bwd2.log
