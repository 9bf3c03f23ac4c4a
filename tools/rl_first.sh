#!/bin/bash
# first contact of the row-per-lane backward: correctness stages, then timing with / without fine checkpoints
mkdir -p gpurun_out
{
for st in tma_bwd rl rl_z_16bit long; do
  timeout 300 python tools/gpu_diag.py $st 2>&1 | tail -5
done
for shape in "12 128 65536" "12 1024 4096" "12 256 16384" "2 768 4096" "1 128 262144"; do
  echo "== $shape fine=1"; NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -2
  echo "== $shape fine=0"; NZ_PROF_FINE=0 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/rl_first.log
