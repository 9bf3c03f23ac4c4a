#!/bin/bash
# timing of the row-per-lane backward on the bench's dominant shapes (+ correctness of one stage) and one ncu capture
mkdir -p gpurun_out
{
timeout 300 python tools/gpu_diag.py rl 2>&1 | tail -3
for shape in "12 128 65536" "12 1024 4096" "12 256 16384" "2 768 4096"; do
  echo "== $shape"; NZ_PROF_FINE=1 timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/rl_time.log
if [ -n "$NCU_K" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -c ${NCU_C:-2} -o gpurun_out/${NCU_O:-rl_ncu} -f python tools/prof_scan.py ${NCU_SHAPE:-12 128 65536} 1 > gpurun_out/ncu_last.log 2>&1
  tail -2 gpurun_out/ncu_last.log
fi
