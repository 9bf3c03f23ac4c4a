#!/bin/bash
# time every built variant (tune_variants/*/libnnuzoo_b200.so) on a few shapes, after a correctness check
mkdir -p gpurun_out
SHAPES=${SHAPES:-"12 1024 4096;12 128 65536;12 256 16384"}
for d in tune_variants/*/; do
  v=$(basename $d)
  export NNUZOO_B200_LIB=$PWD/$d/libnnuzoo_b200.so
  ok=$(timeout 200 python tools/gpu_diag.py tma_bwd 2>&1 | python -c "
import sys,json
worst=0
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        for k,d in json.loads(l).items(): worst=max(worst,max(d.values()))
print('maxrel=%.2e'%worst)")
  IFS=';' read -ra SH <<< "$SHAPES"
  for shape in "${SH[@]}"; do
    r=$(timeout 200 python tools/prof_scan.py $shape 3 2>&1 | tail -1)
    echo "[$v] $ok | $r"
  done
done 2>&1 | tee gpurun_out/tune.log
