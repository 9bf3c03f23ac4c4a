#!/bin/bash
# per-kernel durations (ncu, cold, serialised) of one fwd + bwd of a shape:  bash tools/rl_ktimes.sh "12 128 65536"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:"scan|combine" --csv --log-file gpurun_out/ktimes.csv python tools/prof_scan.py $1 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ktimes.csv')))
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[hi]; k=h.index('Kernel Name'); m=h.index('Metric Name'); v=h.index('Metric Value'); idc=h.index('ID')
out={}
for r in rows[hi+1:]:
    if len(r)<=v: continue
    out.setdefault((int(r[idc]),r[k][:60]),{})[r[m].split('.')[0][-22:]]=r[v]
for (i,n),d in sorted(out.items()):
    print(i,n,' '.join(f"{a}={b}" for a,b in d.items()))
PY
