#!/bin/bash
# Round-end evidence on one B200: GPU test log, bench line, ncu launch list of one bench step (times + DRAM bytes),
# one `--set full` capture of the stage-1 scan (all kernels of a fwd + bwd), reference arm line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -2 gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 200 gpurun_out/bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; tail -c 300 gpurun_out/bench_reference_arm.json
NZ_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-train --no-infer --no-configs > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_final.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scan|combine" -o gpurun_out/final_stage1 -f python tools/prof_scan.py 12 128 65536 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/final_stage1.ncu-rep
