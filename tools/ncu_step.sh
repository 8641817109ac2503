#!/bin/bash
# ncu launch list of ONE eager training step (after 2 warm-up steps) + one full capture of the top kernel.
# usage (on the GPU box, from the repo root):  bash tools/ncu_step.sh
set -u
mkdir -p gpurun_out
CMD="python tools/profile_step.py --eager --no-profiler --warmup 2 --steps 1"
$CMD > gpurun_out/plain_step.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_step.log; exit 1; }
# the script brackets the measured step with cudaProfilerStart/Stop: only that step's ~5.3k launches are listed
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/ncu_gemm.py 1 > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/gemm_top python tools/ncu_gemm.py 1 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
python tools/ncu_gemm.py 1 proj > gpurun_out/ncu_plain_proj.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/gemm_proj python tools/ncu_gemm.py 1 proj > gpurun_out/ncu_full_proj.log 2>&1
echo "projection capture rc=$?"
