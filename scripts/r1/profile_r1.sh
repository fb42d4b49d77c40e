#!/bin/bash
# Round-1 profiling pass (run under gpurun, one GPU).  Numbers printed by runs under ncu are never bench values.
set -u
mkdir -p gpurun_out
export EAR_BENCH_RAYS=2e7
ARGS="bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
# 1. steady-state traversal launches (one closest-hit + one any-hit), full sections
ncu --set full --clock-control none --import-source on -k regex:wf_traverse_kernel -s 40 -c 2 -f -o gpurun_out/r1_traverse python $ARGS > gpurun_out/ncu_traverse.log 2>&1
# 2. shade + visibility-map lookup + splat
ncu --set full --clock-control none --import-source on -k regex:"wf_shade_kernel|wf_vismap_kernel|wf_splat_kernel" -s 60 -c 3 -f -o gpurun_out/r1_shade python $ARGS > gpurun_out/ncu_shade.log 2>&1
# 3. launch list of a whole (small) step: kernel shares
EAR_BENCH_RAYS=4e6 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1_launches.csv python $ARGS > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out | tail -12
