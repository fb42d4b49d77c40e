#!/bin/bash
# 1 GPU: resident blocks of the lookup and splat kernels (register cap), one-wave grids
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --rays 4e7"
C="python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6"
run() { echo -n "$1: "; shift; env "$@" timeout 400 $CMD 2>>gpurun_out/c15_err.log | python scripts/benchline.py; }
{
CMD=$B
run "c4 wave mb5" X=1
run "c4 fixed8 mb5" EAR_B200_GRID_WAVE=0
run "c4 wave mb6" EAR_B200_LIB=build_variants/mb6.so
run "c4 wave mb8" EAR_B200_LIB=build_variants/mb8.so
CMD=$C
run "c5 wave mb5" X=1
run "c5 fixed8 mb5" EAR_B200_GRID_WAVE=0
run "c5 wave mb6" EAR_B200_LIB=build_variants/mb6.so
run "c5 wave mb8" EAR_B200_LIB=build_variants/mb8.so
} > gpurun_out/c15_ab.log 2>&1
cat gpurun_out/c15_ab.log
