#!/bin/bash
# 1 GPU: leaf-vote / fetch-vote thresholds of the traversal kernel after the occupancy and node-step changes (4e7 rays)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --rays 4e7"
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>>gpurun_out/c20_err.log | python scripts/benchline.py; }
{
run "leaf12 fetch8 (default)" X=1
run "leaf8 fetch8" EAR_B200_LEAF_VOTE=8
run "leaf16 fetch8" EAR_B200_LEAF_VOTE=16
run "leaf20 fetch8" EAR_B200_LEAF_VOTE=20
run "leaf12 fetch4" EAR_B200_FETCH_VOTE=4
run "leaf12 fetch12" EAR_B200_FETCH_VOTE=12
run "leaf12 fetch16" EAR_B200_FETCH_VOTE=16
run "leaf16 fetch12" EAR_B200_LEAF_VOTE=16 EAR_B200_FETCH_VOTE=12
} > gpurun_out/c20_ab.log 2>&1
cat gpurun_out/c20_ab.log
