#!/bin/bash
# diagnosis: first render on a fresh scene vs later renders on the same scene (per kernel class)
mkdir -p gpurun_out
EAR_BENCH_VERBOSE=2 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 1e8 2>&1 >/dev/null | grep -E "render #|e2e step|device-timed" | tail -12
