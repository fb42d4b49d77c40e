#!/bin/bash
# diagnosis: per-class times of the render path at full size (1e8 rays)
mkdir -p gpurun_out
EAR_BENCH_VERBOSE=2 EAR_B200_DEBUG=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 1e8 2>&1 >/dev/null | grep -E "render #|e2e step|device-timed|wavefront loop|pool \+ vis|trace_device|api\] render|scene_create" | sed 's/segments\/s, //' | tail -30
