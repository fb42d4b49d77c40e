#!/bin/bash
# diagnosis at full size (1e8 rays): does a stated shared-memory carve-out make the render path's shade class fast?
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 1e8"
run() { echo "== $1"; shift; env EAR_BENCH_VERBOSE=2 "$@" timeout 400 $B 2>&1 >/dev/null | grep -E "render #[123]|device-timed|e2e step" | sed 's/segments\/s.*kernels/kernels/' | tail -6; }
run "default" X=1
run "carve-out stated" EAR_B200_CARVEOUT=1
