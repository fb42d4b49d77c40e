#!/bin/bash
# diagnosis at full size (1e8 rays): small device blocks outside the cache vs inside
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 1e8"
run() { echo "== $1"; shift; env EAR_BENCH_VERBOSE=2 "$@" timeout 400 $B 2>&1 >/dev/null | grep -E "render #[123]|device-timed|e2e step" | sed 's/segments\/s.*kernels/kernels/' | tail -6; }
run "blocks under 1 MB not cached (new default)" X=1
run "every block cached (as before)" EAR_B200_CACHE_MIN_KB=0
