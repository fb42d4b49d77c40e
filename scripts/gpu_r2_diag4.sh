#!/bin/bash
# diagnosis at full size (1e8 rays): shade class of the render path vs stream choice
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 1e8"
run() { echo "== $1"; shift; env EAR_BENCH_VERBOSE=2 "$@" timeout 400 $B 2>&1 >/dev/null | grep -E "render #[123]|device-timed|e2e step" | sed 's/segments\/s.*kernels/kernels/' | tail -6; }
run "library on the null stream" EAR_B200_STREAM=null
run "device leg on a new torch stream" EAR_BENCH_STREAM=new
run "cache off" EAR_B200_CACHE_MB=0
