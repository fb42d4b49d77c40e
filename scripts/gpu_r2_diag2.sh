#!/bin/bash
# diagnosis: why is the shade class 23 % slower through ear_b200_render than through ear_b200_trace_device on torch's stream?
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 4e7"
run() { echo "== $1"; shift; env EAR_BENCH_VERBOSE=2 "$@" timeout 300 $B 2>&1 >/dev/null | grep -E "render #[23]|device-timed" | sed 's/segments\/s.*kernels/kernels/' | tail -3; }
run "default" X=1
run "device leg on a new torch stream" EAR_BENCH_STREAM=new
run "library on the null stream" EAR_B200_STREAM=null
run "library on a blocking stream" EAR_B200_STREAM=blocking
