#!/bin/bash
# Does host CPU contention inflate the per-kernel CUDA-event times?  (N=8 showed +7 % per segment on rank 0.)
run() { EAR_BENCH_RAYS=2e7 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']
print('$1 : %.4g seg/s  ms %.0f  closest %.0f anyhit %.0f shade %.0f splat %.0f' % (d['value'], d['ms_per_step'], k['closest'], k['anyhit'], k['shade'], k['splat']))"; }
run quiet
pids=""
for i in $(seq 1 $(nproc)); do ( while :; do :; done ) & pids="$pids $!"; done
run "$(nproc) spinners"
kill $pids
