#!/bin/bash
# occlusion queries: counting sort by (recorder, cell) vs slot order
for k in 1 0; do
  EAR_B200_SORT_QUERIES=$k EAR_BENCH_RAYS=4e7 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']
print('SORT_QUERIES=$k : %.4g seg/s  ms %.0f  closest %.0f anyhit %.0f shade %.0f splat %.0f' % (d['value'], d['ms_per_step'], k['closest'], k['anyhit'], k['shade'], k['splat']))"
done
