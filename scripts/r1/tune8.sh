#!/bin/bash
# ray binning key: 0 octant+cell12 (15 bit), 1 octant+axis order+cell12 (18 bit), 2 octant+cell15 (18 bit)
for k in 2 3 4; do
  EAR_B200_RAY_KEY=$k EAR_BENCH_RAYS=4e7 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']
print('RAY_KEY=$k : %.4g seg/s  ms %.0f  closest %.0f anyhit %.0f shade %.0f splat %.0f' % (d['value'], d['ms_per_step'], k['closest'], k['anyhit'], k['shade'], k['splat']))"
done
