#!/bin/bash
# visibility-map list cap (longer lists -> fewer BVH fallbacks, longer worst-case lookups)
for c in 64 96 128 192 255; do
  EAR_B200_VISMAP_CAP=$c EAR_BENCH_RAYS=2e7 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']
print('CAP=$c : %.4g seg/s  ms %.0f  closest %.1f anyhit %.1f shade %.1f splat %.1f' % (d['value'], d['ms_per_step'], k['closest'], k['anyhit'], k['shade'], k['splat']))"
done
