#!/bin/bash
# visibility-map resolution with the 256 cap
for r in 512 1024 2048; do
  EAR_B200_VISMAP_RES=$r EAR_BENCH_RAYS=2e7 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']
print('RES=$r : %.4g seg/s  ms %.0f  e2e ms %.0f  closest %.1f anyhit %.1f shade %.1f splat %.1f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], k['closest'], k['anyhit'], k['shade'], k['splat']))"
done
