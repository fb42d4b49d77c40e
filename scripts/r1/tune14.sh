#!/bin/bash
# pool size (rays in flight)
for sl in 16777216 33554432 67108864; do
  echo -n "SLOTS=$sl : "
  EAR_B200_SLOTS=$sl EAR_BENCH_RAYS=8e7 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.4g seg/s  ms %.0f closest %.1f anyhit %.1f shade %.1f splat %.1f'%(d['value'], d['ms_per_step'], k['closest'],k['anyhit'],k['shade'],k['splat']))"
done
