#!/bin/bash
# leaf / fetch votes of the traversal loop, re-checked after the stack and pop changes
for lv in 8 12 16 20; do for fv in 6 8 12; do
  echo -n "LV=$lv FV=$fv : "
  EAR_B200_LEAF_VOTE=$lv EAR_B200_FETCH_VOTE=$fv EAR_BENCH_RAYS=2e7 timeout 120 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.4g seg/s  closest %.1f anyhit %.1f shade %.1f splat %.1f'%(d['value'], k['closest'],k['anyhit'],k['shade'],k['splat']))"
done; done
