#!/bin/bash
# pool generations at the per-rank work of the 8-GPU run (1.25e7 work items) and of the 1-GPU run
for r in 1.25e7; do for g in 4 2 1; do
  echo -n "RAYS=$r GENERATIONS=$g : "
  EAR_B200_GENERATIONS=$g EAR_BENCH_RAYS=$r timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.4g seg/s  ms %.0f closest %.1f anyhit %.1f shade %.1f splat %.1f launches %d'%(d['value'], d['ms_per_step'], k['closest'],k['anyhit'],k['shade'],k['splat'], d['launches_per_step']['closest']))"
done; done
