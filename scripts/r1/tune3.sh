for ml in 4 8; do for ct in 1 2 3; do
  echo -n "MAX_LEAF=$ml SAH_CT=$ct : "
  EAR_B200_MAX_LEAF=$ml EAR_B200_SAH_CT=$ct EAR_BENCH_RAYS=4e7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.3e seg/s  closest %.0f anyhit %.0f shade %.0f splat %.0f'%(d['value'], k['closest'],k['anyhit'],k['shade'],k['splat']))"
done; done
