for lv in 4 8 12; do for fv in 4 8 16; do
  echo -n "LV=$lv FV=$fv : "
  EAR_B200_SLOTS=4194304 EAR_B200_LEAF_VOTE=$lv EAR_B200_FETCH_VOTE=$fv EAR_BENCH_RAYS=2e7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.3e seg/s  closest %.0f anyhit %.0f shade %.0f splat %.0f'%(d['value'], k['closest'],k['anyhit'],k['shade'],k['splat']))"
done; done
