for res in 1024 2048; do for cap in 16 32 48 96; do
  echo -n "VISMAP_RES=$res CAP=$cap : "
  EAR_B200_VISMAP_RES=$res EAR_B200_VISMAP_CAP=$cap EAR_BENCH_RAYS=4e7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.3e seg/s  closest %.0f anyhit %.0f shade %.0f splat %.0f'%(d['value'], k['closest'],k['anyhit'],k['shade'],k['splat']))"
done; done
