for cap in 96 192 384 100000; do
  echo -n "dir-sort RES=1024 CAP=$cap : "
  EAR_B200_VISMAP_RES=1024 EAR_B200_VISMAP_CAP=$cap EAR_BENCH_RAYS=4e7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.3e seg/s  closest %.0f anyhit %.0f shade %.0f splat %.0f'%(d['value'], k['closest'],k['anyhit'],k['shade'],k['splat']))"
done
echo -n "cell-sort RES=1024 CAP=96 : "
EAR_B200_Q_SORT_CELL=1 EAR_B200_VISMAP_RES=1024 EAR_B200_VISMAP_CAP=96 EAR_BENCH_RAYS=4e7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.3e seg/s  closest %.0f anyhit %.0f shade %.0f splat %.0f'%(d['value'], k['closest'],k['anyhit'],k['shade'],k['splat']))"
