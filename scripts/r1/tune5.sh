for res in 0 512 1024 2048; do
  echo -n "VISMAP_RES=$res : "
  EAR_B200_VISMAP_RES=$res EAR_BENCH_RAYS=4e7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); k=d['kernel_ms_per_step']; print('%.3e seg/s  closest %.0f anyhit %.0f shade %.0f splat %.0f'%(d['value'], k['closest'],k['anyhit'],k['shade'],k['splat']))"
done
EAR_B200_VISMAP_RES=2048 EAR_B200_DEBUG=1 EAR_BENCH_RAYS=2e7 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e 2>&1 | grep "ear_b200" | sed -n "10,12p"
