for mb in 4 5 6 8; do for lv in 8 12 16 24; do
  echo -n "MINB=$mb LV=$lv : "
  EAR_B200_MIN_BLOCKS=$mb EAR_B200_LEAF_VOTE=$lv EAR_BENCH_RAYS=8e6 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.3e seg/s  kernel %.1f ms'%(d['value'], d['roofline']['kernel_ms']))"
done; done
