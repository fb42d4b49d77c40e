#!/bin/bash
# One kernel per ncu run, each under its own timeout (a multi-kernel capture of this app once ran into gpurun's limit).
mkdir -p gpurun_out
export EAR_BENCH_RAYS=2e7
ARGS="bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
for k in wf_traverse_kernel wf_vismap_kernel wf_shade_kernel wf_splat_kernel; do
  skip=20; [ $k = wf_traverse_kernel ] && skip=40
  timeout 170 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r1_final_$k python $ARGS > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out | grep r1_final
