#!/bin/bash
# ncu --set full captures of the kernels as they are at the end of round 2 (C4 at 2e7 rays; C5 lookups, shade and splat
# at 3e6 rays), and the launch lists of one small step of each workload
mkdir -p gpurun_out
ARGS="bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
for k in wf_traverse_kernel wf_shade_kernel wf_vismap_kernel wf_splat_kernel; do
  skip=20; [ $k = wf_traverse_kernel ] && skip=40
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r2f_$k python $ARGS --rays 2e7 > gpurun_out/ncu_f_$k.log 2>&1
  echo "$k rc=$?"
done
for k in wf_vismap_kernel wf_shade_kernel wf_splat_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 21 -c 1 -f -o gpurun_out/r2f_c5_$k python $ARGS --workload c5 --rays 3e6 > gpurun_out/ncu_f_c5_$k.log 2>&1
  echo "c5 $k rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2f_launches_c4.csv python $ARGS --rays 4e6 > gpurun_out/ncu_f_list.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file gpurun_out/r2f_launches_c5.csv python $ARGS --workload c5 --rays 1e6 > gpurun_out/ncu_f_list_c5.log 2>&1
echo "c5 launch list rc=$?"
ls -la gpurun_out | grep -E "r2f_|traffic" | awk '{print $5, $9}'
