#!/bin/bash
# ncu --set full captures of the kernels of one iteration (C4, 2e7 rays) and of the C5 occlusion kernels; launch list
mkdir -p gpurun_out
ARGS="bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
for k in wf_traverse_kernel wf_shade_kernel wf_vismap_kernel wf_splat_kernel wf_scatter_kernel; do
  skip=20; [ $k = wf_traverse_kernel ] && skip=40
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r2_$k python $ARGS --rays 2e7 > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
EAR_B200_SPLAT=window timeout 200 ncu --set full --clock-control none --import-source on -k regex:wf_splat_window_kernel -s 20 -c 1 -f -o gpurun_out/r2_wf_splat_window_kernel python $ARGS --rays 2e7 > gpurun_out/ncu_window.log 2>&1
echo "window rc=$?"
for k in wf_vismap_kernel wf_traverse_kernel wf_shade_kernel wf_splat_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 21 -c 1 -f -o gpurun_out/r2_c5_$k python $ARGS --workload c5 --rays 3e6 > gpurun_out/ncu_c5_$k.log 2>&1
  echo "c5 $k rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_c4.csv python $ARGS --rays 4e6 > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file gpurun_out/r2_launches_c5.csv python $ARGS --workload c5 --rays 1e6 > gpurun_out/ncu_list_c5.log 2>&1
echo "c5 launch list rc=$?"
ls -la gpurun_out | grep -E "r2_|traffic"
