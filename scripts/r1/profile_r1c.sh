#!/bin/bash
# steady-state closest-hit launch with the default (16 Mi-slot) pool: DRAM traffic for bench.py's roofline.traffic
mkdir -p gpurun_out
EAR_BENCH_RAYS=4e7 timeout 250 ncu --set full --clock-control none --import-source on -k regex:wf_traverse_kernel -s 40 -c 1 -f -o gpurun_out/r1_traverse_16mi python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_traverse_16mi.log 2>&1
echo rc=$?; ls -la gpurun_out | grep 16mi
