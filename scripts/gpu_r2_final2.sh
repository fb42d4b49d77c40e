#!/bin/bash
# final 1-GPU call B: roofline traffic captures on the bench configurations, ncu --set full of the kernels, launch lists
mkdir -p gpurun_out
rm -f gpurun_out/traffic_*.csv gpurun_out/traffic_*.log
bash scripts/capture_traffic.sh all
bash scripts/profile_r2_final.sh
