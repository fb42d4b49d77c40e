#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --rays 4e7"
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>>gpurun_out/c6_err.log | python scripts/benchline.py; }
{
run default X=1
run mb9 EAR_B200_LIB=build_variants/mb9.so
run mb10s20 EAR_B200_LIB=build_variants/mb10s20.so
run mb12s16 EAR_B200_LIB=build_variants/mb12s16.so
run shade5 EAR_B200_LIB=build_variants/shade5.so
run mb10s20shade5 EAR_B200_LIB=build_variants/mb10s20shade5.so
run sort-at-build EAR_B200_VISMAP_SORT=1
run sort-never EAR_B200_VISMAP_SORT=0
run default-again X=1
} > gpurun_out/c6_ab.log 2>&1
cat gpurun_out/c6_ab.log
( timeout 600 python -m pytest tests/test_convolve_gpu.py -q ) > gpurun_out/c6_conv.log 2>&1; tail -3 gpurun_out/c6_conv.log
EAR_CONVOLUTION=fft timeout 600 python scripts/c3_render.py > gpurun_out/c6_c3_fft.log 2>&1; tail -2 gpurun_out/c6_c3_fft.log
EAR_B200_DEBUG=1 EAR_B200_VISMAP_SORT=1 EAR_BENCH_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 2e7 2>&1 >/dev/null | grep -E "ordered by distance|e2e step|pool \+ vis" | tail -8
bash scripts/capture_traffic.sh c5only > gpurun_out/c6_traffic.log 2>&1; tail -2 gpurun_out/c6_traffic.log
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --cache-control none --clock-control none --kernel-name-base demangled -k "regex:wf_traverse_kernel<0" -s 100 -c 1 --csv --log-file gpurun_out/traffic_c4_n4.csv python bench.py --rays 2.5e7 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/traffic_c4_n4.log 2>&1; echo "c4_n4 rc=$?"
( time EAR_RUN_SLOW=1 timeout 1200 python -m pytest tests/test_gpu_statistics.py -q -k sweep -s ) > gpurun_out/c6_sweep.log 2>&1
tail -4 gpurun_out/c6_sweep.log
