#!/bin/bash
# Roofline evidence for bench.py (profiles/r2_traffic.json): ONE mid-step launch of the dominant kernel captured with ncu
# on the bench configuration itself, per GPU count.  A rank of an N-GPU run traces 1e8/N rays of the same scene, so its
# launches are those of a one-GPU run with --rays 1e8/N: that is what is captured here for N = 2, 4, 8 (ncu must not wrap
# a multi-rank command).  Caches and clocks are left alone (--cache-control none --clock-control none) so the captured
# launch runs as it does inside the step.  Output: gpurun_out/traffic_<key>.csv, turned into JSON by
# profiles/make_traffic_json.py.
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
cap() {   # key workload rays kernel-regex skip
  timeout 500 ncu --metrics $M --cache-control none --clock-control none --kernel-name-base demangled -k "regex:$4" -s $5 -c $6 --csv --log-file gpurun_out/traffic_$1.csv \
    python bench.py --workload $2 --rays $3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/traffic_$1.log 2>&1
  echo "$1 rc=$?"
}
# key workload rays kernel(s) launches-to-skip launches-to-capture; the demangled name selects the closest-hit instance
# <0, 0> (the any-hit instance <1, 0> of the same template runs every iteration too, with an empty list at C4)
if [ "${1:-all}" = all ]; then
cap c4_n1 c4 1e8 "wf_traverse_kernel<.bool.0" 100 1
cap c4_n2 c4 5e7 "wf_traverse_kernel<.bool.0" 100 1
cap c4_n4 c4 2.5e7 "wf_traverse_kernel<.bool.0" 100 1
cap c4_n8 c4 1.25e7 "wf_traverse_kernel<.bool.0" 30 1
fi
cap c5_n1 c5 1e7 "wf_vismap_kernel" 40 1
