#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_statistics.py -q -k "mesh or example2 or coarse or group" -s ) > gpurun_out/c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|band . track|coarse bins" gpurun_out/c5_pytest.log | tail -20
( time EAR_RUN_SLOW=1 timeout 1200 python -m pytest tests/test_gpu_statistics.py -q -k sweep -s ) > gpurun_out/c5_sweep.log 2>&1
tail -5 gpurun_out/c5_sweep.log
timeout 600 python scripts/c3_render.py > gpurun_out/c5_c3_render.log 2>&1; tail -4 gpurun_out/c5_c3_render.log
bash scripts/capture_traffic.sh
bash scripts/profile_r2.sh
