#!/bin/bash
# 1 GPU: branch-free child pick in the node step (closest-hit / any-hit kernels)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_device_bvh.py -q -x ) > gpurun_out/c13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c13_pytest.log | tail -8
echo -n "c4 4e7: "; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rays 4e7 2>>gpurun_out/c13_err.log | tee gpurun_out/c13_c4.json | python scripts/benchline.py
echo -n "c5 3e6: "; timeout 400 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6 2>>gpurun_out/c13_err.log | tee gpurun_out/c13_c5.json | python scripts/benchline.py
