#!/bin/bash
# 1 GPU: warp-cooperative rejection sampling in the shade kernel (A/B against the per-lane loop), parity
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c14_pytest.log | tail -8
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --rays 4e7"
echo -n "coop: "; timeout 300 $B 2>>gpurun_out/c14_err.log | python scripts/benchline.py
echo -n "per-lane: "; EAR_B200_LIB=build_variants/nocoop.so timeout 300 $B 2>>gpurun_out/c14_err.log | python scripts/benchline.py
echo -n "c5 coop: "; timeout 400 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6 2>>gpurun_out/c14_err.log | python scripts/benchline.py
