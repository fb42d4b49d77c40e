#!/bin/bash
# 1 GPU: shade-kernel counters through block sums and partial rows; render path vs trace_device path
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_cli_gpu.py -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c18_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c18_pytest.log | tail -8
EAR_BENCH_VERBOSE=2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rays 1e8 2>gpurun_out/c18_err.log | python scripts/benchline.py
grep -E "render #|e2e step|device-timed" gpurun_out/c18_err.log | tail -8
echo -n "c5 3e6: "; timeout 400 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6 2>>gpurun_out/c18_err.log | python scripts/benchline.py
