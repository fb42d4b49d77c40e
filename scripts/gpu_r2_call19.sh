#!/bin/bash
# 1 GPU: clear kernel instead of the per-iteration memsets; list positions and queue fetch with one atomic per block; parity
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_cli_gpu.py tests/test_gpu_fullsize.py -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c19_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c19_pytest.log | tail -8
EAR_BENCH_VERBOSE=2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rays 4e7 2>gpurun_out/c19_err.log | python scripts/benchline.py
grep -E "render #[23]|e2e step|device-timed" gpurun_out/c19_err.log | sed 's/segments\/s.*kernels/kernels/' | tail -6
echo -n "c5 3e6: "; timeout 400 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6 2>>gpurun_out/c19_err.log | python scripts/benchline.py
echo -n "c5 3e6 sorted: "; EAR_B200_SORT_QUERIES=1 timeout 400 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6 2>>gpurun_out/c19_err.log | python scripts/benchline.py
