#!/bin/bash
# 1 GPU: two-pass query enqueue in the shade kernel (one counter atomic per warp, rank atomics four at a time); parity
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c16_pytest.log | tail -8
echo -n "c4 4e7: "; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rays 4e7 2>>gpurun_out/c16_err.log | python scripts/benchline.py
echo -n "c5 3e6: "; timeout 400 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6 2>>gpurun_out/c16_err.log | python scripts/benchline.py
echo -n "c5 3e6 nosortq: "; EAR_B200_SORT_QUERIES=0 timeout 400 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6 2>>gpurun_out/c16_err.log | python scripts/benchline.py
