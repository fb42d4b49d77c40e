#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_device_bvh.py -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c9_pytest.log | tail -8
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rays 4e7"
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>>gpurun_out/c9_err.log | python scripts/benchline.py; }
{
run sort-build X=1
run atomic-build EAR_B200_VISMAP_BUILD=atomic
} > gpurun_out/c9_ab.log 2>&1
cat gpurun_out/c9_ab.log
EAR_B200_DEBUG=1 EAR_BENCH_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 4e7 2>&1 >/dev/null | grep -E "sort build|e2e step|pool \+ vis|trace_device|wavefront loop|render_sharded|api\] render|scene_create" | tail -30
echo -n "c5: "; timeout 600 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6 2>gpurun_out/c9_c5_err.log | python scripts/benchline.py
