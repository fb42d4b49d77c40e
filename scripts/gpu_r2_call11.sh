#!/bin/bash
# 1 GPU: over-long texel lists tested on a nearest-first prefix before the BVH (C5), count-pass lap of the sorted map build
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_post_gpu.py tests/test_cli_gpu.py tests/test_dropin_gpu.py -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c11_pytest.log | tail -8
B="python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6"
run() { echo -n "$1: "; shift; env "$@" timeout 400 $B 2>>gpurun_out/c11_err.log | python scripts/benchline.py; }
{
run "c5 prefix0" EAR_B200_VISMAP_PREFIX=0
run "c5 prefix8" EAR_B200_VISMAP_PREFIX=8
run "c5 prefix32" EAR_B200_VISMAP_PREFIX=32
run "c5 prefix128" EAR_B200_VISMAP_PREFIX=128
run "c5 prefix32 nosortq" EAR_B200_VISMAP_PREFIX=32 EAR_B200_SORT_QUERIES=0
} > gpurun_out/c11_ab.log 2>&1
cat gpurun_out/c11_ab.log
EAR_B200_DEBUG=1 EAR_BENCH_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 4e7 2>&1 >/dev/null | grep -E "sort build|e2e step|pool \+ vis|scene_create" | tail -24
