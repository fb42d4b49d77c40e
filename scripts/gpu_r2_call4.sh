#!/bin/bash
# round 2, GPU call 4: failing tests re-run, statistics vs EAR_ref, e2e breakdown with the device-memory cache, C5 with sorted maps
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_dropin_gpu.py tests/test_gpu_round2.py tests/test_gpu_statistics.py tests/test_gpu_device_bvh.py -q --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays -s ) > gpurun_out/c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|T60 @|coarse bins" gpurun_out/c4_pytest.log | tail -20
EAR_BENCH_VERBOSE=1 EAR_B200_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench_err.log
python scripts/benchline.py < gpurun_out/c4_bench.json
grep -E "e2e step|render:|scene_create|pool \+ vis|vismap: (alloc|count|scan|fill)" gpurun_out/c4_bench_err.log | tail -40
C5="python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6"
{
echo -n "c5 sort-auto: "; timeout 600 $C5 2>gpurun_out/c4_c5_err.log | python scripts/benchline.py
echo -n "c5 sort-off: "; EAR_B200_VISMAP_SORT=0 timeout 600 $C5 2>>gpurun_out/c4_c5_err.log | python scripts/benchline.py
echo -n "c5 sort-auto cap1024: "; EAR_B200_VISMAP_CAP=1024 timeout 600 $C5 2>>gpurun_out/c4_c5_err.log | python scripts/benchline.py
} > gpurun_out/c4_c5.log 2>&1
cat gpurun_out/c4_c5.log
