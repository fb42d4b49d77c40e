#!/bin/bash
# 1 GPU: map build without the counting REDs (offsets read off the sorted keys), zero-copy track download; parity + C4/C5 lines
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_post_gpu.py tests/test_gpu_fullsize.py -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c12_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c12_pytest.log | tail -8
echo -n "c4 4e7: "; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rays 4e7 2>>gpurun_out/c12_err.log | tee gpurun_out/c12_c4.json | python scripts/benchline.py
EAR_B200_DEBUG=1 EAR_BENCH_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays 4e7 2>&1 >/dev/null | grep -E "sort build|e2e step|pool \+ vis|scene_create|api\] render|render_sharded|trace_device" | tail -30
echo -n "c5 3e6: "; timeout 400 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --rays 3e6 2>>gpurun_out/c12_err.log | tee gpurun_out/c12_c5.json | python scripts/benchline.py
