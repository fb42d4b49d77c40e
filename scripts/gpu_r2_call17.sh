#!/bin/bash
# 1 GPU: rank atomics of the query enqueue aggregated over lanes that share a cell; K6 atomics issued early; sorted vs unsorted queries
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/c17_pytest.log | tail -8
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --rays 4e7"
C="python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6"
echo -n "c4 sorted: "; timeout 300 $B 2>>gpurun_out/c17_err.log | python scripts/benchline.py
echo -n "c4 unsorted: "; EAR_B200_SORT_QUERIES=0 timeout 300 $B 2>>gpurun_out/c17_err.log | python scripts/benchline.py
echo -n "c5 sorted: "; timeout 400 $C 2>>gpurun_out/c17_err.log | python scripts/benchline.py
echo -n "c5 unsorted: "; EAR_B200_SORT_QUERIES=0 timeout 400 $C 2>>gpurun_out/c17_err.log | python scripts/benchline.py
