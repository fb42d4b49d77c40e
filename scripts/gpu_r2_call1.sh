#!/bin/bash
# round 2, GPU call 1: the GPU test suite, closest-hit A/B builds at C4, splat forms, first C5 numbers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
nproc >> gpurun_out/c1_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --rays 4e7"
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>gpurun_out/c1_err.log | python scripts/benchline.py; }
{
run default X=1
run default-again X=1
run noprmt EAR_B200_LIB=build_variants/noprmt.so
run nol2 EAR_B200_LIB=build_variants/nol2.so
run r1like EAR_B200_LIB=build_variants/r1like.so
run mb6 EAR_B200_LIB=build_variants/mb6.so
run mb7 EAR_B200_LIB=build_variants/mb7.so
run window EAR_B200_SPLAT=window
run fetch4 EAR_B200_FETCH_VOTE=4
run fetch12 EAR_B200_FETCH_VOTE=12
run leaf8 EAR_B200_LEAF_VOTE=8
run leaf16 EAR_B200_LEAF_VOTE=16
} > gpurun_out/c1_ab.log 2>&1
cat gpurun_out/c1_ab.log
C5="python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 6e6"
{
echo -n "c5 direct: "; timeout 600 $C5 2>gpurun_out/c1_c5_err.log | python scripts/benchline.py
echo -n "c5 window: "; EAR_B200_SPLAT=window timeout 600 $C5 2>>gpurun_out/c1_c5_err.log | python scripts/benchline.py
} > gpurun_out/c1_c5.log 2>&1
cat gpurun_out/c1_c5.log
( time timeout 600 python -m pytest tests/test_gpu_round2.py -q -k test_default_culling ) > gpurun_out/c1_exact.log 2>&1
tail -3 gpurun_out/c1_exact.log
