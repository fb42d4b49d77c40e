#!/bin/bash
# 8 GPUs: the C4 headline line, the C5 line at full size, the in-process group render tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/c10_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
export EAR_BENCH_VERBOSE=1
timeout 420 $TR --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/c10_c4_n8.json 2> gpurun_out/c10_c4_n8.err
echo "c4 rc=$?"; python scripts/benchline.py < gpurun_out/c10_c4_n8.json
grep -E "device-timed|e2e step" gpurun_out/c10_c4_n8.err | grep "rank 0\|device-timed" | tail -8
timeout 900 $TR --master-port 29612 bench.py --gpus 8 --workload c5 --steps 1 --warmup 2 --warmup-rays 2.4e7 --e2e-steps 1 > gpurun_out/c10_c5_n8.json 2> gpurun_out/c10_c5_n8.err
echo "c5 rc=$?"; python scripts/benchline.py < gpurun_out/c10_c5_n8.json
grep -E "device-timed|e2e step" gpurun_out/c10_c5_n8.err | grep "rank 0\|device-timed" | tail -4
( timeout 600 python -m pytest tests/test_multi_gpu.py -q -x -k "group_render or calc_t60" ) > gpurun_out/c10_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/c10_pytest.log
