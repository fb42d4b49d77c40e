#!/bin/bash
# 2 GPUs: the multi-GPU tests (NCCL path, group render / peer reduce, CLI on 1 vs N GPUs) and a 2-rank bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c7_gpus.txt
( time timeout 900 python -m pytest tests/test_multi_gpu.py -q -v ) > gpurun_out/c7_multi_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c7_multi_gpu_tests.log
grep -E "PASSED|FAILED|SKIPPED|passed|failed|rc=" gpurun_out/c7_multi_gpu_tests.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c7_bench_n2.json 2> gpurun_out/c7_bench_n2.err
python scripts/benchline.py < gpurun_out/c7_bench_n2.json
tail -3 gpurun_out/c7_bench_n2.err
