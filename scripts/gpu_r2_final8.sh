#!/bin/bash
# final 8-GPU call: the C4 headline line and the C5 line at full size
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
export EAR_BENCH_VERBOSE=1
timeout 420 $TR --master-port 29621 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/f8_c4_n8.json 2> gpurun_out/f8_c4_n8.err
echo "c4 rc=$?"; python scripts/benchline.py < gpurun_out/f8_c4_n8.json
grep -E "device-timed|e2e step" gpurun_out/f8_c4_n8.err | grep "rank 0\|device-timed" | tail -6
timeout 600 $TR --master-port 29622 bench.py --gpus 8 --workload c5 --steps 1 --warmup 2 --warmup-rays 2.4e7 --e2e-steps 1 > gpurun_out/f8_c5_n8.json 2> gpurun_out/f8_c5_n8.err
echo "c5 rc=$?"; python scripts/benchline.py < gpurun_out/f8_c5_n8.json
grep -E "device-timed|e2e step" gpurun_out/f8_c5_n8.err | grep "rank 0\|device-timed" | tail -3
