#!/bin/bash
# 2 GPUs: e2e breakdown per rank (library laps), any-hit occupancy variants at C5 on one GPU
mkdir -p gpurun_out
EAR_B200_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err
python scripts/benchline.py < gpurun_out/c8_bench_n2.json
python -c "
import json
for ln in open('gpurun_out/c8_bench_n2.json'):
    if ln.startswith('{'): d=json.loads(ln); print(d['e2e'])
"
grep -E "scene_create|device bvh|pool \+ vis|ordered by|render:" gpurun_out/c8_bench_n2.err | tail -60
C5="python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6"
for v in 6 8 9 10; do echo -n "c5 any$v: "; CUDA_VISIBLE_DEVICES=0 EAR_B200_LIB=build_variants/any$v.so timeout 600 $C5 2>>gpurun_out/c8_c5_err.log | python scripts/benchline.py; done
