#!/bin/bash
# final 4-GPU line (C4) on the end-of-round code
mkdir -p gpurun_out
timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/f4_c4_n4.json 2> gpurun_out/f4_c4_n4.err
echo "c4 n4 rc=$?"; python scripts/benchline.py < gpurun_out/f4_c4_n4.json
