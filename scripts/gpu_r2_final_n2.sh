#!/bin/bash
# final 2-GPU line (C4) on the end-of-round code
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/f2_c4_n2.json 2> gpurun_out/f2_c4_n2.err
echo "c4 n2 rc=$?"; python scripts/benchline.py < gpurun_out/f2_c4_n2.json
