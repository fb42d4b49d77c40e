#!/bin/bash
# final 1-GPU call (after the cache fix): parity subset, the C4 line at full size, the C5 line on one GPU's share
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -x -m gpu ) > gpurun_out/f3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f3_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=|real" gpurun_out/f3_pytest.log | tail -6
EAR_BENCH_VERBOSE=1 timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/f3_bench_c4_n1.json 2> gpurun_out/f3_bench_c4_n1.err; echo "c4 rc=$?"
python scripts/benchline.py < gpurun_out/f3_bench_c4_n1.json
EAR_BENCH_VERBOSE=1 timeout 900 python bench.py --workload c5 --rays 1.25e8 --steps 1 --warmup 2 --warmup-rays 3e6 --e2e-steps 1 > gpurun_out/f3_bench_c5_n1.json 2> gpurun_out/f3_bench_c5_n1.err; echo "c5 rc=$?"
python scripts/benchline.py < gpurun_out/f3_bench_c5_n1.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f3_smoke.log
