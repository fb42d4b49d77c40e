#!/bin/bash
# round 2, GPU call 3: whole GPU suite (mesh sources, group render, drop-in), scene_create breakdown, bench with e2e
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c3_pytest.log | tail -20
EAR_B200_DEBUG=1 timeout 300 python - > gpurun_out/c3_build_debug.log 2>&1 <<'PY'
import time, numpy as np
from ear_b200 import api, scenes
sc, table = scenes.synthetic_hall(n_tris=1_000_000, n_obstacles=2000, n_bands=8)
v, m = sc.triangles(), sc.triangle_materials()
for k in range(5):
    t0 = time.perf_counter(); s = api.Scene(v, m, table); t1 = time.perf_counter(); s.close(); t2 = time.perf_counter()
    print(f"scene_create run {k}: {1e3 * (t1 - t0):.1f} ms, destroy {1e3 * (t2 - t1):.1f} ms", flush=True)
PY
grep -E "scene_create|device bvh" gpurun_out/c3_build_debug.log | tail -40
EAR_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench_err.log
python scripts/benchline.py < gpurun_out/c3_bench.json
grep "e2e step" gpurun_out/c3_bench_err.log
