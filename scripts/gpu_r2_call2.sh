#!/bin/bash
# round 2, GPU call 2: device BVH builder (memcheck + tests + timing vs host builder), rank-returning sort, C5 diagnosis
mkdir -p gpurun_out
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device_bvh.py -q -x -k "tiny or edge" ) > gpurun_out/c2_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/c2_memcheck.log
tail -15 gpurun_out/c2_memcheck.log
( time timeout 900 python -m pytest tests/test_gpu_device_bvh.py tests/test_dropin_gpu.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_cli_gpu.py -q --deselect tests/test_gpu_round2.py::test_default_culling_equals_exact_on_1e8_adversarial_rays ) > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -6 gpurun_out/c2_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rays 4e7"
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>>gpurun_out/c2_err.log | python scripts/benchline.py; }
{
run device-build X=1
run host-build EAR_B200_BUILD=host
run device-build-again X=1
} > gpurun_out/c2_ab.log 2>&1
cat gpurun_out/c2_ab.log
EAR_B200_DEBUG=1 timeout 300 python - > gpurun_out/c2_build_debug.log 2>&1 <<'PY'
import time, numpy as np
from ear_b200 import api, scenes
for n in (1_000_000, 10_000_000):
    sc, table = (scenes.synthetic_hall(n_tris=n, n_obstacles=n // 500, n_bands=3) if n == 1_000_000 else scenes.synthetic_complex(n_tris=n, n_obstacles=n // 500, n_bands=3))
    v, m = sc.triangles(), sc.triangle_materials()
    for k in range(3):
        t0 = time.perf_counter(); s = api.Scene(v, m, table); t1 = time.perf_counter(); s.close()
        print(f"scene_create {n} tris run {k}: {1e3 * (t1 - t0):.1f} ms", flush=True)
PY
grep scene_create gpurun_out/c2_build_debug.log
C5="python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rays 3e6"
EAR_B200_DEBUG=1 timeout 600 $C5 > gpurun_out/c2_c5.json 2> gpurun_out/c2_c5_debug.log
python scripts/benchline.py < gpurun_out/c2_c5.json
grep -E "it [0-9]+:|vismap:.*entries|pool \+ vis" gpurun_out/c2_c5_debug.log | tail -40
