"""C5-shaped run: 1e7 triangles, 64 mono recorders, 3 bands, 50 bounces, a reduced ray budget.
Not a bench line -- it checks that the path holds up at that size (BVH build, map memory budget, 64-recorder queries)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ear_b200 import api, scenes  # noqa: E402

n_tris = int(float(os.environ.get("C5_TRIS", "1e7")))
rays = int(float(os.environ.get("C5_RAYS", "3e5")))
t0 = time.perf_counter()
sc, table = scenes.synthetic_hall(n_tris=n_tris, n_obstacles=20000, n_bands=3, n_recorders=64, samples=rays * 10)
print(f"scene: {sc.triangles().shape[0]} triangles, {len(sc.recorders)} recorders, generated in {time.perf_counter() - t0:.1f} s", flush=True)
t0 = time.perf_counter()
scene = api.Scene(sc.triangles(), sc.triangle_materials(), table)
print(f"scene_create: {time.perf_counter() - t0:.2f} s", flush=True)
ctxs, recs = api.contexts_from_def(sc, n_bands=3)
for k in range(2):
    t0 = time.perf_counter()
    res = scene.render(ctxs, recs, max_bounces=50, seed=7)
    dt = time.perf_counter() - t0
    print(f"render #{k + 1}: {dt:.2f} s wall, {res.device_ms:.0f} ms device, {res.segments} segments "
          f"({res.segments / (res.device_ms * 1e-3):.3g}/s), {res.occlusion_queries} occlusion queries "
          f"({res.occlusion_queries / (res.device_ms * 1e-3):.3g}/s), {res.contributions} contributions, "
          f"{res.bin_updates} bin updates, dropped {res.dropped_updates}", flush=True)
    if k == 0:
        first = (res.segments, res.occlusion_queries, res.contributions, res.bin_updates)
    else:
        assert first == (res.segments, res.occlusion_queries, res.contributions, res.bin_updates), "not reproducible"
print("ok")
