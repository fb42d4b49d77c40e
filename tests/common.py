"""Shared scene / ray generators for the parity tests (seeded, deterministic)."""
from __future__ import annotations

import numpy as np

from ear_b200 import scenes
from ear_b200.earfile import MaterialDef, MeshDef, RecorderDef, SceneDef, SourceDef


def soup_scene(n_tris=3000, seed=5, extent=10.0, size=0.8) -> SceneDef:
    """Random triangle soup: unstructured, overlapping, many near-degenerate configurations."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n_tris, 1, 3))
    v = (c + rng.normal(scale=size, size=(n_tris, 3, 3))).astype(np.float32)
    sc = SceneDef(samples=10000)
    sc.materials.append(MaterialDef("a", (0.9, 0.8, 0.7), (0.0, 0.1, 0.2), (0.2, 0.5, 0.8)))
    sc.materials.append(MaterialDef("b", (0.5, 0.6, 0.7), (0.3, 0.2, 0.1), (0.0, 1.0, 0.4)))
    sc.meshes.append(MeshDef("a", v[: n_tris // 2]))
    sc.meshes.append(MeshDef("b", v[n_tris // 2:]))
    sc.sources.append(SourceDef(["/tmp/click.wav"], position=(0.5, 0.25, 0.1)))
    sc.recorders.append(RecorderDef("/tmp/o.wav", position=(-1.0, 2.0, 0.5)))
    return sc


def small_hall(n_tris=20000, n_obstacles=200, seed=0):
    sc, table = scenes.synthetic_hall(n_tris=n_tris, n_obstacles=n_obstacles, n_bands=3, seed=seed, samples=10000)
    return sc


def named_scene(name: str) -> SceneDef:
    if name == "rt60":
        return scenes.rt60_scene(samples=10000)
    if name == "rt60_saved":   # material as saved in RT60.blend: refl (.95,.05,.95) spec (0,1,0)
        return scenes.rt60_scene(refl=(0.95, 0.05, 0.95), spec=(0.0, 1.0, 0.0), samples=10000)
    if name == "example1":
        return scenes.example1_scene(samples=10000)
    if name == "hall20k":
        return small_hall()
    if name == "soup":
        return soup_scene()
    raise KeyError(name)


def make_rays(sc: SceneDef, n: int, seed: int = 1):
    """A mix of (a) uniform rays from inside the bounds, (b) rays aimed exactly at triangle
    vertices / edge points (shared edges, ties), (c) rays grazing a triangle's plane, (d) rays
    leaving a surface point (the bounce case: origin on a triangle)."""
    rng = np.random.default_rng(seed)
    tris = sc.triangles()
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    T = tris.shape[0]
    q = n // 4
    o = [rng.uniform(lo, hi, (q, 3))]
    d = [rng.normal(size=(q, 3))]
    # (b) aimed at vertices and edge points
    k = rng.integers(0, T, q)
    w = rng.integers(0, 3, q)
    a, b = tris[k, w], tris[k, (w + 1) % 3]
    s = rng.choice([0.0, 0.25, 0.5, 1.0], q)[:, None]
    target = a + s * (b - a)
    ob = rng.uniform(lo, hi, (q, 3))
    o.append(ob)
    d.append(target - ob)
    # (c) grazing: start just above the plane of triangle k, travel almost inside the plane
    k = rng.integers(0, T, q)
    v0, v1, v2 = tris[k, 0], tris[k, 1], tris[k, 2]
    nrm = np.cross(v1 - v0, v2 - v0)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    cen = (v0 + v1 + v2) / 3.0
    inplane = v1 - v0
    inplane /= np.maximum(np.linalg.norm(inplane, axis=1, keepdims=True), 1e-20)
    tilt = 10.0 ** rng.uniform(-6, -1, (q, 1)) * rng.choice([-1.0, 1.0], (q, 1))
    dist = rng.uniform(0.5, 5.0, (q, 1))
    og = cen - inplane * dist - nrm * tilt * dist
    o.append(og)
    d.append(inplane + nrm * tilt)
    # (d) origin on a surface, random outgoing direction
    k = rng.integers(0, T, n - 3 * q)
    u = rng.uniform(0, 1, (n - 3 * q, 2))
    flip = u.sum(1) > 1
    u[flip] = 1 - u[flip]
    p = tris[k, 0] + u[:, :1] * (tris[k, 1] - tris[k, 0]) + u[:, 1:] * (tris[k, 2] - tris[k, 0])
    o.append(p)
    d.append(rng.normal(size=(n - 3 * q, 3)))
    o = np.concatenate(o).astype(np.float32)
    d = np.concatenate(d)
    d = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)).astype(np.float32)
    return o, d


def make_segments(sc: SceneDef, n: int, seed: int = 2):
    """Occlusion queries: surface point -> recorder (the render case), plus random pairs."""
    rng = np.random.default_rng(seed)
    tris = sc.triangles()
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    T = tris.shape[0]
    h = n // 2
    k = rng.integers(0, T, h)
    u = rng.uniform(0, 1, (h, 2))
    flip = u.sum(1) > 1
    u[flip] = 1 - u[flip]
    p1 = tris[k, 0] + u[:, :1] * (tris[k, 1] - tris[k, 0]) + u[:, 1:] * (tris[k, 2] - tris[k, 0])
    x1 = np.tile(np.asarray(sc.recorders[0].position, np.float64)[None, :], (h, 1))
    p2 = rng.uniform(lo, hi, (n - h, 3))
    x2 = rng.uniform(lo, hi, (n - h, 3))
    return np.concatenate([p1, p2]).astype(np.float32), np.concatenate([x1, x2]).astype(np.float32)


def make_segments_to_point(sc: SceneDef, n: int, x, seed: int = 3):
    """Occlusion queries that all end at one point (what the render loop asks): surface points, points exactly
    on triangle vertices / edges, free-space points, and points whose segment grazes a triangle's plane."""
    rng = np.random.default_rng(seed)
    tris = sc.triangles()
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    T = tris.shape[0]
    x = np.asarray(x, np.float64)
    q = n // 4
    k = rng.integers(0, T, q)
    u = rng.uniform(0, 1, (q, 2))
    flip = u.sum(1) > 1
    u[flip] = 1 - u[flip]
    p1 = tris[k, 0] + u[:, :1] * (tris[k, 1] - tris[k, 0]) + u[:, 1:] * (tris[k, 2] - tris[k, 0])
    k = rng.integers(0, T, q)
    w = rng.integers(0, 3, q)
    s = rng.choice([0.0, 0.5, 1.0], q)[:, None]
    p2 = tris[k, w] + s * (tris[k, (w + 1) % 3] - tris[k, w])
    p3 = rng.uniform(lo, hi, (q, 3))
    # grazing: start in the plane of triangle k, on the far side of it as seen from x, so the segment skims the triangle
    k = rng.integers(0, T, n - 3 * q)
    cen = tris[k].mean(1)
    d = cen - x
    p4 = cen + d * rng.uniform(0.01, 0.5, (n - 3 * q, 1)) + rng.normal(scale=1e-4, size=(n - 3 * q, 3))
    p = np.concatenate([p1, p2, p3, p4]).astype(np.float32)
    return p, np.tile(x.astype(np.float32)[None, :], (n, 1))


def edge_case_scene():
    """example1 plus: every triangle stored twice more (equal t: the LOWER original index must win, src/Mesh.cpp:40
    strict '<'), zero-area triangles (two or three equal vertices, collinear), one huge far-away triangle."""
    sc = named_scene("example1")
    tris = sc.triangles()
    rng = np.random.default_rng(3)
    deg = tris[rng.integers(0, tris.shape[0], 12)].copy()
    deg[:4, 1] = deg[:4, 0]                                   # two equal vertices
    deg[4:8, 1] = deg[4:8, 0]
    deg[4:8, 2] = deg[4:8, 0]                                 # a point
    deg[8:, 2] = 0.5 * (deg[8:, 0] + deg[8:, 1])              # collinear
    far = np.array([[[9000.0, 9000.0, -5.0], [9500.0, 9000.0, 300.0], [9000.0, 9500.0, 300.0]]], np.float32)
    extra = np.concatenate([tris, deg, far, tris[::-1]]).astype(np.float32)
    sc.meshes.append(MeshDef(sc.meshes[0].material, extra))
    return sc, tris.shape[0]


def edge_case_queries(sc: SceneDef, n: int = 20000):
    """Rays with exactly zero direction components (infinite inverse directions in the slab test), rays starting
    exactly on vertices, zero-length occlusion segments -- on top of the usual mix."""
    rng = np.random.default_rng(4)
    o, d = make_rays(sc, n, seed=9)
    q = n // 4
    d[:q] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, q)] * rng.choice([-1.0, 1.0], (q, 1)).astype(np.float32)
    d[q: 2 * q, rng.integers(0, 3)] = 0.0
    verts = sc.triangles().reshape(-1, 3)
    o[2 * q: 2 * q + n // 10] = verts[rng.integers(0, verts.shape[0], n // 10)]
    p, x = make_segments(sc, n, seed=10)
    x[: n // 40] = p[: n // 40]
    return o, d, p, x


# ---------------------------------------------------------------------------------------------------------
# the brute-force oracle over many host threads (ctypes releases the GIL during the call): full-size scenes
# ---------------------------------------------------------------------------------------------------------
def oracle_first_hit_parallel(cpu, o, d, workers=None):
    import os
    from concurrent.futures import ThreadPoolExecutor
    workers = workers or max(1, min(32, os.cpu_count() or 1))
    n = o.shape[0]
    cuts = [n * k // workers for k in range(workers + 1)]
    with ThreadPoolExecutor(workers) as ex:
        parts = list(ex.map(lambda k: cpu.first_hit(o[cuts[k]:cuts[k + 1]], d[cuts[k]:cuts[k + 1]]), range(workers)))
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


def oracle_occluded_parallel(cpu, p, x, workers=None):
    import os
    from concurrent.futures import ThreadPoolExecutor
    workers = workers or max(1, min(32, os.cpu_count() or 1))
    n = p.shape[0]
    cuts = [n * k // workers for k in range(workers + 1)]
    with ThreadPoolExecutor(workers) as ex:
        parts = list(ex.map(lambda k: cpu.occluded(p[cuts[k]:cuts[k + 1]], x[cuts[k]:cuts[k + 1]]), range(workers)))
    return np.concatenate(parts)


def oracle_render_parallel(cpu, ctxs, recs, max_bounces, seed, workers=None):
    """Raw (not finalised) oracle render with the ray-id range of every context dealt to host threads; the partial
    tracks are summed in float64.  Returns (tracks[ctx][rec][k] = (data float64, first_sample, real_length), counters)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    workers = workers or max(1, min(32, os.cpu_count() or 1))
    rays = max(int(c.num_samples) for c in ctxs)
    workers = max(1, min(workers, rays))
    cuts = [rays * k // workers for k in range(workers + 1)]
    from oracle import binding as ob

    def part(k):
        return cpu.render(ctxs, recs, max_bounces=max_bounces, rng_mode=ob.RNG_PHILOX, seed=seed, first_ray=cuts[k],
                          ray_count=cuts[k + 1] - cuts[k], finalise=False)
    with ThreadPoolExecutor(workers) as ex:
        parts = list(ex.map(part, range(workers)))
    total, counters = None, {}
    for tracks, cnt in parts:
        for key, v in cnt.items():
            counters[key] = counters.get(key, 0) + v
        if total is None:
            total = [[[[t.data.astype(np.float64), t.first_sample, t.real_length] for t in pair] for pair in per_rec] for per_rec in tracks]
            continue
        for c, per_rec in enumerate(tracks):
            for r, pair in enumerate(per_rec):
                for k, t in enumerate(pair):
                    acc = total[c][r][k]
                    if t.data.shape[0] > acc[0].shape[0]:
                        acc[0] = np.concatenate([acc[0], np.zeros(t.data.shape[0] - acc[0].shape[0])])
                    acc[0][: t.data.shape[0]] += t.data
                    acc[1] = min(acc[1], t.first_sample)
                    acc[2] = max(acc[2], t.real_length)
    return total, counters
