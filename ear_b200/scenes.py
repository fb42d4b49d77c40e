"""Scene fixtures, emitted without Blender.

* `rt60_scene`       BASELINE config 1: testbench/RT60.blend's shoebox "Hall" (unit cube
                     scaled to 10 x 6 x 4 m, 12 triangles after the exporter's quad split
                     `[v0,v1,v2],[v0,v2,v3]`, blender/render_EAR/__init__.py:274).
* `example1_scene`   BASELINE config 2: a 40 x 52 x 18 m hall with a 0.36 m thick, 3.47 m
                     high partition, 44 triangles, stereo recorder with exporter defaults.
* `example2_scene`   BASELINE config 3: the reference's example2 (232 triangles, 2 materials, 10 keyframes,
                     3 sources, animated listener), recovered from its .blend (ear_b200/data/example2.npz).
* `synthetic_complex` BASELINE config 5: grid of coupled halls (C4 halls joined by door tunnels), 64 recorders.
* `synthetic_hall`   BASELINE config 4 generator: 60 x 40 x 20 m shoebox whose six walls are
                     tessellated to an exact triangle count with seeded +-5 cm displacement,
                     plus seeded interior box obstacles; 4 materials x n_bands coefficients.
"""
from __future__ import annotations

import numpy as np

from .earfile import MaterialDef, MeshDef, RecorderDef, SceneDef, SourceDef

_QUAD_FACES = ([0, 1, 2, 3], [4, 7, 6, 5], [0, 4, 5, 1], [1, 5, 6, 2], [2, 6, 7, 3], [4, 0, 3, 7])


def _box_verts(lo, hi) -> np.ndarray:
    """8 corners in the vertex order of RT60.blend's `Hall` cube."""
    (x0, y0, z0), (x1, y1, z1) = lo, hi
    return np.array([(x1, y1, z0), (x1, y0, z0), (x0, y0, z0), (x0, y1, z0),
                     (x1, y1, z1), (x1, y0, z1), (x0, y0, z1), (x0, y1, z1)], np.float32)


def box_triangles(lo, hi, faces=_QUAD_FACES) -> np.ndarray:
    v = _box_verts(lo, hi)
    tris = []
    for q in faces:
        tris.append([v[q[0]], v[q[1]], v[q[2]]])
        tris.append([v[q[0]], v[q[2]], v[q[3]]])
    return np.asarray(tris, np.float32)


def rt60_scene(dims=(10.0, 6.0, 4.0), refl=(0.95, 0.95, 0.95), refr=(0.0, 0.0, 0.0), spec=(0.0, 0.5, 0.0),
               air=(0.0, 0.0, 0.0), samples=1000000, wav="/tmp/click.wav", stereo=False) -> SceneDef:
    dx, dy, dz = dims
    tris = box_triangles((-dx / 2, -dy / 2, 0.0), (dx / 2, dy / 2, dz))
    sc = SceneDef(samples=samples, air_absorption=air)
    sc.materials.append(MaterialDef("Hall_Material", refl, refr, spec))
    sc.meshes.append(MeshDef("Hall_Material", tris))
    # listener at x = d0/2 - 1, source mirrored (testbench script embedded in RT60.blend)
    sc.sources.append(SourceDef([wav], position=(-(dx / 2 - 1.0), 0.0, 1.6)))
    sc.recorders.append(RecorderDef("/tmp/output.wav", position=(dx / 2 - 1.0, 0.0, 1.6), stereo=stereo))
    return sc


def example1_scene(samples=1000000, wav="/tmp/click.wav", stereo=True) -> SceneDef:
    """Hall 40 x 52 x 18 m (12 triangles) + free-standing partition 0.36 m thick, 3.47 m high.
    The partition and a stage block are closed boxes, split into 16 further quads so the scene has
    the 22 quads / 44 triangles of example1.blend (the .blend's exact vertex table is not
    recoverable without Blender; dimensions, material and source/listener placement are)."""
    hall = box_triangles((-20.0, -26.0, 0.0), (20.0, 26.0, 18.0))
    # partition: 5 visible quads (no bottom), split lengthwise into two boxes -> 10 quads
    part_a = box_triangles((-0.18, -14.0, 0.0), (0.18, 0.0, 3.47), faces=_QUAD_FACES[1:])
    part_b = box_triangles((-0.18, 0.0, 0.0), (0.18, 14.0, 3.47), faces=_QUAD_FACES[1:])
    # stage riser: 6 quads
    stage = box_triangles((12.0, -8.0, 0.0), (18.0, 8.0, 0.9))
    tris = np.concatenate([hall, part_a, part_b, stage])
    assert tris.shape[0] == 44
    sc = SceneDef(samples=samples, air_absorption=(0.001, 0.0015, 0.003))
    sc.materials.append(MaterialDef("Material", (0.95, 0.98, 0.99), (0.0, 0.0, 0.0), (0.3, 0.6, 0.9)))
    sc.meshes.append(MeshDef("Material", tris))
    sc.sources.append(SourceDef([wav], position=(-5.0, 5.0, 1.6)))
    sc.recorders.append(RecorderDef("/tmp/example1.out.wav", position=(5.0, -5.0, 1.6), stereo=stereo,
                                    right_ear=(-1.0, 0.0, 0.0), head_size=0.2, head_absorption=(0.1, 0.3, 0.9)))
    return sc


def example2_scene(samples=100000, bach=("/tmp/bach-low.wav", "/tmp/bach-mid.wav", "/tmp/bach-high.wav"),
                   steps="/tmp/steps.wav", door="/tmp/door.wav", out="/tmp/example2.out.wav") -> SceneDef:
    """BASELINE config 3: the reference's example2 -- 232 triangles (three meshes of 24 + 102 + 106, split over two
    materials), air absorption (0.01, 0.03, 0.05), 10 keyframes (frames 1, 51, ..., 451 of 500 @ 24 fps), three sources
    x 10 keyframes x 3 bands = 90 contexts, a keyframed mono listener.  Geometry, materials, settings, the `Bach`
    position and the Listener / Person paths are the reference's own, recovered from example2/example2.blend without
    Blender (tests/golden/make_example2.py -> ear_b200/data/example2.npz).  Substitution: the two
    storyboard sources the add-on would synthesise with bpy ray casts (foot steps of `Listener` and `Person`, the door
    of `Portal`) are a steps recording moving with the Listener and door.wav moving with the Person.  `bach` is the
    triple-band source (3SRC: bach-bwv999-{low,mid,high}.wav, 705 600 samples each in the reference)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "example2.npz"))
    sc = SceneDef(samples=samples, air_absorption=[float(x) for x in d["air"]], drylevel=float(d["dry"]),
                  freq=[float(x) for x in d["freq"]], keys=[float(k) for k in d["keys"]])
    names = [str(n) for n in d["mat_names"]]
    for m, name in enumerate(names):
        refl, refr, spec = d["materials"][m]
        sc.materials.append(MaterialDef(name, [float(x) for x in refl], [float(x) for x in refr], [float(x) for x in spec]))
    tris, tri_mat = d["tris"], d["tri_mat"]
    # MESH blocks in file order: runs of equal material (one block per object and material slot, as the exporter writes)
    start = 0
    for i in range(1, tris.shape[0] + 1):
        if i == tris.shape[0] or tri_mat[i] != tri_mat[start]:
            sc.meshes.append(MeshDef(names[int(tri_mat[start])], tris[start:i].astype(np.float32)))
            start = i
    n_keys = len(sc.keys)
    sc.sources.append(SourceDef(list(bach), animation=np.tile(d["bach"][None, :], (n_keys, 1)).astype(np.float32)))
    # foot steps sound at floor level under the walker (heads are 1.65 m above their floors in the file)
    feet = np.array([0.0, 0.0, -1.65], np.float32)
    sc.sources.append(SourceDef([steps], animation=(d["listener"] + feet).astype(np.float32)))
    sc.sources.append(SourceDef([door], animation=(d["person"] + feet).astype(np.float32)))
    sc.recorders.append(RecorderDef(out, animation=d["listener"].astype(np.float32)))
    return sc


# ----------------------------------------------------------------------------------
# synthetic hall (C4 / C5)
# ----------------------------------------------------------------------------------
def _wall_grid(origin, eu, ev, normal, nu, nv, rng, amp):
    """Tessellate the parallelogram origin + s*eu + t*ev into nu x nv quads (2 triangles each);
    interior vertices are displaced along `normal` by U(-amp, amp); border vertices stay put so
    adjacent walls stay welded."""
    s = np.linspace(0.0, 1.0, nu + 1)
    t = np.linspace(0.0, 1.0, nv + 1)
    S, T = np.meshgrid(s, t, indexing="ij")
    P = (np.asarray(origin)[None, None, :] + S[..., None] * np.asarray(eu)[None, None, :]
         + T[..., None] * np.asarray(ev)[None, None, :])
    d = rng.uniform(-amp, amp, size=S.shape)
    d[0, :] = d[-1, :] = 0.0
    d[:, 0] = d[:, -1] = 0.0
    P = P + d[..., None] * np.asarray(normal)[None, None, :]
    a, b, c, e = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    t1 = np.stack([a, b, c], axis=2)
    t2 = np.stack([a, c, e], axis=2)
    return np.concatenate([t1.reshape(-1, 3, 3), t2.reshape(-1, 3, 3)]).astype(np.float32)


def _split_triangles(tris: np.ndarray, extra: int) -> np.ndarray:
    """Raise the triangle count by exactly `extra` by bisecting the first `extra` triangles."""
    if extra <= 0:
        return tris
    head, tail = tris[:extra], tris[extra:]
    mid = (head[:, 1] + head[:, 2]) * np.float32(0.5)
    a = np.stack([head[:, 0], head[:, 1], mid], axis=1)
    b = np.stack([head[:, 0], mid, head[:, 2]], axis=1)
    return np.concatenate([a, b, tail]).astype(np.float32)


def synthetic_hall(n_tris=1_000_000, n_obstacles=2000, n_bands=8, seed=0, dims=(60.0, 40.0, 20.0),
                   n_recorders=1, samples=10_000_000):
    """Returns (SceneDef, material_table[M, n_bands, 4]).  The SceneDef carries bands 0..2 only
    (the .ear format is hard-wired to three, src/EAR.cpp:176); the full table is what the
    ABI consumes when n_bands > 3 (documented extension)."""
    rng = np.random.default_rng(seed)
    dx, dy, dz = dims
    x0, x1, y0, y1 = -dx / 2, dx / 2, -dy / 2, dy / 2
    n_box_tris = 12 * n_obstacles
    n_wall = n_tris - n_box_tris
    assert n_wall >= 12, "triangle budget too small for the obstacle count"
    # (origin, eu, ev, inward normal, material)
    walls = [
        ((x0, y0, 0.0), (dx, 0, 0), (0, dy, 0), (0, 0, 1), 0),      # floor
        ((x0, y0, dz), (dx, 0, 0), (0, dy, 0), (0, 0, -1), 1),      # ceiling
        ((x0, y0, 0.0), (dx, 0, 0), (0, 0, dz), (0, 1, 0), 1),      # y = y0
        ((x0, y1, 0.0), (dx, 0, 0), (0, 0, dz), (0, -1, 0), 1),     # y = y1
        ((x0, y0, 0.0), (0, dy, 0), (0, 0, dz), (1, 0, 0), 1),      # x = x0
        ((x1, y0, 0.0), (0, dy, 0), (0, 0, dz), (-1, 0, 0), 1),     # x = x1
    ]
    areas = np.array([np.linalg.norm(np.cross(w[1], w[2])) for w in walls])
    per_mat = {0: [], 1: [], 2: [], 3: []}
    made = 0
    for w, area in zip(walls, areas):
        quads = max(1, int(np.floor(n_wall * area / areas.sum() / 2.0)))
        lu, lv = np.linalg.norm(w[1]), np.linalg.norm(w[2])
        nu = max(1, int(np.floor(np.sqrt(quads * lu / lv))))
        nv = max(1, quads // nu)
        g = _wall_grid(w[0], w[1], w[2], w[3], nu, nv, rng, 0.05)
        per_mat[w[4]].append(g)
        made += g.shape[0]
    extra = n_wall - made
    assert extra >= 0
    per_mat[0][0] = _split_triangles(per_mat[0][0], extra)
    # obstacles: seat-like boxes on the floor; every 4th one is a thin glazed screen
    for i in range(n_obstacles):
        cx = rng.uniform(x0 + 3.0, x1 - 3.0)
        cy = rng.uniform(y0 + 3.0, y1 - 3.0)
        if i % 4 == 0:
            sx, sy, sz = (0.06, rng.uniform(0.8, 2.0), rng.uniform(1.0, 2.2))
            if rng.uniform() < 0.5:
                sx, sy = sy, sx
            mat = 3
        else:
            sx, sy, sz = rng.uniform(0.4, 1.2), rng.uniform(0.4, 1.2), rng.uniform(0.4, 1.1)
            mat = 2
        z_lo = 0.06
        per_mat[mat].append(box_triangles((cx - sx / 2, cy - sy / 2, z_lo), (cx + sx / 2, cy + sy / 2, z_lo + sz)))
    # materials: refl ~ U(0.6, 0.97), spec ~ U(0, 0.9) per band; material 3 is glazing (0.7 / 0.24)
    table = np.zeros((4, n_bands, 4), np.float32)
    refl = rng.uniform(0.6, 0.97, size=(4, n_bands)).astype(np.float32)
    spec = rng.uniform(0.0, 0.9, size=(4, n_bands)).astype(np.float32)
    refr = np.zeros((4, n_bands), np.float32)
    refl[3, :] = 0.7
    refr[3, :] = 0.24
    sc = SceneDef(samples=samples, air_absorption=(0.001, 0.0015, 0.003))
    names = ["floor", "shell", "seats", "glazing"]
    for m in range(4):
        sc.materials.append(MaterialDef(names[m], [float(x) for x in refl[m, :3]], [float(x) for x in refr[m, :3]],
                                        [float(x) for x in spec[m, :3]]))
    eps = np.float32(1e-9)
    for m in range(4):
        for b in range(n_bands):
            a = np.float32(1.0)
            a = np.float32(a - np.float32(refl[m, b] - eps))
            a = np.float32(a - np.float32(refr[m, b] - eps))
            table[m, b] = (refl[m, b], refr[m, b], np.float32(np.float32(1.0) - a), spec[m, b])
    for m in range(4):
        if per_mat[m]:
            sc.meshes.append(MeshDef(names[m], np.concatenate(per_mat[m]).astype(np.float32)))
    assert sc.triangles().shape[0] == n_tris, (sc.triangles().shape[0], n_tris)
    sc.sources.append(SourceDef(["/tmp/click.wav"], position=(x0 + 8.0, 1.0, 1.7)))
    rrng = np.random.default_rng(seed + 1)
    for r in range(n_recorders):
        if r == 0:
            pos = (x1 - 10.0, -2.0, 1.8)
        else:
            pos = (rrng.uniform(x0 + 2, x1 - 2), rrng.uniform(y0 + 2, y1 - 2), rrng.uniform(1.2, dz - 2.0))
        sc.recorders.append(RecorderDef(f"/tmp/hall.{r}.wav", position=pos))
    return sc, table


def _wall_grid_holed(origin, eu, ev, normal, nu, nv, rng, amp, hole):
    """_wall_grid with the quads [i0, i1) x [j0, j1) left out (a door).  Vertices on and inside the hole's border are
    not displaced, so the border is a planar rectangle loop.  Returns (triangles, border loop [k][3])."""
    i0, i1, j0, j1 = hole
    s = np.linspace(0.0, 1.0, nu + 1)
    t = np.linspace(0.0, 1.0, nv + 1)
    S, T = np.meshgrid(s, t, indexing="ij")
    P = (np.asarray(origin)[None, None, :] + S[..., None] * np.asarray(eu)[None, None, :]
         + T[..., None] * np.asarray(ev)[None, None, :])
    d = rng.uniform(-amp, amp, size=S.shape)
    d[0, :] = d[-1, :] = 0.0
    d[:, 0] = d[:, -1] = 0.0
    d[i0:i1 + 1, j0:j1 + 1] = 0.0
    P = P + d[..., None] * np.asarray(normal)[None, None, :]
    a, b, c, e = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    keep = np.ones((nu, nv), bool)
    keep[i0:i1, j0:j1] = False
    t1 = np.stack([a, b, c], axis=2)[keep]
    t2 = np.stack([a, c, e], axis=2)[keep]
    loop = ([P[i, j0] for i in range(i0, i1)] + [P[i1, j] for j in range(j0, j1)]
            + [P[i, j1] for i in range(i1, i0, -1)] + [P[i0, j] for j in range(j1, j0, -1)])
    return np.concatenate([t1, t2]).astype(np.float32), np.asarray(loop, np.float64)


def synthetic_complex(n_tris=10_000_000, n_obstacles=20000, n_bands=3, seed=0, hall=(60.0, 40.0, 20.0), grid=(4, 2),
                      n_recorders=64, samples=10_000_000, gap=0.6):
    """BASELINE config 5 (SURVEY.md section 8d "C5"): a complex of grid[0] x grid[1] coupled halls -- the C4 hall
    replicated on a grid, neighbours joined by a door-sized tunnel through their facing walls -- tessellated to an
    exact triangle count with the same seeded +-5 cm wall displacement, seeded box obstacles in every hall, the same
    four materials, ONE point source (hall 0) and `n_recorders` mono recorders spread over all halls.
    Returns (SceneDef, material_table[M, n_bands, 4]) like synthetic_hall."""
    rng = np.random.default_rng(seed)
    gx, gy = grid
    n_halls = gx * gy
    dx, dy, dz = hall
    per_hall_obst = n_obstacles // n_halls
    n_box_tris = 12 * per_hall_obst * n_halls
    per_mat = {0: [], 1: [], 2: [], 3: []}
    # wall tessellation shared by all halls (same dims -> facing walls have matching grids)
    spec = [((dx, 0, 0), (0, dy, 0)), ((dx, 0, 0), (0, dy, 0)), ((dx, 0, 0), (0, 0, dz)), ((dx, 0, 0), (0, 0, dz)),
            ((0, dy, 0), (0, 0, dz)), ((0, dy, 0), (0, 0, dz))]
    areas = np.array([np.linalg.norm(np.cross(e[0], e[1])) for e in spec])
    budget = (n_tris - n_box_tris) // n_halls
    budget -= 4096                                   # head-room for door tunnels and the exact-count top-up
    dims_uv = []
    for (eu, ev), area in zip(spec, areas):
        quads = max(1, int(np.floor(budget * area / areas.sum() / 2.0)))
        lu, lv = np.linalg.norm(eu), np.linalg.norm(ev)
        nu = max(2, int(np.floor(np.sqrt(quads * lu / lv))))
        nv = max(2, quads // nu)
        dims_uv.append((nu, nv))

    def door(nu, nv, lu, lv):
        """quad-index rectangle of a 6 m wide, 3.2 m high door in the middle of a wall of lu x lv metres"""
        w = max(1, int(round(6.0 / lu * nu)))
        h = max(1, int(round(3.2 / lv * nv)))
        i0 = (nu - w) // 2
        return i0, i0 + w, 0, min(h, nv - 1)

    loops = {}                                       # (hall index, wall index) -> border loop of its door
    made = 0
    origins = []
    for cy in range(gy):
        for cx in range(gx):
            h = cy * gx + cx
            x0, y0 = cx * (dx + gap), cy * (dy + gap)
            origins.append((x0, y0))
            x1, y1 = x0 + dx, y0 + dy
            walls = [((x0, y0, 0.0), (0, 0, 1), 0, None), ((x0, y0, dz), (0, 0, -1), 1, None),
                     ((x0, y0, 0.0), (0, 1, 0), 1, cy > 0), ((x0, y1, 0.0), (0, -1, 0), 1, cy + 1 < gy),
                     ((x0, y0, 0.0), (1, 0, 0), 1, cx > 0), ((x1, y0, 0.0), (-1, 0, 0), 1, cx + 1 < gx)]
            for wi, ((org, nrm, mat, has_door), (eu, ev), (nu, nv)) in enumerate(zip(walls, spec, dims_uv)):
                if has_door:
                    g, loop = _wall_grid_holed(org, eu, ev, nrm, nu, nv, rng, 0.05,
                                               door(nu, nv, np.linalg.norm(eu), np.linalg.norm(ev)))
                    loops[(h, wi)] = loop
                else:
                    g = _wall_grid(org, eu, ev, nrm, nu, nv, rng, 0.05)
                per_mat[mat].append(g)
                made += g.shape[0]
    # tunnels: join the two border loops of every pair of facing doors (same grid -> vertex k faces vertex k)
    def tunnel(la, lb):
        n = la.shape[0]
        a0, a1 = la, np.roll(la, -1, axis=0)
        b0, b1 = lb, np.roll(lb, -1, axis=0)
        return np.concatenate([np.stack([a0, a1, b1], axis=1), np.stack([a0, b1, b0], axis=1)]).astype(np.float32).reshape(2 * n, 3, 3)
    for cy in range(gy):
        for cx in range(gx):
            h = cy * gx + cx
            if cx + 1 < gx:
                t = tunnel(loops[(h, 5)], loops[(h + 1, 4)])
                per_mat[1].append(t); made += t.shape[0]
            if cy + 1 < gy:
                t = tunnel(loops[(h, 3)], loops[(h + gx, 2)])
                per_mat[1].append(t); made += t.shape[0]
    extra = n_tris - n_box_tris - made
    assert extra >= 0, extra
    while extra > 0:                                  # top up to the exact count by bisecting wall triangles
        for m in (0, 1):
            for k in range(len(per_mat[m])):
                take = min(extra, per_mat[m][k].shape[0])
                per_mat[m][k] = _split_triangles(per_mat[m][k], take)
                extra -= take
    # obstacles, per hall, as in synthetic_hall
    for h, (ox, oy) in enumerate(origins):
        for i in range(per_hall_obst):
            cx_ = rng.uniform(ox + 3.0, ox + dx - 3.0)
            cy_ = rng.uniform(oy + 3.0, oy + dy - 3.0)
            if i % 4 == 0:
                sx, sy, sz = (0.06, rng.uniform(0.8, 2.0), rng.uniform(1.0, 2.2))
                if rng.uniform() < 0.5:
                    sx, sy = sy, sx
                mat = 3
            else:
                sx, sy, sz = rng.uniform(0.4, 1.2), rng.uniform(0.4, 1.2), rng.uniform(0.4, 1.1)
                mat = 2
            per_mat[mat].append(box_triangles((cx_ - sx / 2, cy_ - sy / 2, 0.06), (cx_ + sx / 2, cy_ + sy / 2, 0.06 + sz)))
    table = np.zeros((4, n_bands, 4), np.float32)
    refl = rng.uniform(0.6, 0.97, size=(4, n_bands)).astype(np.float32)
    spc = rng.uniform(0.0, 0.9, size=(4, n_bands)).astype(np.float32)
    refr = np.zeros((4, n_bands), np.float32)
    refl[3, :] = 0.7
    refr[3, :] = 0.24
    sc = SceneDef(samples=samples, air_absorption=(0.001, 0.0015, 0.003))
    names = ["floor", "shell", "seats", "glazing"]
    eps = np.float32(1e-9)
    for m in range(4):
        sc.materials.append(MaterialDef(names[m], [float(x) for x in refl[m, :3]], [float(x) for x in refr[m, :3]],
                                        [float(x) for x in spc[m, :3]]))
        for b in range(n_bands):
            a = np.float32(1.0)
            a = np.float32(a - np.float32(refl[m, b] - eps))
            a = np.float32(a - np.float32(refr[m, b] - eps))
            table[m, b] = (refl[m, b], refr[m, b], np.float32(np.float32(1.0) - a), spc[m, b])
    for m in range(4):
        if per_mat[m]:
            sc.meshes.append(MeshDef(names[m], np.concatenate(per_mat[m]).astype(np.float32)))
    assert sc.triangles().shape[0] == n_tris, (sc.triangles().shape[0], n_tris)
    sc.sources.append(SourceDef(["/tmp/click.wav"], position=(origins[0][0] + 8.0, origins[0][1] + 21.0, 1.7)))
    rrng = np.random.default_rng(seed + 1)
    for r in range(n_recorders):
        ox, oy = origins[r % n_halls]
        pos = (rrng.uniform(ox + 2, ox + dx - 2), rrng.uniform(oy + 2, oy + dy - 2), rrng.uniform(1.2, dz - 2.0))
        sc.recorders.append(RecorderDef(f"/tmp/complex.{r}.wav", position=pos))
    return sc, table


def air_factors(n_bands: int) -> np.ndarray:
    """Per-band air survival factor per metre (1 - absorption) for the synthetic hall's bands."""
    ab = np.linspace(0.0005, 0.004, n_bands).astype(np.float32)
    return (np.float32(1.0) - ab).astype(np.float32)


def write_click_wav(path: str, n: int = 443) -> str:
    """A short 16-bit mono 44.1 kHz click (same length as example1/click.wav: 443 samples);
    the reference refuses to load a scene whose source wav is missing (src/SoundFile.cpp:41-43)."""
    import wave
    t = np.arange(n, dtype=np.float64)
    sig = np.exp(-t / 40.0) * np.sin(2 * np.pi * 1000.0 * t / 44100.0)
    pcm = np.round(sig / np.abs(sig).max() * 30000.0).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    return path
