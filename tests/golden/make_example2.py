"""Recovers BASELINE config 3's fixture from the reference's example2/example2.blend WITHOUT Blender (tools/blend_read.py
parses the file's own SDNA) and writes ear_b200/data/example2.npz.  Run in the container that has /root/reference:

    python tests/golden/make_example2.py

What is taken from the .blend (Blender 2.62, uncompressed): the three reflective meshes (objects with the add-on's
`is_surface` property: Plane.001 / .003 / .004 = 12 + 51 + 53 quads, identity transforms) with their per-face material
slots, split the way the exporter splits quads ([v0,v1,v2],[v0,v2,v3], blender/render_EAR/__init__.py:274); the two
materials' coefficients (ID properties refl_*, refr_*, exp_*); the scene settings (air absorption, f1/f3, num_samples);
the world position of the `Bach` emitter (parented to the piano); the linear F-curves of `Listener` and `Person`,
evaluated at the exporter's key frames 1, 51, ..., 451 of 500 @ 24 fps."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from blend_read import Blend  # noqa: E402

SRC = os.environ.get("EAR_REFERENCE", "/root/reference") + "/example2/example2.blend"
b = Blend(SRC)


def _i(v):
    return v[0] if isinstance(v, (bytes, bytearray)) else v


def read_prop(blk):
    o = blk[5]
    typ, name, sub = _i(b.get("IDProperty", o, "type")), b.get("IDProperty", o, "name"), _i(b.get("IDProperty", o, "subtype"))
    if typ == 0:
        blk2, ln = b.deref(b.get("IDProperty", o, "data.pointer")), b.get("IDProperty", o, "len")
        return name, b.d[blk2[5]:blk2[5] + ln].split(b"\0")[0].decode("latin1") if blk2 else ""
    if typ == 1:
        return name, b.get("IDProperty", o, "data.val")
    if typ == 2:
        return name, struct.unpack("<f", struct.pack("<i", b.get("IDProperty", o, "data.val")))[0]
    if typ == 8:
        return name, struct.unpack("<d", struct.pack("<ii", b.get("IDProperty", o, "data.val"), b.get("IDProperty", o, "data.val2")))[0]
    if typ == 6:
        out, p = {}, b.get("IDProperty", o, "data.group.first")
        while p:
            blk2 = b.deref(p)
            if not blk2:
                break
            n, v = read_prop(blk2)
            out[n] = v
            p = b.get("IDProperty", blk2[5], "next")
        return name, out
    return name, None


def idprops(sname, off):
    ptr = b.get(sname, off, "id.properties")
    blk = b.deref(ptr) if ptr else None
    return read_prop(blk)[1] if blk else {}


def fcurve_eval(keys, frame):
    """linear keys [(frame, value)], constant extrapolation (every curve of the file is linear, ipo == 1)"""
    if frame <= keys[0][0]:
        return keys[0][1]
    for (f0, v0), (f1, v1) in zip(keys, keys[1:]):
        if frame <= f1:
            return v0 + (v1 - v0) * (frame - f0) / (f1 - f0)
    return keys[-1][1]


def location_curves(off):
    out = {}
    adt = b.get("Object", off, "adt")
    if not adt:
        return out
    act = b.get("AnimData", b.deref(adt)[5], "action")
    p = b.get("bAction", b.deref(act)[5], "curves.first")
    _, bsz = b.layout("BezTriple")
    while p:
        o = b.deref(p)[5]
        rpb = b.deref(b.get("FCurve", o, "rna_path"))
        path = b.d[rpb[5]:rpb[5] + rpb[1]].split(b"\0")[0].decode()
        idx, n = b.get("FCurve", o, "array_index"), b.get("FCurve", o, "totvert")
        bz = b.deref(b.get("FCurve", o, "bezt"))[5]
        keys = []
        for k in range(n):
            vec = b.get("BezTriple", bz + k * bsz, "vec")
            assert _i(b.get("BezTriple", bz + k * bsz, "ipo")) == 1, "non-linear key"
            keys.append((vec[3], vec[4]))
        if path == "location":
            out[idx] = keys
        p = b.get("FCurve", o, "next")
    return out


objs = {b.get("Object", ob[5], "id.name")[2:]: ob[5] for ob in b.blocks_of(b"OB")}
mats = {b.get("Material", ma[5], "id.name")[2:]: idprops("Material", ma[5]) for ma in b.blocks_of(b"MA")}
tris, tri_mat, mat_names = [], [], []
for name in ("Plane.001", "Plane.003", "Plane.004"):
    off = objs[name]
    assert idprops("Object", off).get("is_surface") == 1
    assert np.allclose(b.get("Object", off, "loc"), 0) and np.allclose(b.get("Object", off, "size"), 1) and np.allclose(b.get("Object", off, "rot"), 0)
    me = b.deref(b.get("Object", off, "data"))[5]
    nv, nf, ncol = b.get("Mesh", me, "totvert"), b.get("Mesh", me, "totface"), b.get("Mesh", me, "totcol")
    _, vsz = b.layout("MVert")
    _, fsz = b.layout("MFace")
    mv, mf = b.deref(b.get("Mesh", me, "mvert"))[5], b.deref(b.get("Mesh", me, "mface"))[5]
    co = np.array([b.get("MVert", mv + k * vsz, "co") for k in range(nv)], np.float32)
    slots = b.deref(b.get("Mesh", me, "mat"))
    slot_ptrs = struct.unpack_from("<" + "Q" * ncol, b.d, slots[5])
    slot_names = [b.get("Material", b.deref(p)[5], "id.name")[2:] for p in slot_ptrs]
    faces = []
    for k in range(nf):
        v = [b.get("MFace", mf + k * fsz, f"v{i}") for i in (1, 2, 3, 4)]
        faces.append((v, b.get("MFace", mf + k * fsz, "mat_nr")))
    # the exporter writes one MESH block per material slot of an object (mi_to_fa, __init__.py:261-306)
    for mi in range(ncol):
        for v, mat_nr in faces:
            if mat_nr != mi:
                continue
            assert v[3] != 0, "triangle face: the fixture expects quads"
            for a, c, d in ((0, 1, 2), (0, 2, 3)):
                tris.append([co[v[a]], co[v[c]], co[v[d]]])
                if slot_names[mi] not in mat_names:
                    mat_names.append(slot_names[mi])
                tri_mat.append(mat_names.index(slot_names[mi]))
tris = np.asarray(tris, np.float32)
assert tris.shape[0] == 232, tris.shape
scene = idprops("Scene", b.blocks_of(b"SC")[0][5])
frames = list(range(1, 500, 50))
keys = np.array([(f - 1) / 24.0 for f in frames], np.float32)


def path_of(name):
    off = objs[name]
    loc = b.get("Object", off, "loc")
    cur = location_curves(off)
    return np.array([[fcurve_eval(cur[a], f) if a in cur else loc[a] for a in range(3)] for f in frames], np.float32)


bach = np.array(b.get("Object", objs["Bach"], "obmat"), np.float32)[12:15]
table = np.array([[[mats[m][f"refl_{x}"] for x in ("low", "mid", "high")],
                   [mats[m].get(f"refr_{x}", 0.0) for x in ("low", "mid", "high")],
                   [mats[m][f"exp_{x}"] for x in ("low", "mid", "high")]] for m in mat_names], np.float32)
out = os.path.join(ROOT, "ear_b200", "data", "example2.npz")
np.savez_compressed(out, tris=tris, tri_mat=np.asarray(tri_mat, np.int32), mat_names=np.array(mat_names), materials=table,
                    keys=keys, listener=path_of("Listener"), person=path_of("Person"), bach=bach,
                    air=np.array([scene["ab_low"], scene["ab_mid"], scene["ab_high"]], np.float32),
                    freq=np.array([scene["f1"], 2.0, scene["f3"]], np.float32), dry=np.float32(scene["dry"]),
                    num_samples=np.int32(scene["num_samples"]))
print("wrote", out, tris.shape, mat_names, "tri per material", np.bincount(tri_mat))
print("bounds", tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0))
print("listener", path_of("Listener")[[0, 5, 9]], "person", path_of("Person")[[0, 7, 9]], "bach", bach)
