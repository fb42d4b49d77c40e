"""TEST INFRASTRUCTURE -- ctypes binding of oracle/libear_oracle.so (the CPU restatement) and a
runner for the reference binaries under oracle/_ref/.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs; never by ear_b200/."""
from __future__ import annotations

import ctypes as C
import os
import re
import struct
import subprocess
from typing import Sequence

import numpy as np

from ear_b200.api import ContextC, RecorderC, Context, pack_contexts, pack_recorders, Track

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libear_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
RNG_RAND, RNG_PHILOX = 0, 1

_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "libear_oracle.so"], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        vp, i32, i64, u32, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float
        l.oracle_scene_create.restype = vp
        l.oracle_scene_create.argtypes = [vp, vp, i32, vp, i32, i32]
        l.oracle_scene_destroy.argtypes = [vp]
        l.oracle_scene_set_emitters.argtypes = [vp, vp, i32]
        l.oracle_first_hit.argtypes = [vp, vp, vp, i64, vp, vp]
        l.oracle_occluded.argtypes = [vp, vp, vp, i64, vp]
        l.oracle_render.restype = vp
        l.oracle_render.argtypes = [vp, C.POINTER(ContextC), i32, C.POINTER(RecorderC), i32, i32, i32, u64, i64, i64, i32]
        l.oracle_render_free.argtypes = [vp]
        l.oracle_render_counters.argtypes = [vp, vp]
        l.oracle_render_track_info.argtypes = [vp, i32, i32, i32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        l.oracle_render_track_copy.argtypes = [vp, i32, i32, i32, vp, u32]
        l.oracle_trace_paths.argtypes = [vp, C.POINTER(ContextC), i32, i32, u64, i64, i64, vp, vp]
        l.oracle_power.argtypes = [vp, u32, u32, f32]
        l.oracle_maximum.argtypes = [vp, u32, u32]
        l.oracle_maximum.restype = f32
        l.oracle_get_length.argtypes = [vp, u32, u32, u32, f32]
        l.oracle_get_length.restype = u32
        l.oracle_t60.argtypes = [vp, u32, u32]
        l.oracle_t60.restype = f32
        l.oracle_sabine_eyring.argtypes = [vp, vp, i32, f32, f32, C.POINTER(f32), C.POINTER(f32)]
        l.oracle_convolve.argtypes = [vp, u32, u32, u32, vp, u32, u32, u32, vp, u32, u32, vp]
        _lib = l
    return _lib


class OracleScene:
    def __init__(self, verts, tri_material, materials):
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3, 3)
        self.tri_material = np.ascontiguousarray(tri_material, np.int32)
        self.materials = np.ascontiguousarray(materials, np.float32)
        self.h = lib().oracle_scene_create(self.verts.ctypes.data, self.tri_material.ctypes.data, self.verts.shape[0],
                                           self.materials.ctypes.data, self.materials.shape[0], self.materials.shape[1])

    @classmethod
    def from_def(cls, scene_def, materials=None):
        from ear_b200.api import emitter_table
        tab = scene_def.material_table() if materials is None else materials
        scene = cls(scene_def.triangles(), scene_def.triangle_materials(), tab)
        em = emitter_table(scene_def)
        if em is not None:
            scene.set_emitters(em)
        return scene

    def set_emitters(self, verts):
        self.emitters = np.ascontiguousarray(verts, np.float32).reshape(-1, 3, 3)
        lib().oracle_scene_set_emitters(self.h, self.emitters.ctypes.data, self.emitters.shape[0])

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_scene_destroy(self.h)
            self.h = None

    def first_hit(self, origins, dirs):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        idx = np.empty(o.shape[0], np.int32)
        t = np.empty(o.shape[0], np.float32)
        lib().oracle_first_hit(self.h, o.ctypes.data, d.ctypes.data, o.shape[0], idx.ctypes.data, t.ctypes.data)
        return idx, t

    def occluded(self, p, x):
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 3)
        out = np.empty(p.shape[0], np.uint8)
        lib().oracle_occluded(self.h, p.ctypes.data, x.ctypes.data, p.shape[0], out.ctypes.data)
        return out

    def render(self, contexts: Sequence[Context], recorders, max_bounces=1000, rng_mode=RNG_PHILOX, seed=1,
               first_ray=0, ray_count=-1, finalise=True):
        """Returns (tracks[ctx][rec][track] as ear_b200.api.Track with .length in data.size, counters dict)."""
        ctx = pack_contexts(contexts)
        rec, n_rec = pack_recorders(recorders, len(contexts))
        h = lib().oracle_render(self.h, ctx, len(contexts), rec, n_rec, max_bounces, rng_mode, seed, first_ray,
                                ray_count, 1 if finalise else 0)
        try:
            cnt = np.zeros(5, np.uint64)
            lib().oracle_render_counters(h, cnt.ctypes.data)
            tracks = []
            for c in range(len(contexts)):
                per_rec = []
                for r in range(n_rec):
                    n_tracks = 2 if rec[c * n_rec + r].kind == 2 else 1
                    pair = []
                    for k in range(n_tracks):
                        f, rl, ln = C.c_uint32(), C.c_uint32(), C.c_uint32()
                        lib().oracle_render_track_info(h, c, r, k, C.byref(f), C.byref(rl), C.byref(ln))
                        data = np.empty(ln.value, np.float32)
                        lib().oracle_render_track_copy(h, c, r, k, data.ctypes.data, ln.value)
                        pair.append(Track(data, f.value, rl.value))
                    per_rec.append(pair)
                tracks.append(per_rec)
            counters = dict(zip(["rays", "segments", "occlusion_queries", "contributions", "bin_updates"],
                                [int(x) for x in cnt]))
            return tracks, counters
        finally:
            lib().oracle_render_free(h)

    def trace_paths(self, context: Context, ctx_index, n, max_bounces, seed, first_ray=0):
        hits = np.empty((n, max_bounces), np.int32)
        state = np.empty((n, 8), np.float32)
        cc = context.to_c()
        lib().oracle_trace_paths(self.h, C.byref(cc), ctx_index, max_bounces, seed, first_ray, n, hits.ctypes.data,
                                 state.ctypes.data)
        return hits, state

    def sabine_eyring(self, mesh_tri_counts, kept_mid, air_mid):
        counts = np.ascontiguousarray(mesh_tri_counts, np.int32)
        s, e = C.c_float(), C.c_float()
        lib().oracle_sabine_eyring(self.h, counts.ctypes.data, counts.shape[0], kept_mid, air_mid, C.byref(s), C.byref(e))
        return s.value, e.value


def post_t60(tracks_per_context, which=(0, 0, 0)):
    """Host post chain of `EAR calc T60` (src/EAR.cpp:209-228, 259-262) on oracle or GPU tracks:
    Power(0.335) on every track, global max, threshold max/256, Truncate(getLength), T60 of the
    first context's first recorder's first track.  tracks_per_context: [ctx][rec][track] of Track."""
    l = lib()
    mx = 0.0
    work = []
    for ctx in tracks_per_context:
        for rec in ctx:
            for tr in rec:
                d = np.ascontiguousarray(tr.data, np.float32).copy()
                l.oracle_power(d.ctypes.data, tr.first_sample, tr.real_length, 0.335)
                m = l.oracle_maximum(d.ctypes.data, tr.first_sample, tr.real_length)
                mx = max(mx, m)
                work.append((d, tr))
    thr = np.float32(mx) / np.float32(256.0)
    out = []
    idx = 0
    for ctx in tracks_per_context:
        for rec in ctx:
            ln = 0
            for tr in rec:
                d, _ = work[idx + rec.index(tr)]
                ln = max(ln, l.oracle_get_length(d.ctypes.data, tr.first_sample, tr.real_length, d.shape[0], thr))
            for tr in rec:
                d, _ = work[idx]
                idx += 1
                out.append((d, tr.first_sample, max(1, ln)))
    d, first, real = out[0]
    return float(l.oracle_t60(d.ctypes.data, first, real))


def convolve(response: Track, dry, offset=0, response2: Track = None) -> np.ndarray:
    """RecorderTrack::Process restated (src/Recorder.cpp:247-292); returns the raw result samples."""
    dry = np.ascontiguousarray(dry, np.float32)
    r1 = np.ascontiguousarray(response.data, np.float32)
    first, length = response.first_sample, response.real_length
    r2 = None
    if response2 is not None:
        r2 = np.ascontiguousarray(response2.data, np.float32)
        first, length = min(first, response2.first_sample), max(length, response2.real_length)
    out = np.zeros(max(3 * 44100, dry.shape[0] + offset + length), np.float32)
    lib().oracle_convolve(r1.ctypes.data, r1.shape[0], response.first_sample, response.real_length,
                          r2.ctypes.data if r2 is not None else None, r2.shape[0] if r2 is not None else 0,
                          response2.first_sample if response2 is not None else 0,
                          response2.real_length if response2 is not None else 0,
                          dry.ctypes.data, dry.shape[0], offset, out.ctypes.data)
    return out


# ---------------- the reference itself (oracle/_ref) ----------------
def ref_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ref_harness")) and os.path.exists(os.path.join(REF_DIR, "EAR_ref"))


def ref_first_hit(ear_path, origins, dirs, tmp):
    rays = np.concatenate([np.asarray(origins, np.float32).reshape(-1, 3), np.asarray(dirs, np.float32).reshape(-1, 3)], 1)
    rin, rout = os.path.join(tmp, "rays.bin"), os.path.join(tmp, "hits.bin")
    rays.tofile(rin)
    subprocess.run([os.path.join(REF_DIR, "ref_harness"), "firsthit", ear_path, rin, rout], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    raw = np.fromfile(rout, np.uint8).reshape(-1, 32)
    idx = raw[:, :4].copy().view(np.int32).reshape(-1)
    rest = raw[:, 4:].copy().view(np.float32).reshape(-1, 7)
    return idx, rest[:, 0].copy(), rest[:, 1:4].copy(), rest[:, 4:7].copy()


def ref_occluded(ear_path, p, x, tmp):
    segs = np.concatenate([np.asarray(p, np.float32).reshape(-1, 3), np.asarray(x, np.float32).reshape(-1, 3)], 1)
    sin, sout = os.path.join(tmp, "segs.bin"), os.path.join(tmp, "occ.bin")
    segs.tofile(sin)
    subprocess.run([os.path.join(REF_DIR, "ref_harness"), "occluded", ear_path, sin, sout], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return np.fromfile(sout, np.uint8)


def parse_ref_tracks(path):
    """Reads the dump written by `ref_harness render`: [ctx][rec][track] of Track + context headers."""
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    n_ctx, n_rec = struct.unpack_from("<ii", buf, pos)
    pos += 8
    tracks, headers = [], []
    for _ in range(n_ctx):
        headers.append(struct.unpack_from("<iii", buf, pos))
        pos += 12
        per_rec = []
        for _ in range(n_rec):
            (nt,) = struct.unpack_from("<i", buf, pos)
            pos += 4
            pair = []
            for _ in range(nt):
                first, real = struct.unpack_from("<II", buf, pos)
                pos += 8
                data = np.frombuffer(buf, np.float32, real + 1, pos).copy()
                pos += 4 * (real + 1)
                pair.append(Track(data, first, real))
            per_rec.append(pair)
        tracks.append(per_rec)
    return tracks, headers


def ref_render(ear_path, seed, out_path, t60_only=False, threads=1, timeout=None):
    """Runs the reference's Scene::Render for every context; returns (tracks, headers, stats dict)."""
    cmd = [os.path.join(REF_DIR, "ref_harness"), "render", ear_path, str(seed), out_path]
    if t60_only:
        cmd.append("t60")
    cmd.append(f"threads={threads}")
    r = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=timeout)
    m = re.search(r"REF_RENDER (.*)", r.stdout)
    stats = {k: float(v) for k, v in (kv.split("=") for kv in m.group(1).split())}
    tracks, headers = parse_ref_tracks(out_path)
    return tracks, headers, stats


def ref_calc_t60(ear_path, seed, timeout=None):
    """`EAR_ref calc T60 <file>` with rand() seeded via EAR_REF_SEED; returns (T60_ear, sabine, eyring)."""
    env = dict(os.environ, EAR_REF_SEED=str(seed))
    r = subprocess.run([os.path.join(REF_DIR, "EAR_ref"), "calc", "T60", ear_path], capture_output=True, text=True,
                       env=env, timeout=timeout, stdin=subprocess.DEVNULL)
    vals = [float(x) for x in re.findall(r"T60_\w+\s*: ([-0-9.naninf]+)s", r.stdout)]
    if len(vals) != 3:
        raise RuntimeError("EAR_ref failed: " + r.stdout[-400:])
    return tuple(vals)


def post_all(tracks_per_context, exponent=0.335, divisor=256.0):
    """The whole post chain of Render() (src/EAR.cpp:209-228) on every track: returns (maximum, [[[(data, first_sample,
    real_length, t60)]]]) in [ctx][rec][track] order.  Same steps as post_t60, but nothing is thrown away."""
    l = lib()
    mx = 0.0
    powered = []
    for ctx in tracks_per_context:
        pc = []
        for rec in ctx:
            pr = []
            for tr in rec:
                d = np.ascontiguousarray(tr.data, np.float32).copy()
                l.oracle_power(d.ctypes.data, tr.first_sample, tr.real_length, exponent)
                mx = max(mx, l.oracle_maximum(d.ctypes.data, tr.first_sample, tr.real_length))
                pr.append(d)
            pc.append(pr)
        powered.append(pc)
    thr = np.float32(mx) / np.float32(divisor)
    out = []
    for ctx, pc in zip(tracks_per_context, powered):
        oc = []
        for rec, pr in zip(ctx, pc):
            ln = 0
            if any(tr.real_length > 0 for tr in rec):
                for tr, d in zip(rec, pr):
                    ln = max(ln, l.oracle_get_length(d.ctypes.data, tr.first_sample, tr.real_length, d.shape[0], thr))
            ln = max(1, ln)
            oc.append([(d, tr.first_sample, ln, float(l.oracle_t60(d.ctypes.data, tr.first_sample, ln))) for tr, d in zip(rec, pr)])
        out.append(oc)
    return float(mx), out
