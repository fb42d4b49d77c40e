"""ctypes binding of libear_b200.so (include/ear_b200.h) and the host-side mirror of the
reference's render seam.

The reference has no plugin API; its seam is `Scene::Render(band, sound, absorbtion_factor,
num_samples, dry, recorders, keyframeID)` (src/Scene.h:77) driven once per SceneContext
(src/SceneContext.h:26-48, src/EAR.cpp:170-207).  `Scene.render(contexts)` here takes the same
arguments per context and hands the whole fan-out to the GPU in one call.

There is NO CPU fallback: if the CUDA library is missing or no device is present, loading or the
first compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

MAX_BANDS = 8
SAMPLE_RATE = 44100
MONO, STEREO = 1, 2
POINT_SOURCE, MESH_SOURCE = 0, 1
FIRST_SAMPLE_INIT = 3 * SAMPLE_RATE - 1   # FloatBuffer ctor, src/Recorder.cpp:43-48

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EAR_B200_LIB") or os.path.join(_HERE, "csrc", "libear_b200.so")   # env: A/B builds only


class RecorderC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("position", C.c_float * 3), ("right_ear", C.c_float * 3),
                ("head_size", C.c_float), ("head_absorption", C.c_float * MAX_BANDS)]


class ContextC(C.Structure):
    _fields_ = [("band", C.c_int32), ("stream_id", C.c_int32), ("num_samples", C.c_int64),
                ("absorption_factor", C.c_float), ("dry_level", C.c_float), ("gain", C.c_float),
                ("source_position", C.c_float * 3), ("source_kind", C.c_int32), ("emitter_first", C.c_int32),
                ("emitter_count", C.c_int32), ("reserved", C.c_int32)]


class OptionsC(C.Structure):
    _fields_ = [("max_bounces", C.c_int32), ("n_bins", C.c_int32), ("seed", C.c_uint64),
                ("first_ray", C.c_int64), ("ray_count", C.c_int64), ("finalise", C.c_int32),
                ("reserved", C.c_int32), ("post_exponent", C.c_float), ("post_divisor", C.c_float)]


class TrackC(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("first_sample", C.c_uint32), ("real_length", C.c_uint32),
                ("length", C.c_uint32), ("reserved", C.c_uint32)]


class ResultC(C.Structure):
    _fields_ = [("n_contexts", C.c_int32), ("n_recorders", C.c_int32), ("tracks", C.POINTER(TrackC)),
                ("rays", C.c_uint64), ("segments", C.c_uint64), ("occlusion_queries", C.c_uint64),
                ("contributions", C.c_uint64), ("bin_updates", C.c_uint64), ("dropped_updates", C.c_uint64),
                ("device_ms", C.c_double), ("bvh_build_ms", C.c_double), ("t60", C.POINTER(C.c_float)),
                ("maximum", C.c_float), ("reserved", C.c_float)]


class StatsC(C.Structure):
    _fields_ = [("launches", C.c_uint64 * 8), ("ms", C.c_double * 8), ("iterations", C.c_uint64), ("reserved", C.c_uint64)]


KERNEL_CLASSES = ["shade", "closest", "anyhit", "splat", "sort", "finalise", "vismap", "mapbuild"]

EXPORTS = [
    "ear_b200_last_error", "ear_b200_abi_version", "ear_b200_device_count", "ear_b200_scene_create",
    "ear_b200_scene_destroy", "ear_b200_first_hit", "ear_b200_occluded", "ear_b200_trace_paths",
    "ear_b200_render", "ear_b200_result_free", "ear_b200_trace_device", "ear_b200_finalise_device",
    "ear_b200_default_bins", "ear_b200_scene_stats", "ear_b200_scene_stats_reset", "ear_b200_convolve",
    "ear_b200_scene_image_size", "ear_b200_scene_image_write", "ear_b200_scene_create_from_image", "ear_b200_scene_clone",
    "ear_b200_post_power_device", "ear_b200_post_truncate_device", "ear_b200_tracks_per_recorder",
    "ear_b200_scene_set_emitters", "ear_b200_group_create", "ear_b200_group_destroy", "ear_b200_group_size",
    "ear_b200_group_render", "ear_b200_release_cached_memory", "ear_b200_convolve_fft",
]

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Loads libear_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(ear_b200 has no CPU fallback)")
    lib = C.CDLL(p)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.ear_b200_last_error.restype = C.c_char_p
    lib.ear_b200_abi_version.restype = i32
    lib.ear_b200_device_count.restype = i32
    lib.ear_b200_scene_create.argtypes = [vp, vp, i32, vp, i32, i32, i32, C.POINTER(vp)]
    lib.ear_b200_scene_destroy.argtypes = [vp]
    lib.ear_b200_scene_destroy.restype = None
    lib.ear_b200_scene_image_size.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.ear_b200_scene_image_write.argtypes = [vp, vp, C.c_uint64]
    lib.ear_b200_scene_create_from_image.argtypes = [vp, C.c_uint64, i32, C.POINTER(vp)]
    lib.ear_b200_scene_clone.argtypes = [vp, i32, C.POINTER(vp)]
    lib.ear_b200_post_power_device.argtypes = [vp, C.POINTER(RecorderC), i32, i32, i32, vp, vp, C.c_float, C.POINTER(C.c_float), vp, vp]
    lib.ear_b200_post_truncate_device.argtypes = [vp, C.POINTER(RecorderC), i32, i32, i32, vp, vp, C.c_float, vp, vp]
    lib.ear_b200_first_hit.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.ear_b200_occluded.argtypes = [vp, vp, vp, i64, vp]
    lib.ear_b200_trace_paths.argtypes = [vp, C.POINTER(ContextC), i32, C.POINTER(OptionsC), i64, vp, vp]
    lib.ear_b200_render.argtypes = [vp, C.POINTER(ContextC), i32, C.POINTER(RecorderC), i32, C.POINTER(OptionsC),
                                    C.POINTER(C.POINTER(ResultC))]
    lib.ear_b200_group_create.argtypes = [vp, C.POINTER(i32), i32, C.POINTER(vp)]
    lib.ear_b200_group_destroy.argtypes = [vp]
    lib.ear_b200_group_destroy.restype = None
    lib.ear_b200_group_size.argtypes = [vp]
    lib.ear_b200_group_render.argtypes = [vp, C.POINTER(ContextC), i32, C.POINTER(RecorderC), i32, C.POINTER(OptionsC),
                                          C.POINTER(C.POINTER(ResultC))]
    lib.ear_b200_result_free.argtypes = [C.POINTER(ResultC)]
    lib.ear_b200_result_free.restype = None
    lib.ear_b200_trace_device.argtypes = [vp, C.POINTER(ContextC), i32, C.POINTER(RecorderC), i32,
                                          C.POINTER(OptionsC), i32, vp, vp, vp, vp]
    lib.ear_b200_finalise_device.argtypes = [vp, C.POINTER(ContextC), i32, C.POINTER(RecorderC), i32, i32, vp, vp, vp]
    lib.ear_b200_default_bins.argtypes = [vp, C.POINTER(OptionsC)]
    lib.ear_b200_tracks_per_recorder.argtypes = [C.POINTER(RecorderC), i32]
    lib.ear_b200_tracks_per_recorder.restype = i32
    lib.ear_b200_scene_set_emitters.argtypes = [vp, vp, i32]
    lib.ear_b200_scene_stats.argtypes = [vp, C.POINTER(StatsC)]
    lib.ear_b200_scene_stats_reset.argtypes = [vp]
    lib.ear_b200_scene_stats_reset.restype = None
    u32 = C.c_uint32
    lib.ear_b200_convolve.argtypes = [i32, vp, u32, u32, u32, vp, u32, u32, u32, vp, u32, u32, vp, u32, C.POINTER(u32), C.POINTER(u32)]
    lib.ear_b200_convolve_fft.argtypes = lib.ear_b200_convolve.argtypes
    if path is None:
        _lib = lib
    return lib


class EarError(RuntimeError):
    """Raised with the text of ear_b200_last_error(); the C++ host prints it as `Error: <what>`."""


def _check(lib, rc):
    if rc != 0:
        raise EarError(lib.ear_b200_last_error().decode("utf8", "replace"))


@dataclass
class Context:
    """One SceneContext (src/SceneContext.h:28-45)."""
    band: int
    num_samples: int
    absorption_factor: float
    source_position: Sequence[float]
    dry_level: float = 1.0
    gain: float = 1.0
    stream_id: int = 0      # 0: position in the call keys the random streams; k > 0: key k - 1 (see include/ear_b200.h)
    emitter: Optional[Sequence[int]] = None   # mesh source: (first, count) into the scene's emitter-triangle table

    def to_c(self) -> ContextC:
        c = ContextC()
        c.band, c.num_samples = int(self.band), int(self.num_samples)
        c.stream_id = int(self.stream_id)
        c.absorption_factor = float(np.float32(self.absorption_factor))
        c.dry_level, c.gain = float(self.dry_level), float(self.gain)
        c.source_position[:] = [float(x) for x in self.source_position]
        if self.emitter is not None:
            c.source_kind, c.emitter_first, c.emitter_count = MESH_SOURCE, int(self.emitter[0]), int(self.emitter[1])
        return c


@dataclass
class Recorder:
    """Mono/StereoRecorder parameters evaluated at one keyframe."""
    position: Sequence[float]
    stereo: bool = False
    right_ear: Sequence[float] = (-1.0, 0.0, 0.0)
    head_size: float = 0.2
    head_absorption: Sequence[float] = (0.1, 0.3, 0.9)   # as written in the .ear file

    def to_c(self) -> RecorderC:
        r = RecorderC()
        r.kind = STEREO if self.stereo else MONO
        r.position[:] = [float(x) for x in self.position]
        r.right_ear[:] = [float(x) for x in self.right_ear]
        r.head_size = float(self.head_size)
        ha = list(self.head_absorption) + [self.head_absorption[-1]] * (MAX_BANDS - len(self.head_absorption))
        for i in range(MAX_BANDS):
            # StereoRecorder ctor, src/StereoRecorder.cpp:55-57: max(0, powf(1 - a, 4)) in float32
            base = np.float32(1.0) - np.float32(ha[i])
            r.head_absorption[i] = float(max(np.float32(0.0), np.float32(powf4(base))))
        return r


_libm = C.CDLL("libm.so.6")
_libm.powf.argtypes = [C.c_float, C.c_float]
_libm.powf.restype = C.c_float


def powf4(x: np.float32) -> np.float32:
    """powf(x, 4) through the same libm the C++ host uses (src/StereoRecorder.cpp:56)."""
    return np.float32(_libm.powf(float(x), 4.0))


def make_options(max_bounces=1000, n_bins=0, seed=1, first_ray=0, ray_count=-1, finalise=True, post=None) -> OptionsC:
    """post=(exponent, divisor): run Render()'s post chain on the device before the download (the reference: 0.335, 256)."""
    o = OptionsC()
    o.max_bounces, o.n_bins, o.seed = int(max_bounces), int(n_bins), int(seed)
    o.first_ray, o.ray_count, o.finalise = int(first_ray), int(ray_count), 1 if finalise else 0
    if post is not None:
        o.post_exponent, o.post_divisor = float(post[0]), float(post[1])
    return o


def tracks_per_recorder(rec_c) -> int:
    """Tracks each recorder owns in the device buffers of a call (2 if any recorder of the call is stereo, else 1)."""
    return 2 if any(r.kind == STEREO for r in rec_c) else 1


def pack_contexts(contexts: Sequence[Context]):
    arr = (ContextC * len(contexts))()
    for i, c in enumerate(contexts):
        arr[i] = c.to_c()
    return arr


def pack_recorders(recorders, n_contexts: int):
    """recorders: either one list (same for every context) or a list per context ([n_ctx][R])."""
    if recorders and isinstance(recorders[0], Recorder):
        per_ctx = [recorders] * n_contexts
    else:
        per_ctx = recorders
    n_rec = len(per_ctx[0])
    arr = (RecorderC * (n_contexts * n_rec))()
    for c in range(n_contexts):
        assert len(per_ctx[c]) == n_rec
        for r in range(n_rec):
            arr[c * n_rec + r] = per_ctx[c][r].to_c()
    return arr, n_rec


@dataclass
class Track:
    """FloatBuffer view: data[length], first_sample, real_length (src/Recorder.h:55-65)."""
    data: np.ndarray
    first_sample: int
    real_length: int


@dataclass
class RenderResult:
    tracks: List[List[List[Track]]]    # [context][recorder][track]
    rays: int
    segments: int
    occlusion_queries: int
    contributions: int
    bin_updates: int
    dropped_updates: int
    device_ms: float
    maximum: float = 0.0               # set by the device post chain (sharding.render_sharded(post=...))
    t60: Optional[list] = None         # [context][recorder][track]


def convolve(response: "Track", dry: np.ndarray, offset: int = 0, response2: Optional["Track"] = None, device: int = 0,
             fft: bool = False) -> "Track":
    """RecorderTrack::Process (src/Recorder.cpp:247-292) on the GPU: direct convolution of a dry signal with a
    response track (cross-faded into `response2` for keyframed scenes).  Returns the result track.
    fft=True: the frequency-domain form (src/Recorder.cpp:145-243), equal up to float32 FFT rounding."""
    lib = load_library()
    dry = np.ascontiguousarray(dry, np.float32)
    r1 = np.ascontiguousarray(response.data, np.float32)
    first, length = response.first_sample, response.real_length
    r2 = None
    if response2 is not None:
        r2 = np.ascontiguousarray(response2.data, np.float32)
        first, length = min(first, response2.first_sample), max(length, response2.real_length)
    out_len = max(3 * SAMPLE_RATE, dry.shape[0] + offset + length)
    out = np.zeros(out_len, np.float32)
    of, orl = C.c_uint32(), C.c_uint32()
    fn = lib.ear_b200_convolve_fft if fft else lib.ear_b200_convolve
    _check(lib, fn(
        device, r1.ctypes.data, r1.shape[0], response.first_sample, response.real_length,
        r2.ctypes.data if r2 is not None else None, r2.shape[0] if r2 is not None else 0,
        response2.first_sample if response2 is not None else 0, response2.real_length if response2 is not None else 0,
        dry.ctypes.data, dry.shape[0], offset, out.ctypes.data, out_len, C.byref(of), C.byref(orl)))
    return Track(out, int(of.value), int(orl.value))


class Scene:
    """Triangle soup + BVH + material table resident on one GPU (ear_b200_scene)."""

    def __init__(self, verts: np.ndarray, tri_material: np.ndarray, materials: np.ndarray, device: int = 0):
        self.lib = load_library()
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3, 3)
        self.tri_material = np.ascontiguousarray(tri_material, np.int32)
        self.materials = np.ascontiguousarray(materials, np.float32)
        assert self.materials.ndim == 3 and self.materials.shape[2] == 4
        self.n_bands = self.materials.shape[1]
        self.device = device
        h = C.c_void_p()
        _check(self.lib, self.lib.ear_b200_scene_create(
            self.verts.ctypes.data, self.tri_material.ctypes.data, self.verts.shape[0], self.materials.ctypes.data,
            self.materials.shape[0], self.n_bands, device, C.byref(h)))
        self.handle = h

    @classmethod
    def from_def(cls, scene_def, materials: Optional[np.ndarray] = None, device: int = 0) -> "Scene":
        tab = scene_def.material_table() if materials is None else materials
        scene = cls(scene_def.triangles(), scene_def.triangle_materials(), tab, device)
        em = emitter_table(scene_def)
        if em is not None:
            scene.set_emitters(em)
        return scene

    @classmethod
    def from_image(cls, image_ptr: int, image_bytes: int, n_bands: int, device: int = 0) -> "Scene":
        """Adopts a scene image resident in device memory (ear_b200_scene_create_from_image)."""
        self = cls.__new__(cls)
        self.lib = load_library()
        self.verts = self.tri_material = self.materials = None
        self.n_bands = n_bands
        self.device = device
        h = C.c_void_p()
        _check(self.lib, self.lib.ear_b200_scene_create_from_image(C.c_void_p(image_ptr), image_bytes, device, C.byref(h)))
        self.handle = h
        return self

    def image_size(self) -> int:
        n = C.c_uint64()
        _check(self.lib, self.lib.ear_b200_scene_image_size(self.handle, C.byref(n)))
        return int(n.value)

    def image_write(self, dst_ptr: int, dst_bytes: int) -> None:
        """Copies the scene image into device memory at dst_ptr (e.g. a torch uint8 tensor's data_ptr())."""
        _check(self.lib, self.lib.ear_b200_scene_image_write(self.handle, C.c_void_p(dst_ptr), dst_bytes))

    def set_emitters(self, verts: np.ndarray) -> None:
        """Emitter triangles of the scene's mesh sources, all of them concatenated ([n][3][3])."""
        v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3, 3)
        _check(self.lib, self.lib.ear_b200_scene_set_emitters(self.handle, v.ctypes.data, v.shape[0]))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ear_b200_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def first_hit(self, origins: np.ndarray, dirs: np.ndarray):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        idx = np.empty(o.shape[0], np.int32)
        t = np.empty(o.shape[0], np.float32)
        _check(self.lib, self.lib.ear_b200_first_hit(self.handle, o.ctypes.data, d.ctypes.data, o.shape[0],
                                                      idx.ctypes.data, t.ctypes.data))
        return idx, t

    def occluded(self, p: np.ndarray, x: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 3)
        out = np.empty(p.shape[0], np.uint8)
        _check(self.lib, self.lib.ear_b200_occluded(self.handle, p.ctypes.data, x.ctypes.data, p.shape[0],
                                                     out.ctypes.data))
        return out

    def trace_paths(self, context: Context, ctx_index: int, n: int, max_bounces: int, seed: int, first_ray: int = 0):
        opt = make_options(max_bounces=max_bounces, seed=seed, first_ray=first_ray, ray_count=n)
        hits = np.empty((n, max_bounces), np.int32)
        state = np.empty((n, 8), np.float32)
        cc = context.to_c()
        _check(self.lib, self.lib.ear_b200_trace_paths(self.handle, C.byref(cc), ctx_index, C.byref(opt), n,
                                                        hits.ctypes.data, state.ctypes.data))
        return hits, state

    def stats(self, reset: bool = False) -> dict:
        """Launch counts and CUDA-event milliseconds per kernel class since the last reset."""
        st = StatsC()
        _check(self.lib, self.lib.ear_b200_scene_stats(self.handle, C.byref(st)))
        out = {"iterations": int(st.iterations), "launches": {}, "ms": {}}
        for i, name in enumerate(KERNEL_CLASSES):
            out["launches"][name] = int(st.launches[i])
            out["ms"][name] = float(st.ms[i])
        if reset:
            self.lib.ear_b200_scene_stats_reset(self.handle)
        return out

    def default_bins(self, opt: OptionsC) -> int:
        return int(self.lib.ear_b200_default_bins(self.handle, C.byref(opt)))

    def render(self, contexts: Sequence[Context], recorders, max_bounces: int = 1000, seed: int = 1, n_bins: int = 0,
               first_ray: int = 0, ray_count: int = -1, finalise: bool = True, post=None) -> RenderResult:
        return _render_call(self.lib, self.lib.ear_b200_render, self.handle, contexts, recorders, max_bounces, seed, n_bins,
                            first_ray, ray_count, finalise, post)


class _ResultOwner:
    """Keeps an ear_b200_result alive while numpy views of its tracks exist; frees it with the last of them."""

    def __init__(self, lib, res):
        self.lib, self.res = lib, res

    def __del__(self):
        try:
            if self.res is not None:
                self.lib.ear_b200_result_free(self.res)
                self.res = None
        except Exception:
            pass


def _render_call(lib, fn, handle, contexts, recorders, max_bounces, seed, n_bins, first_ray, ray_count, finalise, post):
    import time
    t0 = time.perf_counter()
    ctx = pack_contexts(contexts)
    rec, n_rec = pack_recorders(recorders, len(contexts))
    opt = make_options(max_bounces, n_bins, seed, first_ray, ray_count, finalise, post)
    res = C.POINTER(ResultC)()
    t1 = time.perf_counter()
    _check(lib, fn(handle, ctx, len(contexts), rec, n_rec, C.byref(opt), C.byref(res)))
    t2 = time.perf_counter()
    owner = _ResultOwner(lib, res)     # the tracks below are views of the library's (page-locked) block: no copy
    r = res.contents
    tracks, t60 = [], []
    for c in range(r.n_contexts):
        per_rec, per_t60 = [], []
        for k in range(r.n_recorders):
            pair, pair_t60 = [], []
            n_tracks = 2 if rec[c * n_rec + k].kind == STEREO else 1
            for tr in range(n_tracks):
                t = r.tracks[(c * r.n_recorders + k) * 2 + tr]
                buf = (C.c_float * t.length).from_address(C.cast(t.data, C.c_void_p).value or 0) if t.length else (C.c_float * 0)()
                buf._owner = owner             # the result is freed when the last track view is gone
                pair.append(Track(np.frombuffer(buf, np.float32), int(t.first_sample), int(t.real_length)))
                if r.t60:
                    pair_t60.append(float(r.t60[(c * r.n_recorders + k) * 2 + tr]))
            per_rec.append(pair)
            per_t60.append(pair_t60)
        tracks.append(per_rec)
        t60.append(per_t60)
    out = RenderResult(tracks, int(r.rays), int(r.segments), int(r.occlusion_queries), int(r.contributions),
                       int(r.bin_updates), int(r.dropped_updates), float(r.device_ms))
    if r.t60:
        out.maximum, out.t60 = float(r.maximum), t60
    if os.environ.get("EAR_B200_DEBUG"):
        t3 = time.perf_counter()
        print(f"[ear_b200.api] render: pack {1e3 * (t1 - t0):.1f} ms, library call {1e3 * (t2 - t1):.1f} ms, "
              f"track views {1e3 * (t3 - t2):.1f} ms", file=sys.stderr)
    return out


class Group:
    """`scene` replicated on several GPUs of this process (ear_b200_group): render() shards the ray ids of every context
    over the GPUs and reduces the partial histograms onto the scene's own GPU through peer memory."""

    def __init__(self, scene: Scene, devices: Sequence[int]):
        self.scene, self.lib = scene, scene.lib
        arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        _check(self.lib, self.lib.ear_b200_group_create(scene.handle, arr, len(devices), C.byref(h)))
        self.handle = h

    def render(self, contexts, recorders, max_bounces: int = 1000, seed: int = 1, n_bins: int = 0, first_ray: int = 0,
               ray_count: int = -1, finalise: bool = True, post=None) -> RenderResult:
        return _render_call(self.lib, self.lib.ear_b200_group_render, self.handle, contexts, recorders, max_bounces, seed, n_bins,
                            first_ray, ray_count, finalise, post)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ear_b200_group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def emitter_table(scene_def) -> Optional[np.ndarray]:
    """Emitter triangles of all mesh sources, concatenated in file order ([E][3][3]) -- what Scene.set_emitters takes."""
    parts = [np.asarray(s.mesh_verts, np.float32).reshape(-1, 3, 3) for s in scene_def.sources if getattr(s, "mesh_verts", None) is not None]
    return np.concatenate(parts) if parts else None


def contexts_from_def(scene_def, t60_only: bool = False, n_bands: int = 3, air_factors=None):
    """The SceneContext list `Render()` builds (src/EAR.cpp:170-191): sound x keyframe x band, with
    rays = samples // 10 (src/EAR.cpp:81) and absorption_factor = 1 - air[band] in float32.
    Returns (contexts, recorders_per_context)."""
    contexts, recs = [], []
    rays = int(scene_def.samples) // 10
    keys = scene_def.keys
    emit_at = 0   # mesh sources: position of their triangles in the concatenated emitter table (emitter_table())
    for src in scene_def.sources:
        emitter = None
        if getattr(src, "mesh_verts", None) is not None:
            n_e = int(np.asarray(src.mesh_verts).reshape(-1, 3, 3).shape[0])
            emitter = (emit_at, n_e)
            emit_at += n_e
        kfs = range(len(keys)) if keys is not None else [-1]
        for kf in kfs:
            for band in range(n_bands):
                if t60_only and band != 1:
                    continue
                if air_factors is not None:
                    af = np.float32(air_factors[band])
                else:
                    af = np.float32(1.0) - np.float32(scene_def.air_absorption[band])
                pos = src.animation[kf] if (kf >= 0 and src.animation is not None) else src.position
                if emitter is not None:
                    pos = (0.0, 0.0, 0.0)   # a mesh source has no location (src/SoundFile.cpp:50-56)
                contexts.append(Context(band, rays, float(af), [float(x) for x in pos], scene_def.drylevel, src.gain, 0, emitter))
                rr = []
                for rec in scene_def.recorders:
                    rpos = rec.animation[kf] if (kf >= 0 and rec.animation is not None) else rec.position
                    ear = (rec.right_ear_animation[kf] if (kf >= 0 and rec.right_ear_animation is not None)
                           else rec.right_ear)
                    rr.append(Recorder([float(x) for x in rpos], rec.stereo, [float(x) for x in ear], rec.head_size,
                                       list(rec.head_absorption)))
                recs.append(rr)
            if t60_only:
                break
        if t60_only:
            break
    return contexts, recs
