"""`.ear` scene files: writer, reader and the flat scene description handed to the C ABI.

The on-disk grammar is the one the reference's Blender exporter emits
(blender/render_EAR/__init__.py:154-196 `pack`/`writeblock`) and its C++ reader
consumes (src/Datatype.cpp:25-147, src/EAR.cpp:92-119):

  file      := ".EAR" block*
  primitive := "int4" i32 | "flt4" f32 | "vec3" flt4 flt4 flt4 | "tri " vec3 vec3 vec3
             | "str " utf8 NUL-padded to a multiple of 4 (always at least one NUL)
  block     := 4-byte id, i32 payload length, payload

Triangle index == position in the concatenation of all MESH blocks in file order
(src/Scene.cpp:103-106, src/Mesh.cpp:116-123); everything downstream relies on that.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

SAMPLE_RATE = 44100

_TRI_DTYPE = np.dtype(
    [("tag", "S4")]
    + [(f"v{v}{name}", fmt) for v in range(3) for name, fmt in
       (("tag", "S4"), ("xt", "S4"), ("x", "<f4"), ("yt", "S4"), ("y", "<f4"), ("zt", "S4"), ("z", "<f4"))]
)
assert _TRI_DTYPE.itemsize == 88


def pack_int(v: int) -> bytes:
    return b"int4" + struct.pack("<i", v)


def pack_float(v: float) -> bytes:
    return b"flt4" + struct.pack("<f", v)


def pack_vec3(v: Sequence[float]) -> bytes:
    return b"vec3" + b"".join(pack_float(float(x)) for x in v)


def pack_str(s: str) -> bytes:
    e = s.encode("utf8")
    return b"str " + e + b"\x00" * (4 - len(e) % 4)


def pack_tris(verts: np.ndarray) -> bytes:
    """verts: [T,3,3] float32 -> T 'tri ' records (88 bytes each), vectorised."""
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3, 3)
    rec = np.empty(verts.shape[0], dtype=_TRI_DTYPE)
    rec["tag"] = b"tri "
    for v in range(3):
        rec[f"v{v}tag"] = b"vec3"
        for c, name in enumerate("xyz"):
            rec[f"v{v}{name}t"] = b"flt4"
            rec[f"v{v}{name}"] = verts[:, v, c]
    return rec.tobytes()


def block(block_id: str, payload: bytes) -> bytes:
    bid = block_id.encode("ascii")
    assert len(bid) == 4
    return bid + struct.pack("<i", len(payload)) + payload


@dataclass
class MaterialDef:
    name: str
    refl: Sequence[float]
    refr: Sequence[float] = (0.0, 0.0, 0.0)
    spec: Sequence[float] = (0.0, 0.0, 0.0)


@dataclass
class MeshDef:
    material: str
    verts: np.ndarray  # [T,3,3] float32


@dataclass
class SourceDef:
    wavs: Sequence[str]                 # 1 path -> SSRC, 3 paths -> 3SRC
    position: Optional[Sequence[float]] = None
    animation: Optional[np.ndarray] = None  # [K,3]
    gain: float = 1.0
    offset: float = 0.0
    # mesh source (src/SoundFile.cpp:50-53): instead of a location the block carries a `mesh` sub-block,
    # [str material name][tri ...] -- the name must be a defined material (Mesh's constructor looks it up)
    mesh_material: Optional[str] = None
    mesh_verts: Optional[np.ndarray] = None  # [E,3,3] float32


@dataclass
class RecorderDef:
    filename: str
    position: Optional[Sequence[float]] = None
    animation: Optional[np.ndarray] = None
    stereo: bool = False
    right_ear: Sequence[float] = (-1.0, 0.0, 0.0)
    right_ear_animation: Optional[np.ndarray] = None
    head_size: float = 0.2
    head_absorption: Sequence[float] = (0.1, 0.3, 0.9)


@dataclass
class SceneDef:
    """Everything an `.ear` file carries (exporter order: VRSN, MAT*, SET, KEYS?, FREQ, objects)."""
    materials: List[MaterialDef] = field(default_factory=list)
    meshes: List[MeshDef] = field(default_factory=list)
    sources: List[SourceDef] = field(default_factory=list)
    recorders: List[RecorderDef] = field(default_factory=list)
    air_absorption: Sequence[float] = (0.0, 0.0, 0.0)
    drylevel: float = 1.0
    samples: int = 1000000            # file value; rays per context = samples // 10 (src/EAR.cpp:81)
    maxthreads: int = 0
    keys: Optional[Sequence[float]] = None
    freq: Sequence[float] = (0.3, 2.0, 6.0)
    debugdir: Optional[str] = None

    # ---- flat arrays for the C ABI (triangle order = file order) ----
    def triangles(self) -> np.ndarray:
        if not self.meshes:
            return np.zeros((0, 3, 3), np.float32)
        return np.ascontiguousarray(np.concatenate([m.verts.reshape(-1, 3, 3) for m in self.meshes]), np.float32)

    def triangle_materials(self) -> np.ndarray:
        names = [m.name for m in self.materials]
        out = [np.full(m.verts.reshape(-1, 3, 3).shape[0], names.index(m.material), np.int32) for m in self.meshes]
        return np.concatenate(out) if out else np.zeros(0, np.int32)

    def material_table(self) -> np.ndarray:
        """[M,3,4] float32 rows {refl, refr, kept, spec} with `kept` derived exactly as
        src/Material.cpp:33-60 does in float32: a=1; a-=(refl-1e-9f); a-=(refr-1e-9f); kept=1-a."""
        tab = np.zeros((len(self.materials), 3, 4), np.float32)
        eps = np.float32(1e-9)
        one = np.float32(1.0)
        for i, m in enumerate(self.materials):
            for b in range(3):
                refl = np.float32(m.refl[b])
                refr = np.float32(m.refr[b])
                a = one
                a = np.float32(a - np.float32(refl - eps))
                a = np.float32(a - np.float32(refr - eps))
                if a < 0:
                    raise ValueError("Invalid material settings")
                tab[i, b] = (refl, refr, np.float32(one - a), np.float32(m.spec[b]))
        return tab

    def to_bytes(self) -> bytes:
        out = [b".EAR", block("VRSN", pack_int(0))]
        for m in self.materials:
            p = pack_str(m.name)
            for arr in (m.refl, m.refr, m.spec):
                p += b"".join(pack_float(float(x)) for x in arr)
            out.append(block("MAT ", p))
        s = pack_str("debug") + pack_int(0)
        s += pack_str("absorption") + pack_vec3(self.air_absorption)
        s += pack_str("drylevel") + pack_float(self.drylevel)
        s += pack_str("samples") + pack_int(int(self.samples))
        s += pack_str("maxthreads") + pack_int(int(self.maxthreads))
        if self.debugdir:
            s += pack_str("debugdir") + pack_str(self.debugdir)
        out.append(block("SET ", s))
        if self.keys is not None:
            out.append(block("KEYS", b"".join(pack_float(float(k)) for k in self.keys)))
        out.append(block("FREQ", b"".join(pack_float(float(f)) for f in self.freq)))
        for mesh in self.meshes:
            out.append(block("MESH", pack_str(mesh.material) + pack_tris(mesh.verts)))

        def loc(position, animation):
            if animation is not None:
                return block("anim", b"".join(pack_vec3(v) for v in np.asarray(animation, np.float32)))
            return pack_vec3(position)

        for src in self.sources:
            p = b"".join(pack_str(w) for w in src.wavs)
            if src.mesh_verts is not None:
                p += block("mesh", pack_str(src.mesh_material) + pack_tris(np.asarray(src.mesh_verts, np.float32)))
            else:
                p += loc(src.position, src.animation)
            p += pack_float(src.gain) + pack_float(src.offset)
            out.append(block("SSRC" if len(src.wavs) == 1 else "3SRC", p))
        for rec in self.recorders:
            p = pack_str(rec.filename) + pack_float(35.0) + loc(rec.position, rec.animation)
            if rec.stereo:
                p += loc(rec.right_ear, rec.right_ear_animation)
                p += pack_float(rec.head_size) + pack_vec3(rec.head_absorption)
            out.append(block("OUT2" if rec.stereo else "OUT1", p))
        return b"".join(out)

    def write(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(self.to_bytes())


# ----------------------------------------------------------------------------------
# reader (mirrors what src/EAR.cpp:92-119 accepts; unknown blocks are skipped)
# ----------------------------------------------------------------------------------
class _Cursor:
    def __init__(self, data: bytes, pos: int = 0, end: Optional[int] = None):
        self.d, self.p, self.e = data, pos, len(data) if end is None else end

    def peek(self) -> bytes:
        return self.d[self.p:self.p + 4]

    def more(self) -> bool:
        return self.p < self.e

    def _tag(self, want: bytes):
        got = self.d[self.p:self.p + 4]
        if got != want:
            raise ValueError(f"Found {got!r} while expecting {want!r}")
        self.p += 4

    def i32(self) -> int:
        self._tag(b"int4")
        (v,) = struct.unpack_from("<i", self.d, self.p)
        self.p += 4
        return v

    def f32(self) -> float:
        self._tag(b"flt4")
        (v,) = struct.unpack_from("<f", self.d, self.p)
        self.p += 4
        return v

    def vec3(self):
        self._tag(b"vec3")
        return [self.f32(), self.f32(), self.f32()]

    def string(self) -> str:
        self._tag(b"str ")
        end = self.d.index(b"\x00", self.p)
        s = self.d[self.p:end].decode("utf8")
        n = end - self.p
        self.p += n + (4 - n % 4)
        return s

    def container(self):
        bid = self.d[self.p:self.p + 4]
        (n,) = struct.unpack_from("<i", self.d, self.p + 4)
        sub = _Cursor(self.d, self.p + 8, self.p + 8 + n)
        self.p += 8 + n
        return bid, sub

    def tris(self) -> np.ndarray:
        start = self.p
        n = 0
        while self.p + 88 <= self.e and self.d[self.p:self.p + 4] == b"tri ":
            self.p += 88
            n += 1
        rec = np.frombuffer(self.d, dtype=_TRI_DTYPE, count=n, offset=start)
        out = np.empty((n, 3, 3), np.float32)
        for v in range(3):
            for c, name in enumerate("xyz"):
                out[:, v, c] = rec[f"v{v}{name}"]
        return out

    def location(self):
        if self.peek() == b"anim":
            _, sub = self.container()
            frames = []
            while sub.more():
                frames.append(sub.vec3())
            return None, np.asarray(frames, np.float32)
        return self.vec3(), None


def read_ear(path: str) -> SceneDef:
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b".EAR":
        raise ValueError("Failed to read file")
    cur = _Cursor(data, 4)
    scene = SceneDef(samples=0)
    have_set = False
    while cur.more():
        bid, sub = cur.container()
        if bid == b"MAT ":
            name = sub.string()
            refl = [sub.f32() for _ in range(3)]
            refr = [sub.f32() for _ in range(3)] if sub.more() and sub.peek() == b"flt4" else [0.0] * 3
            spec = [sub.f32() for _ in range(3)] if sub.more() and sub.peek() == b"flt4" else [0.0] * 3
            scene.materials.append(MaterialDef(name, refl, refr, spec))
        elif bid == b"SET ":
            have_set = True
            while sub.more() and sub.peek() == b"str ":
                key = sub.string()
                tag = sub.peek()
                val = {b"int4": sub.i32, b"flt4": sub.f32, b"vec3": sub.vec3, b"str ": sub.string}[tag]()
                if key == "absorption":
                    scene.air_absorption = val
                elif key == "drylevel":
                    scene.drylevel = val
                elif key == "samples":
                    scene.samples = val
                elif key == "maxthreads":
                    scene.maxthreads = val
                elif key == "debugdir":
                    scene.debugdir = val
        elif bid == b"KEYS":
            keys = []
            while sub.more():
                keys.append(sub.f32())
            scene.keys = keys
        elif bid == b"FREQ":
            scene.freq = [sub.f32() for _ in range(3)]
        elif bid == b"MESH":
            mat = sub.string()
            scene.meshes.append(MeshDef(mat, sub.tris()))
        elif bid in (b"SSRC", b"3SRC"):
            wavs = [sub.string() for _ in range(1 if bid == b"SSRC" else 3)]
            pos = anim = mesh_mat = mesh_verts = None
            if sub.peek() == b"mesh":
                _, inner = sub.container()
                mesh_mat = inner.string()
                mesh_verts = inner.tris()
            else:
                pos, anim = sub.location()
            gain = sub.f32() if sub.more() and sub.peek() == b"flt4" else 1.0
            off = sub.f32() if sub.more() and sub.peek() == b"flt4" else 0.0
            scene.sources.append(SourceDef(wavs, pos, anim, gain, off, mesh_mat, mesh_verts))
        elif bid in (b"OUT1", b"OUT2"):
            fn = sub.string()
            sub.f32()
            pos, anim = sub.location()
            rec = RecorderDef(fn, pos, anim, stereo=(bid == b"OUT2"))
            if rec.stereo:
                ear, ear_anim = sub.location()
                if ear is not None:
                    rec.right_ear = ear
                rec.right_ear_animation = ear_anim
                rec.head_size = sub.f32()
                rec.head_absorption = sub.vec3()
            scene.recorders.append(rec)
        # VRSN and unknown ids (e.g. lower-case ssrc/3src mesh emitters): skipped, as in EAR.cpp:106-118
    if not have_set:
        raise ValueError("No settings block found in file")
    return scene
