"""TEST INFRASTRUCTURE: host build of the product's BVH builder + traversal headers (tests/host_emul)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "host_emul", "libemul.so")


def build():
    srcs = [os.path.join(HERE, "host_emul", "emul.cpp"), os.path.join(ROOT, "ear_b200", "csrc", "bvh_build.cpp")]
    deps = srcs + [os.path.join(ROOT, "ear_b200", "csrc", f) for f in ("traverse.cuh", "device_exact.cuh", "bvh_build.h", "vismap_geom.cuh")]
    deps.append(os.path.join(HERE, "host_emul", "cuda_shim.h"))
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-include",
                    os.path.join(HERE, "host_emul", "cuda_shim.h"), "-o", LIB, *srcs, "-lpthread"], check=True)


class EmulScene:
    def __init__(self, verts, mats=None, exact=False):
        build()
        self.l = C.CDLL(LIB)
        self.l.emul_create.restype = C.c_void_p
        self.l.emul_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        self.l.emul_destroy.argtypes = [C.c_void_p]
        self.l.emul_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        self.l.emul_first_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        self.l.emul_occluded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3, 3)
        self.h = self.l.emul_create(self.verts.ctypes.data, None, self.verts.shape[0])
        self.l.emul_set_exact.argtypes = [C.c_void_p, C.c_int32]
        if exact:
            self.l.emul_set_exact(self.h, 1)

    def __del__(self):
        if getattr(self, "h", None):
            self.l.emul_destroy(self.h)
            self.h = None

    def stats(self):
        n, d, ms = C.c_int32(), C.c_int32(), C.c_double()
        self.l.emul_stats(self.h, C.byref(n), C.byref(d), C.byref(ms))
        return n.value, d.value, ms.value

    def first_hit(self, o, d):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        idx = np.empty(o.shape[0], np.int32)
        t = np.empty(o.shape[0], np.float32)
        self.l.emul_first_hit(self.h, o.ctypes.data, d.ctypes.data, o.shape[0], idx.ctypes.data, t.ctypes.data)
        return idx, t

    def layout_hash(self) -> int:
        self.l.emul_hash.restype = C.c_uint64
        self.l.emul_hash.argtypes = [C.c_void_p]
        return int(self.l.emul_hash(self.h))

    def vismap_violations(self, x, res, points):
        """(violations, accepted, listed) of the visibility-map superset check for recorder position x."""
        x = np.ascontiguousarray(x, np.float32).reshape(3)
        p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        acc, lst = C.c_int64(), C.c_int64()
        self.l.emul_vismap_violations.restype = C.c_int64
        self.l.emul_vismap_violations.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        bad = self.l.emul_vismap_violations(self.h, x.ctypes.data, res, p.ctypes.data, p.shape[0], C.byref(acc), C.byref(lst))
        return int(bad), int(acc.value), int(lst.value)

    def occluded(self, p, x):
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 3)
        out = np.empty(p.shape[0], np.uint8)
        self.l.emul_occluded(self.h, p.ctypes.data, x.ctypes.data, p.shape[0], out.ctypes.data)
        return out
