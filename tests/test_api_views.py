"""CPU test of the Python side of the render call (ear_b200/api.py::_render_call): the Track objects it returns are VIEWS of
the library-owned result (no copy), and the result is freed exactly once, when the last view is gone.

The C ABI behind it here is tests/host_emul/libabi_on_oracle.so -- the test-only implementation of the ABI slice on the
CPU oracle that tests/test_host_surface.py preloads into the CLIs -- loaded directly with ctypes.  The product library is not
involved (it needs a GPU); what is under test is the ownership logic of the binding."""
import ctypes as C
import gc
import os
import subprocess

import numpy as np
import pytest

from ear_b200 import api, scenes
from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_SRC = os.path.join(ROOT, "tests", "host_emul", "abi_on_oracle.cpp")
SHIM = os.path.join(ROOT, "tests", "host_emul", "libabi_on_oracle.so")


class _CountingLib:
    """Forwards to the shim and counts the calls of ear_b200_result_free."""

    def __init__(self, lib):
        self._lib = lib
        self.freed = 0

    def __getattr__(self, name):
        return getattr(self._lib, name)

    def ear_b200_result_free(self, res):
        self.freed += 1
        self._lib.ear_b200_result_free(res)


@pytest.fixture(scope="module")
def shim_lib(oracle_lib):
    deps = [SHIM_SRC, os.path.join(ROOT, "include", "ear_b200.h")]
    if not os.path.exists(SHIM) or any(os.path.getmtime(d) > os.path.getmtime(SHIM) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SHIM, SHIM_SRC, "-L", os.path.join(ROOT, "oracle"),
                        "-lear_oracle", "-Wl,-rpath,$ORIGIN/../../oracle"], check=True)
    lib = C.CDLL(SHIM)
    vp, i32 = C.c_void_p, C.c_int32
    lib.ear_b200_last_error.restype = C.c_char_p
    lib.ear_b200_scene_create.argtypes = [vp, vp, i32, vp, i32, i32, i32, C.POINTER(vp)]
    lib.ear_b200_scene_destroy.argtypes = [vp]
    lib.ear_b200_scene_destroy.restype = None
    lib.ear_b200_render.argtypes = [vp, C.POINTER(api.ContextC), i32, C.POINTER(api.RecorderC), i32, C.POINTER(api.OptionsC),
                                    C.POINTER(C.POINTER(api.ResultC))]
    lib.ear_b200_result_free.argtypes = [C.POINTER(api.ResultC)]
    lib.ear_b200_result_free.restype = None
    return lib


def test_tracks_are_views_of_the_result_and_free_it_with_the_last_view(shim_lib):
    sc = scenes.example1_scene(samples=4000, stereo=True)
    ctxs, recs = api.contexts_from_def(sc)
    verts = np.ascontiguousarray(sc.triangles(), np.float32)
    mats = np.ascontiguousarray(sc.triangle_materials(), np.int32)
    table = np.ascontiguousarray(sc.material_table(), np.float32)
    h = C.c_void_p()
    assert shim_lib.ear_b200_scene_create(verts.ctypes.data, mats.ctypes.data, verts.shape[0], table.ctypes.data, table.shape[0],
                                          table.shape[1], 0, C.byref(h)) == 0
    lib = _CountingLib(shim_lib)
    try:
        res = api._render_call(lib, lib.ear_b200_render, h, ctxs, recs, 30, 5, 0, 0, -1, True, None)
        want, counters = ob.OracleScene.from_def(sc).render(ctxs, recs, max_bounces=30, seed=5)
        assert (res.rays, res.segments, res.contributions) == (counters["rays"], counters["segments"], counters["contributions"])
        flat = [t for c in res.tracks for r in c for t in r]
        flat_want = [t for c in want for r in c for t in r]
        assert len(flat) == len(flat_want) == 2 * len(ctxs)
        for a in flat:
            assert not a.data.flags.owndata          # a view of library memory, not a copy
        keep = flat[1].data                          # one view outlives the result object
        expect = flat_want[1].data.copy()
        first, real = flat[1].first_sample, flat[1].real_length
        assert (first, real) == (flat_want[1].first_sample, flat_want[1].real_length)
        del res, flat, a
        gc.collect()
        assert lib.freed == 0                        # still referenced through `keep`
        junk = [np.full(keep.shape[0], 7.0, np.float32) for _ in range(8)]   # would land in the block if it had been freed
        assert np.array_equal(keep, expect)
        window = keep[first:real + 1]                # a slice keeps the result alive just the same
        del keep, junk
        gc.collect()
        assert lib.freed == 0
        assert np.array_equal(window, expect[first:real + 1])
        del window
        gc.collect()
        assert lib.freed == 1                        # freed once, with the last view
    finally:
        shim_lib.ear_b200_scene_destroy(h)


def test_every_result_is_freed_once(shim_lib):
    sc = scenes.rt60_scene(samples=2000)
    ctxs, recs = api.contexts_from_def(sc)
    verts = np.ascontiguousarray(sc.triangles(), np.float32)
    mats = np.ascontiguousarray(sc.triangle_materials(), np.int32)
    table = np.ascontiguousarray(sc.material_table(), np.float32)
    h = C.c_void_p()
    assert shim_lib.ear_b200_scene_create(verts.ctypes.data, mats.ctypes.data, verts.shape[0], table.ctypes.data, table.shape[0],
                                          table.shape[1], 0, C.byref(h)) == 0
    lib = _CountingLib(shim_lib)
    try:
        for k in range(5):
            res = api._render_call(lib, lib.ear_b200_render, h, ctxs, recs, 20, 3 + k, 0, 0, -1, True, None)
            assert res.rays > 0
            res = None
            gc.collect()
            assert lib.freed == k + 1
    finally:
        shim_lib.ear_b200_scene_destroy(h)
