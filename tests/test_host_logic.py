"""CPU-only checks: .ear round trip, ABI surface, BVH builder + traversal logic (host build of the
device headers, tests/host_emul) against the oracle, multi-rank sharding arithmetic over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from ear_b200 import api, earfile, scenes
from oracle import binding as ob
from tests import common, emul_binding as eb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ear_roundtrip_and_grammar(tmp_path):
    sc = scenes.example1_scene()
    sc.keys = [0.0, 0.5]
    sc.sources[0].animation = np.array([[0, 0, 1], [1, 0, 1]], np.float32)
    sc.sources[0].position = None
    sc.recorders[0].animation = np.array([[2, 0, 1], [3, 0, 1]], np.float32)
    sc.recorders[0].position = None
    raw = sc.to_bytes()
    assert raw[:4] == b".EAR" and raw[4:8] == b"VRSN"
    p = tmp_path / "a.ear"
    sc.write(str(p))
    back = earfile.read_ear(str(p))
    assert back.to_bytes() == raw
    assert back.triangles().shape == (44, 3, 3)
    # str padding: always at least one NUL, total a multiple of 4 (exporter pack(), __init__.py:166-171)
    assert earfile.pack_str("abcd") == b"str abcd\x00\x00\x00\x00"
    assert earfile.pack_str("abc") == b"str abc\x00"
    # tri record is 88 bytes
    assert len(earfile.pack_tris(np.zeros((1, 3, 3), np.float32))) == 88


def test_material_kept_fraction_matches_reference_derivation():
    sc = scenes.rt60_scene(refl=(0.95, 0.5, 0.0), refr=(0.0, 0.25, 0.0))
    tab = sc.material_table()
    # absorption_coefficient = 1 - (1 - (refl - 1e-9f) - (refr - 1e-9f)) in float32 (src/Material.cpp:33-60)
    assert tab.shape == (1, 3, 4)
    assert np.allclose(tab[0, :, 2], [0.95, 0.75, 0.0], atol=1e-6)
    with pytest.raises(ValueError):
        scenes.rt60_scene(refl=(0.9, 0.9, 0.9), refr=(0.2, 0.2, 0.2)).material_table()


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/ear_b200.h declares."""
    import __graft_entry__ as g
    g.build()
    lib = api.load_library()
    header = open(os.path.join(ROOT, "include", "ear_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ear_b200_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(api.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ear_b200_abi_version() == 3
    assert ctypes.sizeof(api.ContextC) == 56 and ctypes.sizeof(api.RecorderC) == 64 and ctypes.sizeof(api.OptionsC) == 48


def test_no_gpu_means_loud_failure_not_fallback():
    lib = api.load_library()
    if lib.ear_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    sc = scenes.rt60_scene()
    with pytest.raises(api.EarError, match="no CUDA device"):
        api.Scene.from_def(sc)


def test_cli_survives_corrupt_ear_files(tmp_path):
    """Truncated files, flipped bytes, wild lengths: the bounds-checked reader must turn every one of them into the
    reference's "Error: <what>" line (src/EAR.cpp:404-408) and a normal exit -- never a crash.  (Files that still
    parse end at "no CUDA device" here, which is an Error line too.)"""
    exe = os.path.join(ROOT, "ear_b200", "csrc", "EAR")
    wav = scenes.write_click_wav(str(tmp_path / "click.wav"))
    sc = scenes.example1_scene(samples=1000, wav=wav)
    good = str(tmp_path / "good.ear")
    sc.write(good)
    data = open(good, "rb").read()
    rng = np.random.default_rng(0)
    seen = set()
    for k in range(60):
        d = bytearray(data)
        if k % 3 == 0:
            d = d[: int(rng.integers(4, len(d)))]
        elif k % 3 == 1:
            for _ in range(int(rng.integers(1, 8))):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
        else:
            i = int(rng.integers(0, len(d) - 8))
            d[i: i + 4] = int(rng.integers(0, 2 ** 31)).to_bytes(4, "little")
        bad = str(tmp_path / "bad.ear")
        open(bad, "wb").write(bytes(d))
        r = subprocess.run([exe, "calc", "T60", bad], capture_output=True, timeout=60)
        out = r.stdout.decode("utf8", "replace")
        assert r.returncode in (0, 1), (k, r.returncode)
        assert "Error:" in out, (k, out[-300:])
        seen.add(out.split("Error:")[1].strip().split("\n")[0][:12])
    assert len(seen) >= 4      # several different diagnoses, not one catch-all


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ear_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("oracle/shim", "").replace("the oracle", "").replace("oracle path", "") \
                    or f in ("bvh_build.cpp", "device_exact.cuh"), f


@pytest.mark.parametrize("name", ["rt60", "example1", "soup", "hall20k"])
def test_bvh_traversal_logic_is_exact_on_host(oracle_lib, name):
    """Builder margins + traversal tie-breaks, compiled for the host: same winner as the O(T) loop for
    uniform, edge/vertex-aimed, grazing and surface-origin rays; same occlusion answers."""
    sc = common.named_scene(name)
    em = eb.EmulScene(sc.triangles())
    cpu = ob.OracleScene.from_def(sc)
    n = 12000
    o, d = common.make_rays(sc, n, seed=31)
    ei, et = em.first_hit(o, d)
    ci, ct = cpu.first_hit(o, d)
    assert np.array_equal(ei, ci)
    assert np.array_equal(et[ci >= 0].view(np.uint32), ct[ci >= 0].view(np.uint32))
    p, x = common.make_segments(sc, n, seed=32)
    assert np.array_equal(em.occluded(p, x), cpu.occluded(p, x))


def test_bvh_ties_and_degenerate_triangles_on_host(oracle_lib):
    """Duplicated triangles (equal t: lowest original index wins), zero-area triangles, a far outlier, axis-aligned
    rays, origins on vertices, zero-length segments: builder + traversal logic against the O(T) loop."""
    sc, n_orig = common.edge_case_scene()
    em = eb.EmulScene(sc.triangles())
    cpu = ob.OracleScene.from_def(sc)
    o, d, p, x = common.edge_case_queries(sc, 12000)
    ei, et = em.first_hit(o, d)
    ci, ct = cpu.first_hit(o, d)
    assert np.array_equal(ei, ci)
    hit = ci >= 0
    assert np.array_equal(et[hit].view(np.uint32), ct[hit].view(np.uint32))
    assert (ci[hit] < n_orig).mean() > 0.9
    assert np.array_equal(em.occluded(p, x), cpu.occluded(p, x))


@pytest.mark.parametrize("name,res", [("rt60", 64), ("example1", 256), ("soup", 256), ("hall20k", 128), ("hall20k", 1024)])
def test_visibility_map_lists_are_supersets_on_host(oracle_lib, name, res):
    """The recorder visibility maps answer occlusion queries from per-texel candidate lists.  Compiled for the host
    (vismap_geom.cuh): for segments from surface points, vertex / edge points, free-space points and plane-grazing
    points to the recorder -- and to a second end point one centimetre off a wall --, every triangle the reference's
    float test accepts (1e-5 < t < 1) is a candidate of the query's texel."""
    sc = common.named_scene(name)
    em = eb.EmulScene(sc.triangles())
    tris = sc.triangles().reshape(-1, 3)
    lo, hi = tris.min(0), tris.max(0)
    near_wall = np.array([lo[0] + 0.01 * (hi[0] - lo[0]) + 0.01, 0.5 * (lo[1] + hi[1]), lo[2] + 0.37 * (hi[2] - lo[2])], np.float32)
    total_accepted = 0
    for x in (np.asarray(sc.recorders[0].position, np.float32), near_wall):
        p, _ = common.make_segments_to_point(sc, 1200, x, seed=int(res))
        bad, accepted, listed = em.vismap_violations(x, res, p)
        assert bad == 0, f"{bad} accepted triangles missing from their texel list"
        assert listed >= accepted
        total_accepted += accepted
    assert total_accepted > 50      # the check saw real occluders


def test_oracle_honours_stream_id(oracle_lib):
    """ear_b200_context.stream_id pins the context word of the Philox key (the CLI uses it when it deals contexts to
    several GPUs): context k rendered alone with stream_id = k + 1 gives the track it has inside the full call."""
    sc = common.named_scene("example1")
    sc.samples = 2000
    cpu = ob.OracleScene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc)
    full, _ = cpu.render(ctxs, recs, max_bounces=30, seed=4)
    ctxs[1].stream_id = 2
    alone, _ = cpu.render([ctxs[1]], [recs[1]], max_bounces=30, seed=4)
    assert np.array_equal(alone[0][0][0].data, full[1][0][0].data)
    ctxs[1].stream_id = 0
    other, _ = cpu.render([ctxs[1]], [recs[1]], max_bounces=30, seed=4)
    assert not np.array_equal(other[0][0][0].data, full[1][0][0].data)


def test_bvh_build_does_not_depend_on_the_thread_count(oracle_lib, monkeypatch):
    """The builder forks subtrees over a worker pool and chunks its passes over big nodes; the node and triangle-record
    arrays must come out byte-identical whatever the number of threads (one rank builds, the others adopt its image)."""
    sc, _ = scenes.synthetic_hall(n_tris=300000, n_obstacles=500, n_bands=3)
    tris = sc.triangles()
    hashes = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("EAR_B200_BUILD_THREADS", threads)
        hashes.append(eb.EmulScene(tris).layout_hash())
    assert hashes[0] == hashes[1] == hashes[2]


def test_bvh_margins_are_load_bearing(oracle_lib):
    """With pad and slack switched off the edge-aimed rays DO lose their reference winner: the
    adversarial set exercises exactly what the margins are there for."""
    sc = common.named_scene("hall20k")
    cpu = ob.OracleScene.from_def(sc)
    o, d = common.make_rays(sc, 12000, seed=33)
    ci, _ = cpu.first_hit(o, d)
    code = ("import sys; sys.path.insert(0, %r); import numpy as np; from tests import common, emul_binding as eb;"
            "sc = common.named_scene('hall20k'); o, d = common.make_rays(sc, 12000, seed=33);"
            "np.save(sys.argv[1], eb.EmulScene(sc.triangles()).first_hit(o, d)[0])" % ROOT)
    out = os.path.join(os.path.dirname(eb.LIB), "bare.npy")
    subprocess.run([sys.executable, "-c", code, out], check=True, env=dict(os.environ, EAR_B200_BVH_MARGIN_SCALE="0"))
    bare = np.load(out)
    os.unlink(out)
    assert (bare != ci).sum() > 0


def test_bvh_depth_is_bounded_for_degenerate_input(oracle_lib):
    # 4096 identical triangles + a geometric progression of sizes: SAH cannot separate them
    rng = np.random.default_rng(0)
    base = rng.normal(size=(1, 3, 3)).astype(np.float32)
    v = np.concatenate([np.repeat(base, 4096, 0), base * (1.5 ** -np.arange(64, dtype=np.float32))[:, None, None]])
    em = eb.EmulScene(v)
    n_nodes, depth, _ = em.stats()
    assert depth <= 30 + 14
    cpu = ob.OracleScene(v, np.zeros(v.shape[0], np.int32), np.ones((1, 3, 4), np.float32))
    o = rng.normal(size=(500, 3)).astype(np.float32) * 3
    d = -o / np.linalg.norm(o, axis=1, keepdims=True)
    assert np.array_equal(em.first_hit(o, d.astype(np.float32))[0], cpu.first_hit(o, d.astype(np.float32))[0])
