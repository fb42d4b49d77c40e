"""K0 on the device (ear_b200/csrc/bvh_device.cuh): the tree it builds must give the reference's first hit and the
reference's occlusion answers like the host builder's tree does -- any valid tree must -- on scenes
of every size class: a 12-triangle room (single bottom-phase subtree that ends in leaves at once), 44 triangles, a
3000-triangle soup with duplicates and degenerates (several top levels + bottom phase), the 20k hall, edge cases."""
import numpy as np
import pytest

from ear_b200 import api, scenes
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    from oracle import binding
    binding.build()
    return binding


@pytest.mark.parametrize("name,n", [("rt60", 20000), ("example1", 20000), ("soup", 40000), ("hall20k", 24000)])
def test_device_built_tree_gives_the_reference_answers(ob, name, n, monkeypatch):
    monkeypatch.setenv("EAR_B200_BUILD", "device")
    sc = common.named_scene(name)
    gpu = api.Scene.from_def(sc)
    cpu = ob.OracleScene.from_def(sc)
    o, d = common.make_rays(sc, n, seed=51)
    gi, gt = gpu.first_hit(o, d)
    ci, ct = cpu.first_hit(o, d)
    assert np.array_equal(gi, ci), f"{(gi != ci).sum()} of {n} first-hit indices differ"
    hit = ci >= 0
    assert np.array_equal(gt[hit].view(np.uint32), ct[hit].view(np.uint32))
    p, x = common.make_segments(sc, n, seed=52)
    assert np.array_equal(gpu.occluded(p, x), cpu.occluded(p, x))
    p, x = common.make_segments_to_point(sc, n, sc.recorders[0].position, seed=53)
    assert np.array_equal(gpu.occluded(p, x), cpu.occluded(p, x))


def test_device_build_edge_cases(ob, monkeypatch):
    """Duplicated triangles (index ties), zero-area triangles, a far outlier; and the degenerate inputs the host
    builder's depth guard exists for: thousands of identical triangles + a geometric progression of sizes."""
    monkeypatch.setenv("EAR_B200_BUILD", "device")
    sc, _ = common.edge_case_scene()
    gpu = api.Scene.from_def(sc)
    cpu = ob.OracleScene.from_def(sc)
    o, d, p, x = common.edge_case_queries(sc, 20000)
    gi, gt = gpu.first_hit(o, d)
    ci, ct = cpu.first_hit(o, d)
    assert np.array_equal(gi, ci)
    assert np.array_equal(gt[ci >= 0].view(np.uint32), ct[ci >= 0].view(np.uint32))
    assert np.array_equal(gpu.occluded(p, x), cpu.occluded(p, x))
    rng = np.random.default_rng(0)
    base = rng.normal(size=(1, 3, 3)).astype(np.float32)
    v = np.concatenate([np.repeat(base, 4096, 0), base * (1.5 ** -np.arange(64, dtype=np.float32))[:, None, None]])
    table = np.ones((1, 3, 4), np.float32)
    gpu2 = api.Scene(v, np.zeros(v.shape[0], np.int32), table)
    cpu2 = ob.OracleScene(v, np.zeros(v.shape[0], np.int32), table)
    o = rng.normal(size=(2000, 3)).astype(np.float32) * 3
    d = (-o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    assert np.array_equal(gpu2.first_hit(o, d)[0], cpu2.first_hit(o, d)[0])


def test_device_and_host_builders_render_the_same_tracks(ob, monkeypatch):
    """Same Philox paths through either tree: identical counters and track ranges, bins equal up to the order of the
    float atomics."""
    sc = common.named_scene("hall20k")
    ctxs, recs = api.contexts_from_def(sc)
    for c in ctxs:
        c.num_samples = 3000
    out = []
    for how in ("host", "device"):
        monkeypatch.setenv("EAR_B200_BUILD", how)
        out.append(api.Scene.from_def(sc).render(ctxs, recs, max_bounces=60, seed=11))
    a, b = out
    assert (a.rays, a.segments, a.occlusion_queries, a.contributions, a.bin_updates) == \
           (b.rays, b.segments, b.occlusion_queries, b.contributions, b.bin_updates)
    for c in range(len(ctxs)):
        ta, tb = a.tracks[c][0][0], b.tracks[c][0][0]
        assert (ta.first_sample, ta.real_length) == (tb.first_sample, tb.real_length)
        assert np.abs(ta.data - tb.data).max() <= 1e-5 * np.abs(ta.data).max()


def test_empty_and_tiny_scenes_on_the_device_builder(monkeypatch):
    monkeypatch.setenv("EAR_B200_BUILD", "device")
    table = np.ones((1, 3, 4), np.float32)
    empty = api.Scene(np.zeros((0, 3, 3), np.float32), np.zeros(0, np.int32), table)
    idx, _ = empty.first_hit(np.zeros((4, 3), np.float32), np.tile(np.array([[0, 0, 1]], np.float32), (4, 1)))
    assert (idx == -1).all()
    one = api.Scene(np.array([[[0, 0, 1], [1, 0, 1], [0, 1, 1]]], np.float32), np.zeros(1, np.int32), table)
    idx, t = one.first_hit(np.array([[0.2, 0.2, 0.0], [2.0, 2.0, 0.0]], np.float32), np.array([[0, 0, 1], [0, 0, 1]], np.float32))
    assert idx.tolist() == [0, -1] and t[0] == 1.0
