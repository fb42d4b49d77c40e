"""Parity at BASELINE.json's full scene size (1M-triangle hall): the brute-force oracle on a sample it can finish
in seconds, plus size-independent properties on 10^5..10^6 queries."""
import numpy as np
import pytest

from ear_b200 import api, scenes
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hall():
    sc, table = scenes.synthetic_hall(n_tris=1_000_000, n_obstacles=2000, n_bands=8, seed=0)
    gpu = api.Scene(sc.triangles(), sc.triangle_materials(), table)
    return sc, table, gpu


def test_first_hit_matches_bruteforce_oracle_on_a_sample(hall):
    from oracle import binding as ob
    sc, table, gpu = hall
    cpu = ob.OracleScene(sc.triangles(), sc.triangle_materials(), table)
    o, d = common.make_rays(sc, 600, seed=41)          # uniform / edge-aimed / grazing / surface-origin mix
    gi, gt = gpu.first_hit(o, d)
    ci, ct = cpu.first_hit(o, d)
    assert np.array_equal(gi, ci)
    assert np.array_equal(gt[ci >= 0].view(np.uint32), ct[ci >= 0].view(np.uint32))
    p, x = common.make_segments(sc, 400, seed=42)
    assert np.array_equal(gpu.occluded(p, x), cpu.occluded(p, x))
    p, x = common.make_segments_to_point(sc, 400, sc.recorders[0].position, seed=43)   # through the visibility map
    assert np.array_equal(gpu.occluded(p, x), cpu.occluded(p, x))


def test_closest_hit_and_occlusion_agree_with_each_other(hall):
    """From a point in free air: the segment to just before the first hit is clear, the segment to beyond it is
    blocked (both answers come from different kernels and a different float test range)."""
    sc, table, gpu = hall
    rng = np.random.default_rng(7)
    n = 400_000
    o = np.tile(np.asarray(sc.sources[0].position, np.float32), (n, 1))
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    idx, t = gpu.first_hit(o, d)
    hit = idx >= 0
    assert hit.mean() > 0.99                       # closed hall: only edge leaks escape
    before = o[hit] + d[hit] * (t[hit] * 0.98)[:, None]
    beyond = o[hit] + d[hit] * (t[hit] * 1.25)[:, None]
    assert gpu.occluded(o[hit], before).mean() < 1e-3
    assert gpu.occluded(o[hit], beyond).mean() > 0.999
    # sortedness property of the winner: no triangle is hit strictly closer along the same ray
    idx2, t2 = gpu.first_hit(o[hit], d[hit])
    assert np.array_equal(idx2, idx[hit]) and np.array_equal(t2.view(np.uint32), t[hit].view(np.uint32))


def test_render_is_reproducible_and_shards_add_up_at_full_scene_size(hall):
    sc, table, gpu = hall
    af = scenes.air_factors(8)
    ctxs = [api.Context(b, 40000, float(af[b]), sc.sources[0].position) for b in (0, 7)]
    recs = [api.Recorder(sc.recorders[0].position)]
    a = gpu.render(ctxs, recs, max_bounces=50, seed=99, finalise=False)
    b = gpu.render(ctxs, recs, max_bounces=50, seed=99, finalise=False)
    assert (a.rays, a.segments, a.occlusion_queries, a.contributions, a.bin_updates) == \
           (b.rays, b.segments, b.occlusion_queries, b.contributions, b.bin_updates)
    assert a.dropped_updates == 0 and a.segments > 40 * a.rays
    for c in range(2):
        ta, tb = a.tracks[c][0][0], b.tracks[c][0][0]
        assert (ta.first_sample, ta.real_length) == (tb.first_sample, tb.real_length)
        assert np.abs(ta.data - tb.data).max() <= 1e-5 * np.abs(ta.data).max()      # atomics reorder the sums
    lo = gpu.render(ctxs, recs, max_bounces=50, seed=99, finalise=False, first_ray=0, ray_count=15000)
    hi = gpu.render(ctxs, recs, max_bounces=50, seed=99, finalise=False, first_ray=15000, ray_count=25000)
    assert lo.segments + hi.segments == a.segments and lo.bin_updates + hi.bin_updates == a.bin_updates
    for c in range(2):
        s = lo.tracks[c][0][0].data.astype(np.float64) + hi.tracks[c][0][0].data
        assert np.abs(s - a.tracks[c][0][0].data).max() <= 1e-5 * np.abs(a.tracks[c][0][0].data).max()
