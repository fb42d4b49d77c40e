"""Round-2 GPU parity additions ("what's missing" / "parity holes"):
  * a pool that turns over in lockstep traces every ray
  * C5 shape: >= 1e6 triangles x 66 recorders against the oracle on a ray sample, through every occlusion route
    (visibility maps, map budget exhausted -> BVH any-hit, recorders beyond the 64-map cache) and both splat forms
  * the 1M-triangle hall against the oracle on >= 1e4 rays, and a histogram there against the ORACLE
  * default culling == the rigorous EXACT mode on 1e8 adversarial rays of the 1M-triangle hall
"""
import os

import numpy as np
import pytest

from ear_b200 import api, scenes
from tests import common

pytestmark = pytest.mark.gpu
REL_TOL = 2e-4


@pytest.fixture(scope="module")
def ob():
    from oracle import binding
    binding.build()
    return binding


def test_pool_turning_over_in_lockstep_traces_every_ray(ob, monkeypatch):
    """Closed box, every ray reaches the bounce cap, cap a multiple of the host's check interval, rays >> slots: all
    slots empty at the same launch while the queue still holds rays.  The loop must go on (it used to stop)."""
    monkeypatch.setenv("EAR_B200_SLOTS", "256")
    sc = scenes.rt60_scene(refl=(1.0 - 1e-6,) * 3, spec=(0.0, 0.5, 0.0), samples=10000)
    gpu = api.Scene.from_def(sc)
    cpu = ob.OracleScene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc, t60_only=True)
    for cap in (8, 16, 2):
        ctxs[0].num_samples = 5000
        res = gpu.render(ctxs, recs, max_bounces=cap, seed=5, finalise=False)
        tracks, cnt = cpu.render(ctxs, recs, max_bounces=cap, seed=5, finalise=False)
        assert res.rays == 5000
        assert (res.segments, res.occlusion_queries, res.contributions, res.bin_updates) == \
               (cnt["segments"], cnt["occlusion_queries"], cnt["contributions"], cnt["bin_updates"])
        a, b = res.tracks[0][0][0], tracks[0][0][0]
        assert (a.first_sample, a.real_length) == (b.first_sample, b.real_length)
        n = b.real_length + 1
        assert np.abs(a.data[:n] - b.data[:n]).max() <= REL_TOL * max(np.abs(b.data[:n]).max(), 1e-30)


def test_max_bounces_beyond_the_pool_field_is_rejected():
    gpu = api.Scene.from_def(common.named_scene("rt60"))
    ctxs, recs = api.contexts_from_def(common.named_scene("rt60"), t60_only=True)
    with pytest.raises(api.EarError):
        gpu.render(ctxs, recs, max_bounces=70000, seed=1)


# ---------------------------------------------------------------------------------------------------------
# C5 shape
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def complex_1m():
    sc, table = scenes.synthetic_complex(n_tris=1_000_000, n_obstacles=2000, n_bands=3, seed=0, n_recorders=66)
    return sc, np.ascontiguousarray(table, np.float32)


@pytest.fixture(scope="module")
def complex_oracle(ob, complex_1m):
    sc, table = complex_1m
    cpu = ob.OracleScene(sc.triangles(), sc.triangle_materials(), table)
    af = scenes.air_factors(3)
    ctxs = [api.Context(b, 24, float(af[b]), sc.sources[0].position) for b in (0, 2)]
    recs = [api.Recorder(r.position) for r in sc.recorders]
    want, cnt = common.oracle_render_parallel(cpu, ctxs, recs, max_bounces=10, seed=21)
    return ctxs, recs, want, cnt


@pytest.mark.parametrize("route", ["maps", "budget", "nomaps", "window"])
def test_c5_shape_matches_oracle_through_every_route(complex_1m, complex_oracle, route, monkeypatch):
    """1e6 triangles x 66 mono recorders x 2 bands: counters exact, track ranges exact, bins within REL_TOL of the
    oracle -- with all maps (64 cached + 2 recorders through the BVH), with the map memory budget exhausted after a
    few maps, without maps, and with the shared-memory windowed splat."""
    sc, table = complex_1m
    ctxs, recs, want, cnt = complex_oracle
    if route == "budget":
        monkeypatch.setenv("EAR_B200_VISMAP_BUDGET", str(400e6))
    if route == "nomaps":
        monkeypatch.setenv("EAR_B200_VISMAP_RES", "0")
    if route == "window":
        monkeypatch.setenv("EAR_B200_SPLAT", "window")
    gpu = api.Scene(sc.triangles(), sc.triangle_materials(), table)
    res = gpu.render(ctxs, recs, max_bounces=10, seed=21, finalise=False)
    assert res.rays == 48 and res.dropped_updates == 0
    assert (res.segments, res.occlusion_queries, res.contributions, res.bin_updates) == \
           (cnt["segments"], cnt["occlusion_queries"], cnt["contributions"], cnt["bin_updates"])
    assert res.occlusion_queries > 60 * res.contributions // 66 > 0
    for c in range(len(ctxs)):
        for r in range(len(recs)):
            a, (data, first, real) = res.tracks[c][r][0], want[c][r][0]
            assert (a.first_sample, a.real_length) == (first, real), (c, r)
            n = real + 1
            scale = max(np.abs(data[:n]).max(), 1e-30)
            assert np.abs(a.data[:n] - data[:n]).max() <= REL_TOL * scale, (c, r)
    gpu.close()


# ---------------------------------------------------------------------------------------------------------
# C4 scene size against the oracle
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hall():
    sc, table = scenes.synthetic_hall(n_tris=1_000_000, n_obstacles=2000, n_bands=8, seed=0)
    return sc, np.ascontiguousarray(table, np.float32)


def test_hall_1m_first_hit_and_occlusion_on_1e4_rays(ob, hall):
    sc, table = hall
    gpu = api.Scene(sc.triangles(), sc.triangle_materials(), table)
    cpu = ob.OracleScene(sc.triangles(), sc.triangle_materials(), table)
    o, d = common.make_rays(sc, 12000, seed=141)
    gi, gt = gpu.first_hit(o, d)
    ci, ct = common.oracle_first_hit_parallel(cpu, o, d)
    assert np.array_equal(gi, ci), f"{(gi != ci).sum()} first-hit indices differ"
    assert np.array_equal(gt[ci >= 0].view(np.uint32), ct[ci >= 0].view(np.uint32))
    p, x = common.make_segments_to_point(sc, 6000, sc.recorders[0].position, seed=143)
    assert np.array_equal(gpu.occluded(p, x), common.oracle_occluded_parallel(cpu, p, x))
    p, x = common.make_segments(sc, 6000, seed=142)
    assert np.array_equal(gpu.occluded(p, x), common.oracle_occluded_parallel(cpu, p, x))
    gpu.close()


@pytest.mark.parametrize("splat", ["direct", "window"])
def test_hall_1m_histogram_matches_oracle(ob, hall, splat, monkeypatch):
    sc, table = hall
    monkeypatch.setenv("EAR_B200_SPLAT", splat)
    gpu = api.Scene(sc.triangles(), sc.triangle_materials(), table)
    cpu = ob.OracleScene(sc.triangles(), sc.triangle_materials(), table)
    af = scenes.air_factors(8)
    ctxs = [api.Context(b, 150, float(af[b]), sc.sources[0].position) for b in (1, 6)]
    recs = [api.Recorder(sc.recorders[0].position)]
    want, cnt = common.oracle_render_parallel(cpu, ctxs, recs, max_bounces=50, seed=77)
    res = gpu.render(ctxs, recs, max_bounces=50, seed=77, finalise=False)
    assert (res.rays, res.segments, res.occlusion_queries, res.contributions, res.bin_updates) == \
           (300, cnt["segments"], cnt["occlusion_queries"], cnt["contributions"], cnt["bin_updates"])
    for c in range(2):
        a, (data, first, real) = res.tracks[c][0][0], want[c][0][0]
        assert (a.first_sample, a.real_length) == (first, real)
        n = real + 1
        assert np.abs(a.data[:n] - data[:n]).max() <= REL_TOL * np.abs(data[:n]).max()
    gpu.close()


def test_default_culling_equals_exact_on_1e8_adversarial_rays(hall, monkeypatch):
    """The default child-culling rule (entry later than best_t (1 + 2^-10) + s0) is a bound only under an assumption on
    the reference's own hit point; EAR_B200_EXACT_SLACK=1 is rigorous.  1e8 rays of the usual
    adversarial mix (a quarter each: uniform, aimed at vertices / edge points, grazing a triangle's plane, leaving a
    surface point) on the bench scene must give bit-identical (index, t) in both modes."""
    sc, table = hall
    fast = api.Scene(sc.triangles(), sc.triangle_materials(), table)
    monkeypatch.setenv("EAR_B200_EXACT_SLACK", "1")
    exact = api.Scene(sc.triangles(), sc.triangle_materials(), table)
    total = int(float(os.environ.get("EAR_TEST_EXACT_RAYS", "1e8")))
    chunk = 12_500_000
    done = mismatches = hits = 0
    seed = 1000
    while done < total:
        n = min(chunk, total - done)
        o, d = common.make_rays(sc, n, seed=seed)
        fi, ft = fast.first_hit(o, d)
        ei, et = exact.first_hit(o, d)
        hit = ei >= 0
        mismatches += int((fi != ei).sum()) + int((ft[hit].view(np.uint32) != et[hit].view(np.uint32)).sum())
        hits += int(hit.sum())
        done += n
        seed += 1
    assert mismatches == 0, f"{mismatches} of {done} rays differ between default and EXACT culling"
    assert hits > 0.5 * done


# ---------------------------------------------------------------------------------------------------------
# mesh-emitter sources (src/SoundFile.cpp:215-228, src/Mesh.cpp:143-154, src/Triangle.cpp:44-53)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("stereo", [False, True])
def test_mesh_source_paths_and_histograms_match_oracle(ob, stereo):
    from tests.test_oracle_pinning import _mesh_source_scene
    sc = _mesh_source_scene("/tmp/click.wav", samples=60000, stereo=stereo)
    gpu = api.Scene.from_def(sc)
    cpu = ob.OracleScene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc)
    assert ctxs[0].emitter == (0, 3)
    # bounce paths: emission point / direction and every later hit identical (same Philox streams)
    hg, sg = gpu.trace_paths(ctxs[1], 1, 4000, 40, seed=17)
    hc, sc_ = cpu.trace_paths(ctxs[1], 1, 4000, 40, seed=17)
    assert np.array_equal(hg, hc)
    # geometry of the final state bit for bit; the running intensity within an ulp or two: it is a product of
    # pow(af, length) factors, and the GPU's pow (exp2(y log2 x) in double, rounded once) differs from the host libm's powf
    # by one ulp in ~3e-4 of the evaluations (which variant of powf runs even depends on the CPU: glibc selects an FMA build)
    assert np.array_equal(sg[:, :6].view(np.uint32), sc_[:, :6].view(np.uint32)) and np.array_equal(sg[:, 7], sc_[:, 7])
    assert np.abs(sg[:, 6] - sc_[:, 6]).max() <= 1e-6 * np.abs(sc_[:, 6]).max()
    res = gpu.render(ctxs, recs, max_bounces=120, seed=17)
    tracks, cnt = cpu.render(ctxs, recs, max_bounces=120, seed=17)
    assert (res.rays, res.segments, res.occlusion_queries, res.contributions, res.bin_updates) == \
           (cnt["rays"], cnt["segments"], cnt["occlusion_queries"], cnt["contributions"], cnt["bin_updates"])
    # bounce 0 is recorded for mesh sources: more Connect() calls than segments that hit something
    assert res.occlusion_queries > res.segments - res.rays
    for c in range(3):
        for k in range(2 if stereo else 1):
            a, b = res.tracks[c][0][k], tracks[c][0][k]
            assert (a.first_sample, a.real_length) == (b.first_sample, b.real_length)
            n = b.real_length + 1
            assert np.abs(a.data[:n] - b.data[:n]).max() <= REL_TOL * np.abs(b.data[:n]).max()


def test_group_render_on_one_gpu_equals_render(ob):
    """ear_b200_group_render with a single device is ear_b200_render (the CLI always goes through the group)."""
    sc = common.named_scene("example1")
    gpu = api.Scene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc)
    for c in ctxs:
        c.num_samples = 4000
    a = gpu.render(ctxs, recs, max_bounces=100, seed=5)
    grp = api.Group(gpu, [0])
    b = grp.render(ctxs, recs, max_bounces=100, seed=5)
    assert (a.rays, a.segments, a.contributions, a.bin_updates) == (b.rays, b.segments, b.contributions, b.bin_updates)
    for c in range(len(ctxs)):
        ta, tb = a.tracks[c][0][0], b.tracks[c][0][0]
        assert (ta.first_sample, ta.real_length) == (tb.first_sample, tb.real_length)
        assert np.abs(ta.data - tb.data).max() <= 1e-5 * np.abs(ta.data).max()
    # the device post chain inside render: Power / Truncate / T60 as the oracle's post chain gives them
    post = grp.render(ctxs, recs, max_bounces=100, seed=5, post=(0.335, 256.0))
    tracks, _ = ob.OracleScene.from_def(sc).render(ctxs, recs, max_bounces=100, seed=5)
    want_max, want = ob.post_all(tracks)
    assert post.maximum == pytest.approx(want_max, rel=2e-6)
    for c in range(len(ctxs)):
        data, first, real, w_t60 = want[c][0][0]
        t = post.tracks[c][0][0]
        assert (t.first_sample, t.real_length) == (first, real)
        assert abs(post.t60[c][0][0] - w_t60) <= 1.5 / 44100.0
    grp.close()


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 3: example2 (232 triangles, 2 materials, 10 keyframes, 3 sources -> 90 contexts)
# ---------------------------------------------------------------------------------------------------------
def test_example2_ninety_contexts_match_oracle(ob):
    sc = scenes.example2_scene(samples=10000)          # 1000 rays per context here; C3 itself uses 1e4
    gpu = api.Scene.from_def(sc)
    cpu = ob.OracleScene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc)
    assert len(ctxs) == 90
    res = gpu.render(ctxs, recs, max_bounces=100, seed=33, finalise=False)
    want, cnt = common.oracle_render_parallel(cpu, ctxs, recs, max_bounces=100, seed=33)
    assert (res.rays, res.segments, res.occlusion_queries, res.contributions, res.bin_updates) == \
           (90 * 1000, cnt["segments"], cnt["occlusion_queries"], cnt["contributions"], cnt["bin_updates"])
    for c in range(90):
        a, (data, first, real) = res.tracks[c][0][0], want[c][0][0]
        assert (a.first_sample, a.real_length) == (first, real), c
        n = real + 1
        assert np.abs(a.data[:n] - data[:n]).max() <= REL_TOL * max(np.abs(data[:n]).max(), 1e-30), c
    # and the bounce paths of a keyframed context late in the list
    hg, _ = gpu.trace_paths(ctxs[77], 77, 3000, 60, seed=33)
    hc, _ = cpu.trace_paths(ctxs[77], 77, 3000, 60, seed=33)
    assert np.array_equal(hg, hc)
