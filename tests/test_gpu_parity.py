"""GPU parity tests: libear_b200.so (through the C ABI) against the CPU oracle on the same inputs.

Bars: first-hit triangle indices bit-exact and hit distances bit-exact (north star asks 1e-5
relative); occlusion answers identical; bounce paths (emission, Material::Bounce, Sample_Hemi,
reflection) identical triangle by triangle under the shared Philox streams; histograms within
REL_TOL of the oracle's per bin (float atomics reorder the sums; powf differs by <= 2 ulp)."""
import numpy as np
import pytest

from ear_b200 import api
from tests import common

pytestmark = pytest.mark.gpu

REL_TOL = 2e-4     # per-bin |gpu - oracle| <= REL_TOL * max|oracle| over the track


@pytest.fixture(scope="module")
def ob():
    from oracle import binding
    binding.build()
    return binding


def _pair(ob, sc):
    return api.Scene.from_def(sc), ob.OracleScene.from_def(sc)


@pytest.mark.parametrize("name,n", [("rt60", 40000), ("example1", 40000), ("soup", 40000), ("hall20k", 24000)])
def test_first_hit_bit_exact(ob, name, n):
    sc = common.named_scene(name)
    gpu, cpu = _pair(ob, sc)
    o, d = common.make_rays(sc, n)
    gi, gt = gpu.first_hit(o, d)
    ci, ct = cpu.first_hit(o, d)
    assert np.array_equal(gi, ci), f"{(gi != ci).sum()} of {n} first-hit indices differ"
    hit = ci >= 0
    assert hit.mean() > 0.3
    assert np.array_equal(gt[hit].view(np.uint32), ct[hit].view(np.uint32))


@pytest.mark.parametrize("name,n", [("rt60", 1000000), ("example1", 1000000), ("soup", 100000)])
def test_first_hit_and_occlusion_at_survey_ray_counts(ob, name, n):
    """SURVEY.md section 8(d) parity procedure (i): 10^6 seeded rays per scene (uniform, edge/vertex-aimed, grazing,
    surface-leaving) -- indices identical, distances bit-identical, occlusion answers identical."""
    sc = common.named_scene(name)
    gpu, cpu = _pair(ob, sc)
    o, d = common.make_rays(sc, n, seed=77)
    gi, gt = gpu.first_hit(o, d)
    ci, ct = cpu.first_hit(o, d)
    assert np.array_equal(gi, ci), f"{(gi != ci).sum()} of {n} first-hit indices differ"
    hit = ci >= 0
    assert np.array_equal(gt[hit].view(np.uint32), ct[hit].view(np.uint32))
    p, x = common.make_segments(sc, n, seed=78)
    assert np.array_equal(gpu.occluded(p, x), cpu.occluded(p, x))


@pytest.mark.parametrize("name,n", [("rt60", 40000), ("example1", 40000), ("soup", 40000), ("hall20k", 24000)])
def test_occlusion_identical(ob, name, n):
    sc = common.named_scene(name)
    gpu, cpu = _pair(ob, sc)
    p, x = common.make_segments(sc, n)
    g = gpu.occluded(p, x)
    c = cpu.occluded(p, x)
    assert np.array_equal(g, c), f"{(g != c).sum()} of {n} occlusion answers differ"


@pytest.mark.parametrize("name,n", [("rt60", 40000), ("example1", 40000), ("soup", 40000), ("hall20k", 24000)])
def test_occlusion_to_recorder_identical(ob, name, n):
    """All segments end at the recorder: answered through the recorder's visibility map (+ BVH fallback), the
    way the render loop does it."""
    sc = common.named_scene(name)
    gpu, cpu = _pair(ob, sc)
    for x in (sc.recorders[0].position, sc.sources[0].position):
        p, xx = common.make_segments_to_point(sc, n, x)
        g = gpu.occluded(p, xx)
        c = cpu.occluded(p, xx)
        assert np.array_equal(g, c), f"{(g != c).sum()} of {n} occlusion answers differ"
        assert 0.01 < c.mean() < 0.99 or name == "rt60"


def test_empty_and_tiny_scenes(ob):
    # one triangle, and a ray batch of size 1 / 0
    sc = common.soup_scene(n_tris=2, seed=9)
    gpu, cpu = _pair(ob, sc)
    o, d = common.make_rays(sc, 64)
    assert np.array_equal(gpu.first_hit(o, d)[0], cpu.first_hit(o, d)[0])
    gi, _ = gpu.first_hit(o[:1], d[:1])
    assert gi.shape == (1,)
    gi, _ = gpu.first_hit(o[:0], d[:0])
    assert gi.shape == (0,)


@pytest.mark.parametrize("name,band", [("rt60", 1), ("rt60_saved", 1), ("example1", 2), ("soup", 0), ("hall20k", 1)])
def test_bounce_paths_identical(ob, name, band):
    sc = common.named_scene(name)
    gpu, cpu = _pair(ob, sc)
    ctx = api.Context(band, 1000, 0.999, sc.sources[0].position)
    n, mb = (300, 40) if name == "hall20k" else (1500, 60)
    hg, sg = gpu.trace_paths(ctx, 3, n, mb, seed=11, first_ray=17)
    hc, s_c = cpu.trace_paths(ctx, 3, n, mb, seed=11, first_ray=17)
    assert np.array_equal(hg, hc), f"{(hg != hc).any(axis=1).sum()} of {n} paths differ"
    # geometry of the final ray state is bit-exact; intensity goes through powf (<= a few ulp per bounce)
    assert np.array_equal(sg[:, :6].view(np.uint32), s_c[:, :6].view(np.uint32))
    assert np.allclose(sg[:, 6], s_c[:, 6], rtol=1e-4, atol=0)
    assert np.array_equal(sg[:, 7].view(np.uint32), s_c[:, 7].view(np.uint32))


def _compare_tracks(res, tracks):
    worst = 0.0
    for c in range(len(tracks)):
        for r in range(len(tracks[c])):
            for k in range(len(tracks[c][r])):
                a, b = res.tracks[c][r][k], tracks[c][r][k]
                assert (a.first_sample, a.real_length) == (b.first_sample, b.real_length)
                n = b.real_length + 1
                scale = max(float(np.abs(b.data[:n]).max()), 1e-30)   # a recorder no ray reached has an all-zero track
                err = np.abs(a.data[:n].astype(np.float64) - b.data[:n]).max() / scale
                worst = max(worst, err)
                assert not a.data[n:].any()
    return worst


@pytest.mark.parametrize("name,stereo", [("rt60", False), ("rt60", True), ("example1", True), ("soup", False)])
def test_histograms_match_oracle(ob, name, stereo):
    sc = common.named_scene(name)
    sc.samples = 20000
    for rec in sc.recorders:
        rec.stereo = stereo
    gpu, cpu = _pair(ob, sc)
    ctxs, recs = api.contexts_from_def(sc)
    res = gpu.render(ctxs, recs, max_bounces=120, seed=5)
    tracks, cnt = cpu.render(ctxs, recs, max_bounces=120, seed=5)
    assert res.rays == cnt["rays"] and res.segments == cnt["segments"]
    assert res.occlusion_queries == cnt["occlusion_queries"]
    assert res.contributions == cnt["contributions"] and res.bin_updates == cnt["bin_updates"]
    assert res.dropped_updates == 0
    assert _compare_tracks(res, tracks) < REL_TOL


def test_multiple_recorders_and_raw_shards_add_up(ob):
    """Two ray shards traced without finalise sum to the unsharded raw histogram (the multi-GPU
    contract: partial sums, one reduce, then finalise)."""
    sc = common.named_scene("example1")
    sc.samples = 8000
    sc.recorders.append(type(sc.recorders[0])("/tmp/b.wav", position=(-8.0, 3.0, 2.0), stereo=True, right_ear=(0.0, 1.0, 0.0)))
    gpu, cpu = _pair(ob, sc)
    ctxs, recs = api.contexts_from_def(sc)
    full = gpu.render(ctxs, recs, max_bounces=60, seed=2, finalise=False)
    a = gpu.render(ctxs, recs, max_bounces=60, seed=2, finalise=False, first_ray=0, ray_count=300)
    b = gpu.render(ctxs, recs, max_bounces=60, seed=2, finalise=False, first_ray=300, ray_count=500)
    assert a.rays + b.rays == full.rays and a.segments + b.segments == full.segments
    for c in range(len(ctxs)):
        for r in range(2):
            for k in range(2):
                s = a.tracks[c][r][k].data.astype(np.float64) + b.tracks[c][r][k].data
                f = full.tracks[c][r][k].data
                assert np.abs(s - f).max() <= 1e-4 * np.abs(f).max()
                assert min(a.tracks[c][r][k].first_sample, b.tracks[c][r][k].first_sample) == full.tracks[c][r][k].first_sample
                assert max(a.tracks[c][r][k].real_length, b.tracks[c][r][k].real_length) == full.tracks[c][r][k].real_length
    tracks, _ = cpu.render(ctxs, recs, max_bounces=60, seed=2, finalise=False)
    assert _compare_tracks(full, tracks) < REL_TOL


def test_errors_are_reported_not_swallowed():
    sc = common.named_scene("rt60")
    gpu = api.Scene.from_def(sc)
    with pytest.raises(api.EarError):
        gpu.render([api.Context(7, 10, 1.0, (0, 0, 1))], [api.Recorder((1, 0, 1))])   # band outside the table
    with pytest.raises(api.EarError):
        api.Scene(sc.triangles(), np.full(12, 3, np.int32), sc.material_table())       # bad material index


def test_scene_image_round_trip(ob):
    """Multi-GPU replication path on one GPU: a scene adopted from another scene's device image answers every
    query and renders exactly like the scene that was built from the triangles; a damaged image is refused."""
    import torch
    sc = common.named_scene("example1")
    sc.samples = 6000
    built = api.Scene.from_def(sc)
    n = built.image_size()
    assert n > 0 and n % 256 == 0
    buf = torch.zeros((n,), dtype=torch.uint8, device="cuda:0")
    built.image_write(buf.data_ptr(), n)
    torch.cuda.synchronize()
    adopted = api.Scene.from_image(buf.data_ptr(), n, built.n_bands, device=0)
    del buf
    o, d = common.make_rays(sc, 20000)
    bi, bt = built.first_hit(o, d)
    ai, at = adopted.first_hit(o, d)
    assert np.array_equal(ai, bi) and np.array_equal(at.view(np.uint32)[bi >= 0], bt.view(np.uint32)[bi >= 0])
    ctxs, recs = api.contexts_from_def(sc)
    ra = adopted.render(ctxs, recs, max_bounces=40, seed=5)
    rb = built.render(ctxs, recs, max_bounces=40, seed=5)
    assert ra.segments == rb.segments and ra.contributions == rb.contributions
    for c in range(len(ctxs)):
        f, g = ra.tracks[c][0][0], rb.tracks[c][0][0]
        assert (f.first_sample, f.real_length) == (g.first_sample, g.real_length)
        assert np.abs(f.data - g.data).max() <= 1e-4 * np.abs(g.data).max()
    bad = torch.zeros((n,), dtype=torch.uint8, device="cuda:0")
    built.image_write(bad.data_ptr(), n)
    bad[0] = 0   # magic
    with pytest.raises(api.EarError):
        api.Scene.from_image(bad.data_ptr(), n, built.n_bands, device=0)
    with pytest.raises(api.EarError):
        api.Scene.from_image(bad.data_ptr(), 128, built.n_bands, device=0)
    with pytest.raises(api.EarError):
        built.image_write(bad.data_ptr(), n - 256)


def test_many_recorders_match_oracle(ob):
    """C5's recorder count and beyond: 66 recorders (mono and stereo mixed) in one call.  The library caches 64
    visibility maps per scene, so the last recorders take the BVH any-hit path -- the answers must not depend
    on which path a recorder got."""
    sc = common.named_scene("example1")
    sc.samples = 3000
    rng = np.random.default_rng(5)
    first = sc.recorders[0]
    sc.recorders = []
    for i in range(66):
        pos = (float(rng.uniform(-9, 9)), float(rng.uniform(-5, 5)), float(rng.uniform(0.5, 3.5)))
        ear = rng.normal(size=3)
        ear /= np.linalg.norm(ear)
        sc.recorders.append(type(first)(f"/tmp/r{i}.wav", position=pos, stereo=(i % 3 == 0),
                                        right_ear=tuple(float(x) for x in ear)))
    gpu, cpu = _pair(ob, sc)
    ctxs, recs = api.contexts_from_def(sc)
    res = gpu.render(ctxs, recs, max_bounces=30, seed=21)
    tracks, cnt = cpu.render(ctxs, recs, max_bounces=30, seed=21)
    assert res.segments == cnt["segments"] and res.occlusion_queries == cnt["occlusion_queries"]
    assert res.contributions == cnt["contributions"]
    assert _compare_tracks(res, tracks) < REL_TOL


def test_ties_degenerate_triangles_and_axis_aligned_rays(ob):
    """Collisions and nulls as this domain has them (tests/common.py::edge_case_scene / edge_case_queries): duplicated
    triangles (lowest index wins the tie), zero-area triangles, a huge far triangle, axis-aligned rays, origins exactly
    on vertices, zero-length segments."""
    sc, n_orig = common.edge_case_scene()
    gpu, cpu = _pair(ob, sc)
    o, d, p, x = common.edge_case_queries(sc)
    gi, gt = gpu.first_hit(o, d)
    ci, ct = cpu.first_hit(o, d)
    assert np.array_equal(gi, ci), f"{(gi != ci).sum()} of {o.shape[0]} first-hit indices differ"
    hit = ci >= 0
    assert np.array_equal(gt[hit].view(np.uint32), ct[hit].view(np.uint32))
    assert (ci[hit] < n_orig).mean() > 0.9                         # the duplicates (higher indices) lose the ties
    assert np.array_equal(gpu.occluded(p, x), cpu.occluded(p, x))


@pytest.mark.parametrize("env", [{"EAR_B200_VISMAP_CAP": "3"}, {"EAR_B200_VISMAP_RES": "0"}, {"EAR_B200_VISMAP_RES": "64"},
                                 {"EAR_B200_EXACT_SLACK": "1"}])
def test_occlusion_paths_agree(ob, monkeypatch, env):
    """The render loop answers occlusion queries three ways -- visibility-map lists, BVH any-hit for texels whose list
    is over the cap, BVH any-hit when maps are off -- and in the rigorous-slack mode; every mix must give the oracle's
    histogram (knobs are read at scene creation)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sc = common.named_scene("hall20k")
    sc.samples = 3000
    gpu, cpu = _pair(ob, sc)
    ctxs, recs = api.contexts_from_def(sc)
    ctxs, recs = ctxs[:1], recs[:1]
    res = gpu.render(ctxs, recs, max_bounces=25, seed=13)
    tracks, cnt = cpu.render(ctxs, recs, max_bounces=25, seed=13)
    assert res.segments == cnt["segments"] and res.occlusion_queries == cnt["occlusion_queries"]
    assert res.contributions == cnt["contributions"] and res.bin_updates == cnt["bin_updates"]
    assert _compare_tracks(res, tracks) < REL_TOL


def test_stream_id_pins_the_random_streams(ob):
    """A context rendered alone with stream_id = k + 1 follows the paths it has as context k of the full call
    (how the CLI keeps its output independent of the number of GPUs the contexts are dealt to)."""
    sc = common.named_scene("example1")
    sc.samples = 8000
    gpu, cpu = _pair(ob, sc)
    ctxs, recs = api.contexts_from_def(sc)
    full = gpu.render(ctxs, recs, max_bounces=60, seed=4)
    segs = 0
    for k in range(len(ctxs)):
        ctxs[k].stream_id = k + 1
        alone = gpu.render([ctxs[k]], [recs[k]], max_bounces=60, seed=4)
        segs += alone.segments
        for a, f in zip(alone.tracks[0][0], full.tracks[k][0]):
            assert (a.first_sample, a.real_length) == (f.first_sample, f.real_length)
            assert np.abs(a.data - f.data).max() <= 1e-4 * np.abs(f.data).max()
    assert segs == full.segments
    ctxs[1].stream_id = 0
    other = gpu.render([ctxs[1]], [recs[1]], max_bounces=60, seed=4)      # keyed as context 0 now: different paths
    assert not np.array_equal(other.tracks[0][0][0].data, full.tracks[1][0][0].data)
