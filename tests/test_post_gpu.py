"""Device post chain against the oracle's restatement of src/EAR.cpp:209-228 and
src/Recorder.cpp:76-118,303-340 on the SAME input tracks: Power values within one float32 ulp (pow route differs from
glibc's only near rounding boundaries), track maxima and truncated lengths identical, T60 within one sample."""
import ctypes as C

import numpy as np
import pytest

from ear_b200 import api
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    from oracle import binding
    binding.build()
    return binding


def _device_post(scene, res, recs, n_ctx, exponent=0.335, divisor=256.0):
    import torch
    lib = scene.lib
    rec_c, n_rec = api.pack_recorders(recs, n_ctx)
    tpr = api.tracks_per_recorder(rec_c)   # device layout [context][recorder][tpr][bins]: 1 when every recorder is mono
    flat = []
    for c in range(n_ctx):
        for r in range(n_rec):
            pair = res.tracks[c][r]
            flat.append(pair[0])
            if tpr == 2:
                flat.append(pair[1] if len(pair) > 1 else None)
    n_bins = max(t.data.shape[0] for t in flat if t is not None)
    hist = np.zeros((len(flat), n_bins), np.float32)
    rng = np.zeros((len(flat), 2), np.uint32)
    rng[:, 0] = api.FIRST_SAMPLE_INIT
    for i, t in enumerate(flat):
        if t is not None:
            hist[i, : t.data.shape[0]] = t.data
            rng[i] = (t.first_sample, t.real_length)
    d_hist = torch.from_numpy(hist).cuda()
    d_rng = torch.from_numpy(rng.view(np.int32)).cuda()
    mx = C.c_float(0.0)
    tmax = np.zeros(len(flat), np.float32)
    api._check(lib, lib.ear_b200_post_power_device(scene.handle, rec_c, n_ctx, n_rec, n_bins, d_hist.data_ptr(), d_rng.data_ptr(),
                                                   exponent, C.byref(mx), tmax.ctypes.data, None))
    thr = float(np.float32(mx.value) / np.float32(divisor))
    t60 = np.zeros(len(flat), np.float32)
    api._check(lib, lib.ear_b200_post_truncate_device(scene.handle, rec_c, n_ctx, n_rec, n_bins, d_hist.data_ptr(), d_rng.data_ptr(),
                                                      thr, t60.ctypes.data, None))
    return float(mx.value), d_hist.cpu().numpy(), d_rng.cpu().numpy().view(np.uint32), t60, tmax, tpr


@pytest.mark.parametrize("name,stereo,samples", [("rt60", False, 40000), ("example1", True, 30000), ("rt60", True, 600)])
def test_post_chain_matches_oracle(ob, name, stereo, samples):
    sc = common.named_scene(name)
    sc.samples = samples
    for rec in sc.recorders:
        rec.stereo = stereo
    scene = api.Scene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc)
    res = scene.render(ctxs, recs, max_bounces=300, seed=3)
    want_max, want = ob.post_all(res.tracks)
    got_max, hist, rng, t60, tmax, tpr = _device_post(scene, res, recs, len(ctxs))
    assert got_max == pytest.approx(want_max, rel=2e-7)
    i = 0
    for c in range(len(ctxs)):
        for r in range(len(recs[c])):
            for k in range(tpr):
                if k >= len(want[c][r]):
                    assert t60[i] == 0.0 and tmax[i] == 0.0
                    i += 1
                    continue
                data, first, real, w_t60 = want[c][r][k]
                assert (int(rng[i, 0]), int(rng[i, 1])) == (first, real)
                n = data.shape[0]
                scale = np.abs(data).max()
                assert np.abs(hist[i, :n] - data).max() <= 2e-7 * scale
                assert abs(float(t60[i]) - w_t60) <= 1.5 / 44100.0
                i += 1


def test_post_chain_on_empty_and_single_sample_tracks(ob):
    """Edge cases of the FloatBuffer bookkeeping: a recorder no ray reached (real_length 0 -> length 1, T60 0) and
    a track with one touched bin (the last touched bin is outside [first_sample, real_length))."""
    sc = common.named_scene("rt60")
    scene = api.Scene.from_def(sc)
    recs = [[api.Recorder((1.0, 0.0, 1.0))], [api.Recorder((1.0, 0.0, 1.0), stereo=True, right_ear=(0.0, 1.0, 0.0))]]
    n_bins = 5000
    empty = api.Track(np.zeros(n_bins, np.float32), api.FIRST_SAMPLE_INIT, 0)
    one = api.Track(np.zeros(n_bins, np.float32), 100, 100)
    one.data[100] = 0.5
    two = api.Track(np.zeros(n_bins, np.float32), 40, 43)
    two.data[40:44] = (0.25, -0.125, 0.01, 0.02)
    res = api.RenderResult([[[empty]], [[one, two]]], 0, 0, 0, 0, 0, 0, 0.0)
    want_max, want = ob.post_all(res.tracks)
    got_max, hist, rng, t60, tmax, tpr = _device_post(scene, res, recs, 2)
    assert tpr == 2
    assert got_max == pytest.approx(want_max, rel=2e-7)
    rows = {0: want[0][0][0], 2: want[1][0][0], 3: want[1][0][1]}
    for i, (data, first, real, w_t60) in rows.items():
        assert (int(rng[i, 0]), int(rng[i, 1])) == (first, real)
        assert np.abs(hist[i, : data.shape[0]] - data).max() <= 2e-7
        assert abs(float(t60[i]) - w_t60) <= 1.5 / 44100.0
