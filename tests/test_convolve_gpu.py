"""RecorderTrack::Process (direct convolution, src/Recorder.cpp:247-292) on the GPU against
the oracle's restatement.  The kernel adds every output sample's terms in the reference's order with one float
multiply and one float add per term, so the comparison is bit-exact."""
import numpy as np
import pytest

from ear_b200 import api

pytestmark = pytest.mark.gpu


def _track(rng, first, real, length=None, scale=1.0):
    length = length or max(3 * 44100, real + 44100)
    data = np.zeros(length, np.float32)
    data[first:real + 1] = (rng.normal(size=real + 1 - first) * scale * np.exp(-np.arange(real + 1 - first) / 5000.0)).astype(np.float32)
    return api.Track(data, first, real)


@pytest.mark.parametrize("n_dry,offset,first,real", [(443, 0, 1028, 30000), (1, 5, 0, 10), (2000, 44100, 4589, 60000),
                                                      (300, 0, 132299, 140000), (777, 123, 500, 500)])
def test_convolution_bit_exact(n_dry, offset, first, real):
    from oracle import binding as ob
    rng = np.random.default_rng(n_dry + first)
    resp = _track(rng, first, real)
    dry = rng.normal(size=n_dry).astype(np.float32)
    got = api.convolve(resp, dry, offset)
    want = ob.convolve(resp, dry, offset)
    assert got.data.shape == want.shape
    assert np.array_equal(got.data.view(np.uint32), want.view(np.uint32))
    if real > first:
        assert got.first_sample == min(132299, offset + first) and got.real_length == n_dry - 1 + offset + real - 1
    else:
        assert (got.first_sample, got.real_length) == (132299, 0)      # the reference's loops do not run


def test_keyframe_crossfade_bit_exact():
    from oracle import binding as ob
    rng = np.random.default_rng(3)
    a = _track(rng, 2000, 50000)
    b = _track(rng, 1500, 42000, length=45000)          # shorter allocation: reads as zero beyond it
    dry = rng.normal(size=1500).astype(np.float32)
    got = api.convolve(a, dry, 88200, response2=b)
    want = ob.convolve(a, dry, 88200, response2=b)
    assert np.array_equal(got.data.view(np.uint32), want.view(np.uint32))
    assert got.first_sample == min(132299, 88200 + 1500) and got.real_length == 1499 + 88200 + 50000 - 1


def test_matches_numpy_convolve_numerically():
    rng = np.random.default_rng(4)
    resp = _track(rng, 100, 20000)
    dry = rng.normal(size=500).astype(np.float32)
    got = api.convolve(resp, dry, 0)
    ref = np.convolve(dry.astype(np.float64), resp.data[:20000].astype(np.float64))   # [first, real_length) only
    n = ref.shape[0]
    assert np.abs(got.data[:n] - ref).max() <= 1e-4 * np.abs(ref).max()
