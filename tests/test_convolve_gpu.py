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


FFT_TOL = 2e-5    # |fft - direct| <= FFT_TOL * peak of the direct result (float32 transforms of up to 2^21 points)


@pytest.mark.parametrize("n_dry,offset,first,real,fade", [(443, 0, 1028, 30000, False), (2000, 44100, 4589, 60000, False),
                                                           (1500, 88200, 2000, 50000, True), (705600, 0, 3000, 400000, False),
                                                           (100000, 1000, 1500, 300000, True)])
def test_fft_convolution_agrees_with_the_direct_form(n_dry, offset, first, real, fade):
    """ear_b200_convolve_fft (the reference's USE_FFTW form of RecorderTrack::Process, src/Recorder.cpp:145-243) against the
    bit-exact direct form on the same inputs, including BASELINE config 3's 705 600-sample dry signal and the keyframe
    cross-fade (faded dry signals in the FFT form == interpolated responses in the direct form)."""
    rng = np.random.default_rng(n_dry + first)
    a = _track(rng, first, real)
    b = _track(rng, max(0, first - 400), real - 7000, length=real - 2000) if fade else None
    dry = (rng.normal(size=n_dry) * np.hanning(n_dry)).astype(np.float32)
    direct = api.convolve(a, dry, offset, response2=b)
    fft = api.convolve(a, dry, offset, response2=b, fft=True)
    assert (fft.first_sample, fft.real_length) == (direct.first_sample, direct.real_length)
    assert fft.data.shape == direct.data.shape
    peak = np.abs(direct.data).max()
    assert peak > 0 and np.abs(fft.data - direct.data).max() <= FFT_TOL * peak
