"""BASELINE config 3 in miniature: keyframed source and listener (KEYS + anim blocks), a triple-band source (3SRC)
next to a single-wav one (SSRC, band-split by the equaliser), mono listener -- through the C++ host CLI.  Checks the
context count (sound x keyframe x band), the keyframe cross-fade convolution path, and that the run is
reproducible for a seed."""
import os
import re
import subprocess
import wave

import numpy as np
import pytest

from ear_b200 import scenes
from ear_b200.earfile import SourceDef

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EAR = os.path.join(ROOT, "ear_b200", "csrc", "EAR")


def _tone(path, freq, n=4000):
    t = np.arange(n) / 44100.0
    pcm = np.round(np.sin(2 * np.pi * freq * t) * np.hanning(n) * 20000).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(44100); w.writeframes(pcm.tobytes())
    return path


def _read(path):
    with wave.open(path) as w:
        return np.frombuffer(w.readframes(w.getnframes()), "<i2").copy(), w.getnchannels()


def test_animated_two_sources_three_bands(tmp_path):
    sc = scenes.example1_scene(samples=8000, wav=_tone(str(tmp_path / "a.wav"), 440.0), stereo=False)
    sc.keys = [0.0, 0.05, 0.1]
    sc.sources[0].position = None
    sc.sources[0].animation = np.array([[-5, 5, 1.6], [-4, 5, 1.6], [-3, 5, 1.6]], np.float32)
    sc.sources.append(SourceDef([_tone(str(tmp_path / "lo.wav"), 120.0), _tone(str(tmp_path / "mid.wav"), 1000.0),
                                 _tone(str(tmp_path / "hi.wav"), 5000.0)],
                                animation=np.array([[8, -8, 1.2]] * 3, np.float32), gain=0.5, offset=0.02))
    sc.recorders[0].position = None
    sc.recorders[0].animation = np.array([[5, -5, 1.6], [5, -4, 1.6], [5, -3, 1.6]], np.float32)
    sc.recorders[0].filename = str(tmp_path / "out.wav")
    path = str(tmp_path / "anim.ear")
    sc.write(path)
    env = dict(os.environ, EAR_SEED="9", EAR_MAX_BOUNCES="80")
    r = subprocess.run([EAR, "render", path], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-800:]
    assert "Saved" in r.stdout
    first, ch = _read(sc.recorders[0].filename)
    assert ch == 1 and first.shape[0] > 4000 and np.abs(first).max() > 20000
    # same seed -> same paths; float atomics may reorder sums, so allow one LSB of the 16-bit output
    r2 = subprocess.run([EAR, "render", path], capture_output=True, text=True, env=env, timeout=600)
    second, _ = _read(sc.recorders[0].filename)
    assert second.shape == first.shape and np.abs(second.astype(np.int32) - first).max() <= 2
    # 2 sources x 3 keyframes x 3 bands = 18 contexts traced
    m = re.search(r"Traced (\d+) ray-bounce segments", r.stdout)
    assert m and int(m.group(1)) > 18 * 800 * 10
