"""BASELINE config 3 at full size through the CLI: example2 (232 triangles, 90 contexts x 1e4 rays, 1000-bounce cap) with
705 600-sample triple-band sources (the length of the reference's bach-bwv999-*.wav; synthetic content) -- the direct-form
convolution of RecorderTrack::Process (src/Recorder.cpp:247-292) at its real size.  Prints wall times per stage."""
import os
import subprocess
import sys
import tempfile
import time
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ear_b200 import scenes  # noqa: E402


def noise(path, n, seed, band):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 44100.0
    sig = np.sin(2 * np.pi * band * t) * (0.3 + 0.7 * rng.uniform(size=n)) * np.minimum(1.0, t * 4)
    pcm = np.round(sig / np.abs(sig).max() * 20000).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(44100); w.writeframes(pcm.tobytes())
    return path


tmp = tempfile.mkdtemp()
n = int(os.environ.get("C3_DRY_SAMPLES", 705600))
bach = [noise(os.path.join(tmp, f"bach-{b}.wav"), n, k, f) for k, (b, f) in enumerate((("low", 110.0), ("mid", 900.0), ("high", 4000.0)))]
steps = noise(os.path.join(tmp, "steps.wav"), 222000, 7, 300.0)        # the nine step-0x.wav files end to end
door = noise(os.path.join(tmp, "door.wav"), 88200, 8, 1500.0)
sc = scenes.example2_scene(samples=100000, bach=bach, steps=steps, door=door, out=os.path.join(tmp, "out.wav"))
path = os.path.join(tmp, "example2.ear")
sc.write(path)
t0 = time.perf_counter()
r = subprocess.run([os.path.join(ROOT, "ear_b200", "csrc", "EAR"), "render", path], capture_output=True, text=True,
                   env=dict(os.environ, EAR_SEED="1"), stdin=subprocess.DEVNULL)
dt = time.perf_counter() - t0
print(r.stdout[-600:])
print(f"EAR render example2 (90 contexts, dry {n} samples): {dt:.2f} s wall, rc {r.returncode}")
