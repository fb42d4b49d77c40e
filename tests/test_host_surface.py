"""Host surface of this repo's `EAR` against the REFERENCE's own host code, on the CPU (VERDICT r1 "missing" 9, J1,
(f)3, (f)4): both binaries -- ear_b200/csrc/EAR and oracle/_ref/EAR_ref_gpu (the reference's main / parser / equalizer
/ post chain / convolution / merge / WAV writer with INTEGRATION.md's binding in place of the thread fan-out) -- are
run with tests/host_emul/libabi_on_oracle.so preloaded, a test-only implementation of the C ABI on the CPU oracle, so
they receive IDENTICAL tracks; every file they write must then be byte-identical:

  out*.wav                           merge / normalise / save          src/EAR.cpp:357-386, lib/wave/WaveFile.cpp:190-256
  debug/sound-S.band-B*.wav          band split of the dry signal      lib/equalizer/Equalizer.cpp:27-96
  debug/response-R.sound-S...{wav,bin}  Power / Truncate / dumps        src/EAR.cpp:209-241, src/Recorder.cpp:76-123
  debug/rec-R.sound-S...wav          RecorderTrack::Process bookkeeping src/Recorder.cpp:247-292,343-363

and `calc T60` must print the same three numbers.  (The same comparison with the real GPU library underneath is
tests/test_dropin_gpu.py.)  Needs oracle/_ref (built where /root/reference exists; it travels with the repo)."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from ear_b200 import scenes
from ear_b200.earfile import RecorderDef, SourceDef
from tests.test_cli_animated_gpu import _tone

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EAR = os.path.join(ROOT, "ear_b200", "csrc", "EAR")
REF = os.path.join(ROOT, "oracle", "_ref", "EAR_ref_gpu")
SHIM_SRC = os.path.join(ROOT, "tests", "host_emul", "abi_on_oracle.cpp")
SHIM = os.path.join(ROOT, "tests", "host_emul", "libabi_on_oracle.so")


@pytest.fixture(scope="module")
def shim(oracle_lib):
    if not (os.path.exists(EAR) and os.path.exists(REF)):
        pytest.skip("EAR / oracle/_ref/EAR_ref_gpu not built")
    deps = [SHIM_SRC, os.path.join(ROOT, "include", "ear_b200.h")]
    if not os.path.exists(SHIM) or any(os.path.getmtime(d) > os.path.getmtime(SHIM) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SHIM, SHIM_SRC, "-L", os.path.join(ROOT, "oracle"),
                        "-lear_oracle", "-Wl,-rpath,$ORIGIN/../../oracle"], check=True)
    return SHIM


def _both(tmp_path, sc, verb, env, shim):
    out = {}
    for tag, exe in (("ours", EAR), ("ref", REF)):
        d = tmp_path / tag
        (d / "debug").mkdir(parents=True)
        for k, rec in enumerate(sc.recorders):
            rec.filename = str(d / f"out{k}.wav")
        sc.debugdir = str(d / "debug") + "/"
        path = str(d / "scene.ear")
        sc.write(path)
        e = dict(os.environ, LD_PRELOAD=shim, **env)
        r = subprocess.run([exe, *verb, path], capture_output=True, text=True, env=e, timeout=900, cwd=str(d), stdin=subprocess.DEVNULL)
        out[tag] = (d, r)
    return out


def _assert_same_tree(a_dir, b_dir, at_least):
    def files(d):
        return sorted(os.path.relpath(os.path.join(r, f), d) for r, _, fs in os.walk(d) for f in fs if not f.endswith(".ear"))
    fa, fb = files(a_dir), files(b_dir)
    assert fa == fb, (fa, fb)
    assert len(fa) >= at_least, fa
    for n in fa:
        a, b = open(os.path.join(a_dir, n), "rb").read(), open(os.path.join(b_dir, n), "rb").read()
        assert a == b, f"{n}: {len(a)} vs {len(b)} bytes"


def test_example1_stereo_render_is_byte_identical(tmp_path, shim):
    wav = scenes.write_click_wav(str(tmp_path / "click.wav"))
    sc = scenes.example1_scene(samples=3000, wav=wav, stereo=True)
    out = _both(tmp_path, sc, ["render"], {"EAR_SEED": "9", "EAR_MAX_BOUNCES": "60"}, shim)
    assert out["ours"][1].returncode == 0 and out["ref"][1].returncode == 0, (out["ours"][1].stdout[-400:], out["ref"][1].stdout[-400:])
    _assert_same_tree(out["ours"][0], out["ref"][0], 13)


def test_mono_and_stereo_recorders_with_air_absorption(tmp_path, shim):
    wav = scenes.write_click_wav(str(tmp_path / "click.wav"))
    sc = scenes.rt60_scene(samples=2000, wav=wav, air=(0.001, 0.002, 0.004))
    sc.recorders.append(RecorderDef("x", position=(1.0, 2.0, 1.2), stereo=True, right_ear=(0.0, 1.0, 0.0), head_size=0.25,
                                    head_absorption=(0.2, 0.4, 0.8)))
    out = _both(tmp_path, sc, ["render"], {"EAR_SEED": "3", "EAR_MAX_BOUNCES": "80"}, shim)
    assert out["ours"][1].returncode == 0 and out["ref"][1].returncode == 0
    _assert_same_tree(out["ours"][0], out["ref"][0], 2 + 3 + 2 * 9)


def test_keyframed_scene_with_triple_band_source(tmp_path, shim):
    """BASELINE config 3 in miniature: KEYS + anim blocks, a 3SRC next to an SSRC with gain and offset, moving mono
    listener: sound x keyframe x band contexts, the cross-fade convolution between successive keyframes
    (src/Recorder.cpp:267-292), SoundFile::Section offsets (src/EAR.cpp:297-327)."""
    sc = scenes.example1_scene(samples=1500, wav=_tone(str(tmp_path / "a.wav"), 440.0), stereo=False)
    sc.keys = [0.0, 0.05, 0.1]
    sc.sources[0].position = None
    sc.sources[0].animation = np.array([[-5, 5, 1.6], [-4, 5, 1.6], [-3, 5, 1.6]], np.float32)
    sc.sources.append(SourceDef([_tone(str(tmp_path / "lo.wav"), 120.0), _tone(str(tmp_path / "mid.wav"), 1000.0),
                                 _tone(str(tmp_path / "hi.wav"), 5000.0)],
                                animation=np.array([[8, -8, 1.2]] * 3, np.float32), gain=0.5, offset=0.02))
    sc.recorders[0].position = None
    sc.recorders[0].animation = np.array([[5, -5, 1.6], [5, -4, 1.6], [5, -3, 1.6]], np.float32)
    out = _both(tmp_path, sc, ["render"], {"EAR_SEED": "9", "EAR_MAX_BOUNCES": "40"}, shim)
    assert out["ours"][1].returncode == 0 and out["ref"][1].returncode == 0, (out["ours"][1].stdout[-400:], out["ref"][1].stdout[-400:])
    _assert_same_tree(out["ours"][0], out["ref"][0], 1 + 6 + 3 * 18)


def test_calc_t60_prints_the_same_numbers(tmp_path, shim):
    wav = scenes.write_click_wav(str(tmp_path / "click.wav"))
    for k, (dims, refl) in enumerate([((10.0, 6.0, 4.0), 0.95), ((5.0, 4.0, 3.0), 0.75)]):
        sc = scenes.rt60_scene(dims=dims, refl=(0.9, refl, 0.9), samples=6000, wav=wav)
        sub = tmp_path / f"room{k}"
        sub.mkdir()
        out = _both(sub, sc, ["calc", "T60"], {"EAR_SEED": str(5 + k), "EAR_MAX_BOUNCES": "400"}, shim)
        a = re.findall(r"T60_\w+\s*: ([0-9.]+)s", out["ours"][1].stdout)
        b = re.findall(r"T60_\w+\s*: ([0-9.]+)s", out["ref"][1].stdout)
        assert len(a) == 3 and a == b, (a, b)
        _assert_same_tree(out["ours"][0], out["ref"][0], 3)


def test_mesh_source_scene(tmp_path, shim):
    """A source that emits from a mesh (`mesh` sub-block inside SSRC, src/SoundFile.cpp:50-53): parsed by both readers,
    rays start on the emitter triangles, bounce 0 is recorded, no direct lobe."""
    wav = scenes.write_click_wav(str(tmp_path / "click.wav"))
    sc = scenes.rt60_scene(samples=1500, wav=wav)
    sc.sources[0].position = None
    sc.sources[0].mesh_material = sc.materials[0].name
    sc.sources[0].mesh_verts = np.array([[[-4, -1, 1], [-4, 1, 1], [-4, 0, 2.5]], [[-4, 1, 1], [-3, 1, 1], [-4, 0, 2.5]]], np.float32)
    out = _both(tmp_path, sc, ["render"], {"EAR_SEED": "12", "EAR_MAX_BOUNCES": "60"}, shim)
    assert out["ours"][1].returncode == 0 and out["ref"][1].returncode == 0, (out["ours"][1].stdout[-400:], out["ref"][1].stdout[-400:])
    _assert_same_tree(out["ours"][0], out["ref"][0], 10)


def test_error_lines_match(tmp_path, shim):
    sc = scenes.rt60_scene(samples=1000, wav=str(tmp_path / "missing.wav"))
    out = _both(tmp_path, sc, ["render"], {"EAR_SEED": "1"}, shim)
    for tag in ("ours", "ref"):
        assert out[tag][1].returncode == 1 and "Error: Failed to open sound file" in out[tag][1].stdout, out[tag][1].stdout[-300:]


def test_example2_fixture_is_the_reference_scene():
    """BASELINE config 3: 232 triangles (24 + 102 + 106 over three meshes) in two materials, 10 keyframes at
    (frame - 1) / 24 s for frames 1, 51, ..., 451, three sources x 10 keyframes x 3 bands = 90 contexts with 1e4 rays,
    animated mono listener (values recovered from example2.blend by tests/golden/make_example2.py)."""
    from ear_b200 import api
    sc = scenes.example2_scene()
    assert sc.triangles().shape == (232, 3, 3) and len(sc.materials) == 2
    assert [m.name for m in sc.materials] == ["Material", "Ground floor"]
    assert np.allclose(sc.materials[0].refl, (0.93, 0.97, 0.99)) and np.allclose(sc.materials[1].spec, (0.2, 0.3, 0.4))
    assert len(sc.keys) == 10 and np.allclose(sc.keys, [(f - 1) / 24.0 for f in range(1, 500, 50)])
    assert np.allclose(sc.air_absorption, (0.01, 0.03, 0.05)) and np.allclose(sc.freq, (0.4, 2.0, 6.0))
    ctxs, recs = api.contexts_from_def(sc)
    assert len(ctxs) == 90 and ctxs[0].num_samples == 10000 and len(sc.sources[0].wavs) == 3
    assert not recs[0][0].stereo and recs[0][0].position != recs[89][0].position     # the listener walks
    import tempfile
    from ear_b200.earfile import read_ear
    with tempfile.NamedTemporaryFile(suffix=".ear") as f:
        sc.write(f.name)
        back = read_ear(f.name)
    assert back.triangles().shape == (232, 3, 3) and len(back.keys) == 10 and len(back.sources) == 3


def test_example2_render_is_byte_identical(tmp_path, shim):
    """example2 through both host surfaces (reduced ray count, short dry signals): 90 contexts, the keyframe cross-fade
    between every pair of successive keyframes, a 3SRC and two SSRC sources, animated listener."""
    wavs = [_tone(str(tmp_path / f"w{k}.wav"), f, n=3000) for k, f in enumerate((110.0, 900.0, 4000.0, 300.0, 1500.0))]
    sc = scenes.example2_scene(samples=600, bach=wavs[:3], steps=wavs[3], door=wavs[4])
    out = _both(tmp_path, sc, ["render"], {"EAR_SEED": "21", "EAR_MAX_BOUNCES": "25"}, shim)
    assert out["ours"][1].returncode == 0 and out["ref"][1].returncode == 0, (out["ours"][1].stdout[-400:], out["ref"][1].stdout[-400:])
    _assert_same_tree(out["ours"][0], out["ref"][0], 1 + 9 + 3 * 90)
