"""Generates the committed golden fixtures by RUNNING THE REFERENCE ITSELF (oracle/_ref, built from
the unmodified sources under /root/reference by oracle/build_ref.sh).  Run from the repo root in the
build container:  python tests/golden/make_golden.py
Outputs (small): tests/golden/golden.json + tests/golden/queries_<scene>.npz
  * first-hit: triangle index / t / hit point / flipped normal from Mesh::RayIntersection
  * occlusion: Mesh::LineIntersection answers
  * tracks: per (context, recorder, track) first_sample, real_length, SHA-256 of the raw float32 track
    and 16 probe bins, from Scene::Render with rand() seeded through time()
  * T60: the three numbers `EAR_ref calc T60 <file>` prints for a given seed
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ear_b200 import scenes  # noqa: E402
from oracle import binding as ob  # noqa: E402
from tests import common  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = {"rt60": 500, "rt60_saved": 500, "example1": 500, "soup": 500}   # file "samples" -> 50 rays / context
SEED = 20240


def track_digest(tr):
    raw = np.ascontiguousarray(tr.data[: tr.real_length + 1], np.float32)
    probes = np.linspace(tr.first_sample, tr.real_length, 16).astype(np.int64)
    return {"first_sample": int(tr.first_sample), "real_length": int(tr.real_length),
            "sha256": hashlib.sha256(raw.tobytes()).hexdigest(),
            "probe_bins": [int(i) for i in probes], "probe_bits": [int(raw[i:i + 1].view(np.uint32)[0]) for i in probes]}


def main():
    assert ob.ref_available(), "build oracle/_ref first (oracle/build_ref.sh)"
    tmp = tempfile.mkdtemp()
    wav = scenes.write_click_wav(os.path.join(tmp, "click.wav"))
    out = {"seed": SEED, "scenes": {}}
    for name, samples in SCENES.items():
        sc = common.named_scene(name)
        sc.samples = samples
        for s in sc.sources:
            s.wavs = [wav]
        path = os.path.join(tmp, name + ".ear")
        sc.write(path)
        o, d = common.make_rays(sc, 4000, seed=21)
        idx, t, p, n = ob.ref_first_hit(path, o, d, tmp)
        sp, sx = common.make_segments(sc, 4000, seed=22)
        occ = ob.ref_occluded(path, sp, sx, tmp)
        np.savez_compressed(os.path.join(HERE, f"queries_{name}.npz"), o=o, d=d, idx=idx, t=t, p=p, n=n, sp=sp, sx=sx, occ=occ)
        tracks, headers, stats = ob.ref_render(path, SEED, os.path.join(tmp, name + ".tracks"))
        entry = {"samples": samples, "headers": [list(h) for h in headers], "segments": int(stats["segments"]),
                 "tracks": [[[track_digest(tr) for tr in rec] for rec in ctx] for ctx in tracks]}
        if name.startswith("rt60"):
            entry["calc_t60"] = list(ob.ref_calc_t60(path, SEED))
        out["scenes"][name] = entry
        print(name, "segments", entry["segments"], entry.get("calc_t60"))
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
