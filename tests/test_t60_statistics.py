"""T60 of the RT60 testbench room: the GPU's generator (Philox) against the reference's (rand()), both through
the oracle (which is bit-identical to the reference in rand mode and path-identical to the GPU in Philox mode).
North-star bar: T60 within 2 %.  The estimator is an extreme-value statistic with a single-run sigma of ~3 % at
this ray budget, so the bar is applied to the means over M independent seeds per side."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = 16
SAMPLES = 60000     # file value -> 6000 rays per run


def _one(args):
    mode, seed = args
    sys.path.insert(0, ROOT)
    from ear_b200 import api, scenes
    from oracle import binding as ob
    sc = scenes.rt60_scene(samples=SAMPLES)
    cpu = ob.OracleScene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc, t60_only=True)
    tracks, _ = cpu.render(ctxs, recs, rng_mode=mode, seed=seed)
    return ob.post_t60(tracks)


def test_t60_mean_philox_vs_rand_within_two_percent():
    from oracle import binding as ob
    ob.build()
    jobs = [(ob.RNG_RAND, 1000 + i) for i in range(M)] + [(ob.RNG_PHILOX, 2000 + i) for i in range(M)]
    with mp.get_context("spawn").Pool(min(8, os.cpu_count() or 1)) as pool:
        vals = pool.map(_one, jobs)
    rand, philox = np.array(vals[:M]), np.array(vals[M:])
    rel = abs(rand.mean() - philox.mean()) / rand.mean()
    print(f"T60 rand {rand.mean():.4f} +- {rand.std(ddof=1):.4f}   philox {philox.mean():.4f} +- {philox.std(ddof=1):.4f}   diff {100 * rel:.2f} %")
    assert rel < 0.02
    # both sit where the reference's own runs sit: between Norris-Eyring (3.039 s) and Sabine (3.118 s), +- 8 %
    for v in (rand.mean(), philox.mean()):
        assert 2.80 < v < 3.40
