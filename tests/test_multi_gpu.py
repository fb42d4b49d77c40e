"""Two-rank NCCL test of the public multi-GPU path (needs 2 GPUs; skipped on a one-GPU box):
create_replicated_scene (rank 0 builds, image broadcast, rank 1 adopts) + render_sharded (ray shards, ONE reduce,
finalise on rank 0) must give the tracks of the single-GPU ear_b200_render for the same seed."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ear_b200 import api, scenes
    from ear_b200.sharding import create_replicated_scene, render_sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    sc = scenes.example1_scene(samples=30000)
    ctxs, recs = api.contexts_from_def(sc)
    scene = create_replicated_scene(sc.triangles(), sc.triangle_materials(), sc.material_table(), device=rank)
    res = render_sharded(scene, ctxs, recs, max_bounces=40, seed=11)
    if rank == 0:
        flat = [t for c in res.tracks for r in c for t in r]
        np.savez(os.path.join(out_dir, "sharded.npz"), hist=np.stack([t.data for t in flat]),
                 first=np.array([t.first_sample for t in flat]), real=np.array([t.real_length for t in flat]),
                 seg=np.array([res.segments, res.rays, res.contributions]))
    else:
        assert res is None
    dist.barrier()
    scene.close()
    dist.destroy_process_group()


def test_replicated_scene_and_sharded_render_match_one_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from ear_b200 import api, scenes
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    sc = scenes.example1_scene(samples=30000)
    ctxs, recs = api.contexts_from_def(sc)
    one = api.Scene.from_def(sc).render(ctxs, recs, max_bounces=40, seed=11)
    flat = [t for c in one.tracks for r in c for t in r]
    assert tuple(got["seg"]) == (one.segments, one.rays, one.contributions)
    assert np.array_equal(got["first"], [t.first_sample for t in flat])
    assert np.array_equal(got["real"], [t.real_length for t in flat])
    ref = np.stack([t.data for t in flat])
    assert np.abs(got["hist"] - ref).max() <= 1e-4 * np.abs(ref).max()


def test_cli_output_does_not_depend_on_gpu_count(tmp_path):
    """18 contexts, rays sharded over 1 GPU and over 2 GPUs: the contexts carry their global stream ids, so both runs trace the
    same paths and the written wav agrees to the last bit or two of the 16-bit samples (float atomics reorder sums)."""
    import subprocess
    import wave

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from ear_b200 import scenes
    from ear_b200.earfile import SourceDef
    from tests.test_cli_animated_gpu import EAR, _read, _tone
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
    outs = []
    for gpus in ("1", "2"):
        env = dict(os.environ, EAR_SEED="9", EAR_MAX_BOUNCES="80", EAR_GPUS=gpus)
        r = subprocess.run([EAR, "render", path], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-800:]
        assert f"on {gpus} GPU(s)" in r.stdout
        outs.append(_read(sc.recorders[0].filename)[0])
    assert outs[0].shape == outs[1].shape
    assert np.abs(outs[0].astype(np.int32) - outs[1]).max() <= 2


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def test_group_render_shards_rays_and_reduces_over_peer_memory():
    """ear_b200_group_render (one process, several GPUs; src/EAR.cpp:196-207 is the seam): every GPU traces a share of
    the ray ids of every context, the partial histograms meet on GPU 0 in one sliced peer reduce.  Same paths as the
    one-GPU render (Philox keyed by context and ray id): counters and track ranges identical, bins up to float sum order.
    Covers mono + stereo, three contexts, and a ray count that does not divide by the GPU count."""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    from ear_b200 import api, scenes
    sc = scenes.example1_scene(samples=30010, stereo=True)
    ctxs, recs = api.contexts_from_def(sc)
    gpu = api.Scene.from_def(sc)
    one = gpu.render(ctxs, recs, max_bounces=60, seed=23)
    for g in sorted({2, min(n, 3), n}):
        grp = api.Group(gpu, list(range(g)))
        res = grp.render(ctxs, recs, max_bounces=60, seed=23)
        assert (res.rays, res.segments, res.occlusion_queries, res.contributions, res.bin_updates) == \
               (one.rays, one.segments, one.occlusion_queries, one.contributions, one.bin_updates), g
        for c in range(len(ctxs)):
            for k in range(2):
                a, b = res.tracks[c][0][k], one.tracks[c][0][k]
                assert (a.first_sample, a.real_length) == (b.first_sample, b.real_length)
                assert np.abs(a.data - b.data).max() <= 1e-4 * np.abs(b.data).max()
        grp.close()


def test_cli_calc_t60_uses_every_gpu_for_one_context(tmp_path):
    """`EAR calc T60` has ONE context: with whole contexts dealt to GPUs it could only ever use one GPU.  Ray sharding
    gives the same T60 on 1 and on all GPUs (same paths; the T60 estimator thresholds the summed track, so allow a
    few samples of difference from the float sum order)."""
    import re
    import subprocess
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    from ear_b200 import scenes
    from tests.test_cli_animated_gpu import EAR
    wav = scenes.write_click_wav(str(tmp_path / "click.wav"))
    sc = scenes.rt60_scene(samples=200000, wav=wav)
    path = str(tmp_path / "rt60.ear")
    sc.write(path)
    t60 = []
    for gpus in ("1", str(n)):
        r = subprocess.run([EAR, "calc", "T60", path], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, EAR_SEED="77", EAR_GPUS=gpus))
        assert r.returncode == 0, r.stdout[-500:]
        assert f"on {gpus} GPU(s)" in r.stdout
        t60.append(float(re.search(r"T60_ear\s*: ([0-9.]+)s", r.stdout).group(1)))
    assert abs(t60[0] - t60[1]) <= 5.0 / 44100.0, t60
