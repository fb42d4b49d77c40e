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
