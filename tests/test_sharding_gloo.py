"""world_size-2 gloo test of the multi-rank path: shard bounds, raw partial histograms, ONE reduce,
ranges by min/max.  The tracer here is the oracle (the CUDA kernels need a GPU; the same check runs
against them in tests/test_gpu_parity.py::test_multiple_recorders_and_raw_shards_add_up)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_exactly_once():
    from ear_b200.sharding import shard_bounds
    for n in (0, 1, 7, 12500000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ear_b200 import api, scenes
    from ear_b200.sharding import reduce_partials, shard_bounds
    from oracle import binding as ob
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = scenes.example1_scene(samples=4000)
    cpu = ob.OracleScene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc)
    lo, hi = shard_bounds(ctxs[0].num_samples, rank, world)
    tracks, cnt = cpu.render(ctxs, recs, max_bounces=40, seed=9, first_ray=lo, ray_count=hi - lo, finalise=False)
    flat = [t for c in tracks for r in c for t in r]
    n_bins = max(t.data.shape[0] for t in flat)
    n_bins_t = torch.tensor([n_bins]); dist.all_reduce(n_bins_t, op=dist.ReduceOp.MAX); n_bins = int(n_bins_t.item())
    hist = torch.zeros((len(flat), n_bins), dtype=torch.float32)
    for i, t in enumerate(flat):
        hist[i, : t.data.shape[0]] = torch.from_numpy(t.data)
    first = torch.tensor([t.first_sample for t in flat], dtype=torch.int32)
    real = torch.tensor([t.real_length for t in flat], dtype=torch.int32)
    seg = torch.tensor([cnt["segments"], cnt["rays"]], dtype=torch.int64)
    reduce_partials(hist, first, real, dst=0)
    dist.reduce(seg, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), hist=hist.numpy(), first=first.numpy(), real=real.numpy(), seg=seg.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_matches_unsharded(tmp_path):
    import torch.multiprocessing as mp
    from ear_b200 import api, scenes
    from oracle import binding as ob
    ob.build()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "reduced.npz")
    sc = scenes.example1_scene(samples=4000)
    cpu = ob.OracleScene.from_def(sc)
    ctxs, recs = api.contexts_from_def(sc)
    tracks, cnt = cpu.render(ctxs, recs, max_bounces=40, seed=9, finalise=False)
    flat = [t for c in tracks for r in c for t in r]
    assert int(got["seg"][0]) == cnt["segments"] and int(got["seg"][1]) == cnt["rays"]
    for i, t in enumerate(flat):
        assert int(got["first"][i]) == t.first_sample and int(got["real"][i]) == t.real_length
        n = t.real_length + 1
        ref = t.data[:n].astype(np.float64)
        assert np.abs(got["hist"][i, :n] - ref).max() <= 1e-5 * np.abs(ref).max()
