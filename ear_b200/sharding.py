"""Ray sharding across one-process-per-GPU ranks.

The scene (triangles, BVH, materials) is replicated on every GPU; rank g traces global ray ids
[g*N/G, (g+1)*N/G) of EVERY context (the Philox stream is keyed by (seed, context, ray id), so the
union over ranks is the same set of paths whatever G is).  Each rank owns a full-size partial
histogram.  The only exchange is at the end: ONE sum-reduce of the histograms to rank 0 plus a min /
max reduce of the two track-range words, then rank 0 finalises (x 1/N, direct sound, x gain^2,
src/Scene.cpp:286-316).  Works with any torch.distributed backend (NCCL over NVLink on the GPU box,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, exhaustive, non-overlapping ray-id ranges; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_rays * rank // world, n_rays * (rank + 1) // world


def reduce_partials(hist, first_sample, real_length, dst: int = 0) -> None:
    """In-place reduce to `dst`: hist by SUM, first_sample by MIN, real_length by MAX.
    `hist` [tracks, bins] float32; `first_sample` / `real_length` [tracks] int32 (separate, contiguous)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    dist.reduce(hist, dst=dst, op=dist.ReduceOp.SUM)
    dist.reduce(first_sample, dst=dst, op=dist.ReduceOp.MIN)
    dist.reduce(real_length, dst=dst, op=dist.ReduceOp.MAX)


def create_replicated_scene(verts, tri_material, materials, device: int, src: int = 0, how: str = "auto"):
    """The scene on every rank's GPU.  The BVH is built on the device (a few ms for 1M triangles), so by default every
    rank simply builds its own copy from the host triangles (the builder is deterministic: the copies are identical).
    how="broadcast" is the form the host builder needs (EAR_B200_BUILD=host): rank `src` builds, broadcasts the device
    image (header + nodes + triangle records + materials, one contiguous buffer) over NCCL / NVLink, and the other
    ranks adopt the bytes (ear_b200_scene_create_from_image) -- one host build per job instead of one per rank."""
    import os
    import torch
    import torch.distributed as dist
    from . import api
    world = dist.get_world_size() if dist.is_initialized() else 1
    if how == "auto":
        how = "broadcast" if os.environ.get("EAR_B200_BUILD") == "host" else "each"
    if world == 1 or how == "each":
        return api.Scene(verts, tri_material, materials, device=device)
    rank = dist.get_rank()
    dev = torch.device("cuda", device)
    size = torch.zeros((2,), dtype=torch.int64, device=dev)
    scene = None
    if rank == src:
        scene = api.Scene(verts, tri_material, materials, device=device)
        size[0] = scene.image_size()
        size[1] = scene.n_bands
    dist.broadcast(size, src=src)
    n_bytes, n_bands = int(size[0].item()), int(size[1].item())
    image = torch.empty((n_bytes,), dtype=torch.uint8, device=dev)
    if rank == src:
        scene.image_write(image.data_ptr(), n_bytes)
    dist.broadcast(image, src=src)
    if rank != src:
        scene = api.Scene.from_image(image.data_ptr(), n_bytes, n_bands, device=device)
    return scene


def render_sharded(scene, contexts, recorders, max_bounces: int = 1000, seed: int = 1, n_bins: int = 0, dst: int = 0,
                   post=None):
    """Scene::Render (src/Scene.cpp:111-318) over all ranks of the default process group: every rank traces its
    ray-id range of every context into a device-resident partial histogram, the partials meet in ONE reduce on
    rank `dst`, which finalises and downloads the tracks.  Host contexts in, host tracks out (RenderResult on
    `dst`, None elsewhere).  Identical to Scene.render() when there is one rank.

    post=(exponent, divisor), e.g. (0.335, 256.0): also run Render()'s post chain (src/EAR.cpp:209-228) on the device
    before the download -- Power, global maximum, Truncate(getLength(maximum / divisor)), T60 per track; the result
    then carries `maximum` and `t60` ([context][recorder][track]) and the tracks come back compressed and truncated."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import api
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    if world == 1 and post is None:
        return scene.render(contexts, recorders, max_bounces=max_bounces, seed=seed, n_bins=n_bins)
    lib = scene.lib
    dev = torch.device("cuda", scene.device)
    n_ctx = len(contexts)
    ctx_c = api.pack_contexts(contexts)
    rec_c, n_rec = api.pack_recorders(recorders, n_ctx)
    rays = max(int(c.num_samples) for c in contexts) if n_ctx else 0
    lo, hi = shard_bounds(rays, rank, world)
    opt = api.make_options(max_bounces, n_bins, seed, lo, hi - lo, False)
    if n_bins <= 0:
        n_bins = scene.default_bins(opt)
    tpr = api.tracks_per_recorder(rec_c)          # device layout: [context][recorder][tpr][bins]
    n_tracks = n_ctx * n_rec * tpr
    import os
    import sys
    import time
    dbg = bool(os.environ.get("EAR_B200_DEBUG"))
    laps = [("start", time.perf_counter())]

    def lap(what):
        if dbg:
            torch.cuda.synchronize(dev)
            laps.append((what, time.perf_counter()))
    with torch.cuda.device(dev):
        hist = torch.zeros((n_tracks, n_bins), dtype=torch.float32, device=dev)
        rng = torch.empty((n_tracks, 2), dtype=torch.int32, device=dev)
        rng[:, 0] = api.FIRST_SAMPLE_INIT
        rng[:, 1] = 0
        counters = torch.zeros((8,), dtype=torch.int64, device=dev)
        sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        api._check(lib, lib.ear_b200_trace_device(scene.handle, ctx_c, n_ctx, rec_c, n_rec, C.byref(opt), n_bins,
                                                  hist.data_ptr(), rng.data_ptr(), counters.data_ptr(), sp))
        lap("buffers + trace")
        if world > 1:
            first = rng[:, 0].contiguous()
            real = rng[:, 1].contiguous()
            reduce_partials(hist, first, real, dst=dst)
            dist.reduce(counters, dst=dst, op=dist.ReduceOp.SUM)
            if rank != dst:
                return None
            rng[:, 0] = first
            rng[:, 1] = real
        lap("reduce")
        api._check(lib, lib.ear_b200_finalise_device(scene.handle, ctx_c, n_ctx, rec_c, n_rec, n_bins,
                                                     hist.data_ptr(), rng.data_ptr(), sp))
        maximum, t60 = 0.0, None
        if post is not None:
            exponent, divisor = post
            mx = C.c_float(0.0)
            api._check(lib, lib.ear_b200_post_power_device(scene.handle, rec_c, n_ctx, n_rec, n_bins, hist.data_ptr(),
                                                           rng.data_ptr(), exponent, C.byref(mx), None, sp))
            maximum = float(mx.value)
            threshold = float(np.float32(maximum) / np.float32(divisor))
            h_t60 = np.zeros((n_tracks,), np.float32)
            api._check(lib, lib.ear_b200_post_truncate_device(scene.handle, rec_c, n_ctx, n_rec, n_bins, hist.data_ptr(),
                                                              rng.data_ptr(), threshold, h_t60.ctypes.data, sp))
            t60 = h_t60.reshape(n_ctx, n_rec, tpr).tolist()
        h_rng = rng.cpu().numpy()
        c = counters.cpu().numpy()
        # whole rows into ONE page-locked buffer (torch's host allocator keeps it for the next call); the device rows are
        # zero beyond what was recorded, and the Track objects below are views of this buffer -- no host-side copies
        h_pinned = torch.empty((n_tracks, n_bins), dtype=torch.float32, pin_memory=True)
        h_pinned.copy_(hist, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        h_hist = h_pinned.numpy()
    lap("finalise + download")
    tracks = []
    for ci in range(n_ctx):
        per_rec = []
        for k in range(n_rec):
            pair = []
            for tr in range(2 if rec_c[ci * n_rec + k].kind == api.STEREO else 1):
                t = (ci * n_rec + k) * tpr + tr
                real_len = int(h_rng[t, 1])
                if post is not None:
                    h_hist[t, real_len + 1:] = 0.0     # the device row keeps the samples beyond the truncated length
                pair.append(api.Track(h_hist[t], int(h_rng[t, 0]), real_len))   # spans the whole buffer, like Scene.render()
            per_rec.append(pair)
        tracks.append(per_rec)
    lap("track objects")
    if dbg:
        print("[ear_b200.sharding] render_sharded: " + ", ".join(f"{w} {1e3 * (t - laps[i][1]):.1f} ms" for i, (w, t) in enumerate(laps[1:])), file=sys.stderr)
    res = api.RenderResult(tracks, int(c[0]), int(c[1]), int(c[2]), int(c[3]), int(c[4]), int(c[5]), 0.0)
    res.maximum = maximum
    res.t60 = t60
    return res
