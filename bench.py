#!/usr/bin/env python
"""bench.py -- ray-bounce segments/sec of the Scene::Render replacement on B200.

Workloads:
  c4 (default, the headline; BASELINE.json configs[3]): synthetic 1M-triangle hall (60 x 40 x 20 m, displaced wall
     tessellation + 2000 box obstacles, 4 materials), 8 frequency bands = 8 contexts, 1e8 rays in total (1.25e7 per
     band), max 50 bounces, one mono recorder, point source.
  c5 (--workload c5; configs[4]): 10M-triangle complex of 8 coupled halls, 3 bands, 1e9 rays, 64 mono recorders.
A "step" is one full render of that ray budget.  Total work is fixed as N grows (strong scaling): rank g traces ray ids
[g*R/N, (g+1)*R/N) of every context into its own partial histogram, then ONE NCCL reduce (sum) of the histograms
(+ min/max of the track ranges) to rank 0, which finalises.  No other collective is on the data path.

  value : segments / second, whole job, scene + contexts already resident in HBM (ear_b200_trace_device +
          ear_b200_finalise_device on torch's current stream, CUDA events, max over ranks)
  e2e   : same metric through the public API with host buffers, every step: sharding.create_replicated_scene
          (triangle upload + device BVH build on every rank) + sharding.render_sharded (context upload, visibility
          maps, trace, one reduce, finalise, track download); at N=1 that is ear_b200_scene_create + ear_b200_render
  roofline : c4 -- the closest-hit traversal kernel; c5 -- the occlusion-query kernels (visibility-map lookups + BVH
          any-hit).  Duration live from the library's per-launch CUDA events; traffic and issue statistics from the
          committed ncu capture of the same configuration (profiles/r2_traffic.json, scripts/capture_traffic.sh)
  --impl reference : the reference's own CPU render (oracle/_ref/ref_harness, built from the unmodified sources) on
          this host's cores, one 50-ray context per process, time-boxed (50 is the reference's minimum:
          DrawProgressBar divides by samples/50).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAX_BOUNCES = 50
# workload -> (triangles, bands, rays, recorders); C4 is the headline (BASELINE.json configs[3]), C5 = configs[4]
WORKLOADS = {"c4": (1_000_000, 8, 1e8, 1), "c5": (10_000_000, 3, 1e9, 64)}
N_TRIS, N_BANDS, TOTAL_RAYS, N_RECORDERS, WORKLOAD, WORKLOAD_KEY = 0, 0, 0, 0, "", "c4"


def select_workload(key, rays=None, tris=None):
    global N_TRIS, N_BANDS, TOTAL_RAYS, N_RECORDERS, WORKLOAD, WORKLOAD_KEY
    t, b, r, n = WORKLOADS[key]
    WORKLOAD_KEY = key
    N_TRIS = int(float(tris if tris is not None else os.environ.get("EAR_BENCH_TRIS", t)))
    N_BANDS, N_RECORDERS = b, n
    TOTAL_RAYS = int(float(rays if rays is not None else os.environ.get("EAR_BENCH_RAYS", r)))
    what = "hall" if key == "c4" else "complex of 8 coupled halls"
    WORKLOAD = (f"synthetic {N_TRIS}-triangle {what}, {N_BANDS} bands, {TOTAL_RAYS:.3g} rays, {MAX_BOUNCES} bounces, "
                f"{N_RECORDERS} mono recorder{'s' if N_RECORDERS > 1 else ''}")
    if TOTAL_RAYS != int(r) or N_TRIS != t:
        WORKLOAD += f" [REDUCED from the named {t} triangles / {r:.0e} rays]"


def algorithmic_bytes(n_tris, segments, occlusion, bin_updates):
    """64*S + Q(T)*(S+O) + 8*U with Q(T) = 32*ceil(log2(ceil(T/4))) + 192."""
    depth = math.ceil(math.log2(max(2, math.ceil(n_tris / 4))))
    q = 32 * depth + 192
    return 64 * segments + q * (segments + occlusion) + 8 * bin_updates


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower() == "active" for r in self.rows if len(r) >= 6)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def build_workload():
    import numpy as np
    from ear_b200 import api, scenes
    if WORKLOAD_KEY == "c5":
        sc, table = scenes.synthetic_complex(n_tris=N_TRIS, n_obstacles=max(8, N_TRIS // 500), n_bands=N_BANDS, seed=0,
                                             n_recorders=N_RECORDERS)
    else:
        sc, table = scenes.synthetic_hall(n_tris=N_TRIS, n_obstacles=max(1, N_TRIS // 500), n_bands=N_BANDS, seed=0)
    af = scenes.air_factors(N_BANDS)
    rays_per_ctx = TOTAL_RAYS // N_BANDS
    ctxs = [api.Context(b, rays_per_ctx, float(af[b]), sc.sources[0].position, 1.0, 1.0) for b in range(N_BANDS)]
    recs = [api.Recorder(r.position) for r in sc.recorders[:N_RECORDERS]]
    return sc, np.ascontiguousarray(table, np.float32), ctxs, recs


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from ear_b200 import api
    from ear_b200.sharding import reduce_partials, shard_bounds

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (ear_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = api.load_library()

    sc, table, ctxs, recs = build_workload()
    verts, tri_mat = sc.triangles(), sc.triangle_materials()
    scene = api.Scene(verts, tri_mat, table, device=local)
    n_ctx, n_rec = len(ctxs), len(recs)
    ctx_c = api.pack_contexts(ctxs)
    rec_c, _ = api.pack_recorders(recs, n_ctx)
    rays_per_ctx = ctxs[0].num_samples
    lo, hi = shard_bounds(rays_per_ctx, rank, world)
    opt = api.make_options(max_bounces=MAX_BOUNCES, seed=1234, first_ray=lo, ray_count=hi - lo, finalise=False)
    opt_warm = opt
    if args.warmup_rays is not None:   # C5 at full size: a warm-up step of 1e9 rays costs as much as the timed one
        wlo, whi = shard_bounds(max(1, int(float(args.warmup_rays)) // n_ctx), rank, world)
        opt_warm = api.make_options(max_bounces=MAX_BOUNCES, seed=1234, first_ray=wlo, ray_count=whi - wlo, finalise=False)
    n_bins = scene.default_bins(opt)
    n_tracks = n_ctx * n_rec * api.tracks_per_recorder(rec_c)
    hist = torch.zeros((n_tracks, n_bins), dtype=torch.float32, device=dev)
    rng_first = torch.empty((n_tracks,), dtype=torch.int32, device=dev)
    rng_real = torch.empty((n_tracks,), dtype=torch.int32, device=dev)
    rng = torch.empty((n_tracks, 2), dtype=torch.int32, device=dev)
    counters = torch.zeros((8,), dtype=torch.int64, device=dev)
    flush = torch.empty((64 * 1024 * 1024,), dtype=torch.float32, device=dev)   # 256 MiB > 126 MB L2
    if os.environ.get("EAR_BENCH_STREAM") == "new":     # diagnosis: the device-timed leg on a stream of its own
        torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def step(opt=opt):
        flush.zero_()
        hist.zero_()
        counters.zero_()
        rng[:, 0] = api.FIRST_SAMPLE_INIT
        rng[:, 1] = 0
        api._check(lib, lib.ear_b200_trace_device(scene.handle, ctx_c, n_ctx, rec_c, n_rec, C.byref(opt), n_bins,
                                                  hist.data_ptr(), rng.data_ptr(), counters.data_ptr(), sp))
        if world > 1:
            # one reduce of the partial histograms over NVLink; track ranges reduce by min / max
            rng_first.copy_(rng[:, 0]); rng_real.copy_(rng[:, 1])
            reduce_partials(hist, rng_first, rng_real, dst=0)
            rng[:, 0] = rng_first; rng[:, 1] = rng_real
        if rank == 0:
            api._check(lib, lib.ear_b200_finalise_device(scene.handle, ctx_c, n_ctx, rec_c, n_rec, n_bins,
                                                         hist.data_ptr(), rng.data_ptr(), sp))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(opt_warm)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    scene.stats(reset=True)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seg_total = occ_total = bins_total = 0
    t0.record(stream)
    for _ in range(args.steps):
        step()
        c = counters.cpu().numpy()
        seg_total += int(c[1]); occ_total += int(c[2]); bins_total += int(c[4])
    t1.record(stream)
    barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    tot = torch.tensor([seg_total, occ_total, bins_total, int(counters[5].item())], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    clocks = sampler.stop() if sampler else None
    st = scene.stats()
    n_launch = sum(st["launches"].values())
    total_ms = float(ms.item())
    segments, occl, bins, dropped = (int(x) for x in tot.tolist())
    value = segments / (total_ms * 1e-3)
    if rank == 0 and os.environ.get("EAR_BENCH_VERBOSE"):
        print(f"[bench] device-timed: {value:.4e} segments/s, {total_ms / args.steps:.1f} ms/step, kernels {st['ms']}", file=sys.stderr)

    # ---- e2e: host buffers through the public C ABI, every step ----
    e2e_ms = []
    e2e_create_ms = []
    h2d = verts.nbytes + tri_mat.nbytes + table.nbytes + C.sizeof(ctx_c) + C.sizeof(rec_c)
    d2h = 0
    e2e_segments = 0
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps if args.e2e_steps is None else args.e2e_steps, 5))
    from ear_b200.sharding import create_replicated_scene, render_sharded
    e2e_warm = 1 if (e2e_steps and args.warmup > 0 and args.warmup_rays is None) else 0   # one untimed pass: page-locked buffers, allocator caches
    for it in range(e2e_warm + e2e_steps):
        timed = it >= e2e_warm
        barrier()
        w0 = time.perf_counter()
        # the public multi-GPU path: one BVH build (rank 0), image broadcast over NVLink, sharded trace, one reduce
        s2 = create_replicated_scene(verts, tri_mat, table, device=local)
        torch.cuda.synchronize()
        wc = time.perf_counter()
        res = render_sharded(s2, ctxs, recs, max_bounces=MAX_BOUNCES, seed=1234)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        s2.close()
        if res is not None and timed:
            d2h = sum(t.data.nbytes for c in res.tracks for r in c for t in r)   # whole track buffers come down
            e2e_segments += res.segments
        res = None            # hands the page-locked track block back before the next step asks for one
        if not timed:
            continue
        e2e_create_ms.append((wc - w0) * 1e3)
        e2e_ms.append((w1 - w0) * 1e3)
        if os.environ.get("EAR_BENCH_VERBOSE"):
            print(f"[bench] rank {rank} e2e step: create {e2e_create_ms[-1]:.1f} ms, render {(w1 - wc) * 1e3:.1f} ms", file=sys.stderr)
            if os.environ.get("EAR_BENCH_VERBOSE") == "2":   # diagnosis only: the same render again on the now-warm scene
                s3 = create_replicated_scene(verts, tri_mat, table, device=local)
                for k in range(3):
                    s3.stats(reset=True)
                    torch.cuda.synchronize(); wa = time.perf_counter()
                    render_sharded(s3, ctxs, recs, max_bounces=MAX_BOUNCES, seed=1234)
                    torch.cuda.synchronize()
                    print(f"[bench] rank {rank} render #{k + 1} on one scene: {(time.perf_counter() - wa) * 1e3:.1f} ms, "
                          f"kernels {({n: round(v, 1) for n, v in s3.stats()['ms'].items()})}", file=sys.stderr)
                s3.close()
    e2 = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    es = torch.tensor([e2e_segments], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
    e2e_value = float(es.item()) / (float(e2.item()) * 1e-3) if e2e_steps else None

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_base = reference_sample(sc)
        except Exception as exc:  # the baseline is reported, never required
            cpu_base = {"error": str(exc)[:200]}
    if rank == 0:
        peak, which = measured_peak()
        # Dominant kernel class of the step and its SURVEY 8(d) per-unit bytes: C4 -- the closest-hit traversal, Q(T) per
        # segment; C5 -- the occlusion queries (visibility-map lookups + BVH any-hit fallback), Q(T) per query.
        # Duration: CUDA events recorded by the library around every launch on the launch stream, summed over the
        # timed region (this rank's launches).
        depth = math.ceil(math.log2(max(2, math.ceil(N_TRIS / 4))))
        q_bytes = 32 * depth + 192
        if WORKLOAD_KEY == "c5":
            dom_name, units = "wf_vismap_kernel + wf_traverse_kernel<true> (occlusion queries)", occl / world
            dom_ms = st["ms"]["vismap"] + st["ms"]["anyhit"]
        else:
            dom_name, units = "wf_traverse_kernel<false> (closest hit)", segments / world
            dom_ms = st["ms"]["closest"]
        dom_launches = st["launches"]["closest"]      # one launch of every class per iteration
        dom_bytes = q_bytes * units
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        # DRAM bytes and issue statistics of ONE mid-step launch of that kernel on this configuration, from the committed
        # ncu --set full capture (scripts/capture_traffic.sh writes the entry; never measured inside a bench run)
        traffic, ncu_info = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tj = json.load(f).get(f"{WORKLOAD_KEY}_n{world}")
            if tj:
                traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
                ncu_info = {k: tj[k] for k in tj if k not in ("dram_bytes_read", "dram_bytes_write")}
        except Exception:
            traffic = None
        whole = algorithmic_bytes(N_TRIS, segments / world, occl / world, bins / world) / (total_ms * 1e-3) / 1e9
        line = {
            "metric": "ray-bounce segments/sec", "value": value, "unit": "segments/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, **({"warmup_rays": int(float(args.warmup_rays))} if args.warmup_rays is not None else {}), "triangles": N_TRIS, "bands": N_BANDS, "rays": TOTAL_RAYS,
                       "max_bounces": MAX_BOUNCES, "recorders": N_RECORDERS, "bins_per_track": n_bins,
                       "sharding": f"ray ranges over {world} GPU(s), one NCCL reduce",
                       "l2": "flushed by a 256 MiB memset before every step"},
            "segments_per_step": segments // args.steps, "occlusion_queries_per_step": occl // args.steps,
            "bin_updates_per_step": bins // args.steps, "dropped_updates": dropped,
            "e2e": {"value": e2e_value, "unit": "segments/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "includes": "scene upload + BVH build + image broadcast + visibility maps + trace + reduce + finalise + track download",
                    "steps": e2e_steps, "warmup": e2e_warm,
                    "ms_per_step": sum(e2e_ms) / max(1, len(e2e_ms)), "scene_create_ms": sum(e2e_create_ms) / max(1, len(e2e_create_ms)),
                    "step_ms_rank0": [round(x, 1) for x in e2e_ms], "scene_create_ms_rank0": [round(x, 1) for x in e2e_create_ms]},
            "gpu_launches": n_launch, "kernel_ms_per_step": {k: v / args.steps for k, v in st["ms"].items()},
            "launches_per_step": {k: v / args.steps for k, v in st["launches"].items()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of one "
                                                             "mid-step launch on this configuration, profiles/r2_traffic.json)",
                         "ncu": ncu_info,
                         "algorithmic_bytes_per_launch": dom_bytes / max(1, dom_launches),
                         "kernel": dom_name,
                         "kernel_ms": dom_ms / max(1, dom_launches), "kernel_launches": dom_launches,
                         "kernel_share_of_step": dom_ms / total_ms,
                         "whole_step_achieved": whole, "whole_step_frac": whole / peak,
                         "peak_source": which,
                         **({"note": "frac > 1 is possible here: SURVEY 8(d) prices every occlusion query at a BVH walk (Q(T) bytes); "
                                     "the visibility maps answer most of them from a short triangle list instead"} if WORKLOAD_KEY == "c5" else {}),
                         "bytes_model": "dominant kernel: (32*ceil(log2(ceil(T/4)))+192) B per query it answers (SURVEY 8d Q(T)); "
                                        "whole step: 64*S + Q(T)*(S+O) + 8*U"},
            "cpu_baseline": cpu_base, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ------------------------------------------------------------------------------------------
def reference_sample(sc, procs=None, budget=20):
    """The reference's CPU render of the same scene on this host: `procs` concurrent processes, each running
    one mid-band context through the unmodified Scene::Render (1000-bounce cap, 1e-8 cutoff) for `budget`
    seconds (brute force over 1M triangles needs minutes for the reference's 50-ray minimum; the harness
    reports the segments finished when the budget expires)."""
    from ear_b200 import scenes
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        raise RuntimeError("oracle/_ref/ref_harness missing (build it where /root/reference exists)")
    cores = os.cpu_count() or 1
    procs = procs or cores
    tmp = tempfile.mkdtemp(prefix="ear_ref_")
    scenes.write_click_wav(os.path.join(tmp, "click.wav"))
    sc.sources[0].wavs = [os.path.join(tmp, "click.wav")]
    sc.samples = 500   # -> 50 rays per context, the reference's minimum (DrawProgressBar: samples/50)
    path = os.path.join(tmp, "hall.ear")
    sc.write(path)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([harness, "render", path, str(100 + i), os.path.join(tmp, f"t{i}.bin"), "t60", "threads=1",
                            f"budget={budget}"],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for i in range(procs)]
    outs = [p.communicate()[0] for p in ps]
    wall = time.perf_counter() - t0
    import re
    seg, secs = 0.0, 0.0
    for o in outs:
        m = re.search(r"REF_RENDER .*seconds=([0-9.]+) .*segments=([0-9.]+)", o)
        if not m:
            raise RuntimeError("reference run failed")
        secs = max(secs, float(m.group(1)))
        seg += float(m.group(2))
    for f in os.listdir(tmp):
        os.unlink(os.path.join(tmp, f))
    os.rmdir(tmp)
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            model = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), "unknown")
    except OSError:
        pass
    return {"value": seg / secs, "unit": "segments/s", "cores": procs, "cpu_model": model, "nproc": cores, "kind": "reference",
            "segments": int(seg), "render_seconds": secs,
            "sample": f"{procs} processes x 1 context, {budget} s budget each (reference's 1000-bounce loop), {int(seg)} segments, "
                      f"{secs:.1f} s render, {wall:.1f} s wall incl. parse"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sc, table, ctxs, recs = build_workload()
    # each sample is a time-boxed run of the reference; the whole arm stays within ~3 minutes whatever --steps is
    samples = args.steps + min(args.warmup, 1)
    budget = max(4, min(20, int(160 / max(1, samples))))
    for _ in range(min(args.warmup, 1)):
        reference_sample(sc, procs=max(1, (os.cpu_count() or 1)), budget=budget)
    t0 = time.perf_counter()
    seg = secs = 0.0
    last = None
    for _ in range(args.steps):
        last = reference_sample(sc, budget=budget)
        seg += last["segments"]
        secs += last["render_seconds"]
    wall = time.perf_counter() - t0
    value = seg / secs          # all steps: segments traced / time the slowest process of each step rendered
    line = {"impl": "reference", "metric": "ray-bounce segments/sec", "value": value, "unit": "segments/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "triangles": N_TRIS, "bands": N_BANDS, "rays": TOTAL_RAYS,
                       "max_bounces": MAX_BOUNCES, "recorders": N_RECORDERS},
            "cpu_baseline": last,
            "e2e": {"value": value, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS), help="c4 (headline, default) or c5")
    ap.add_argument("--rays", default=None, help="override the workload's ray budget (the line then says REDUCED)")
    ap.add_argument("--tris", default=None, help="override the workload's triangle count (the line then says REDUCED)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--warmup-rays", default=None, help="ray budget of the untimed warm-up steps (default: the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the end-to-end leg (default: min(steps, 5))")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the host-buffer end-to-end leg")
    args = ap.parse_args()
    select_workload(args.workload, args.rays, args.tris)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
