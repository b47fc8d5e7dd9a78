#!/usr/bin/env python
"""bench.py — the reference's headline metric (Mrays/s, path segments per second) on config C2 of
BASELINE.json: the LTS example scene (Cornell box + glass cone, analytic triangles) at 1024x1024,
1024 spp, max depth 16, REFERENCE integrator semantics, COMPAT traversal (the reference's own
intersect_bvh results), float32 arithmetic, Philox4x32-10 random numbers.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c1|c2|c3|c4|c4b|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one full pass of the walk over the film (all spp of every pixel).  A ray = one path segment = one
closest-hit or shadow query of the reference's control flow (R/src/utils.py:59 + R/src/light_samples.py:48).
N > 1: weak scaling (c5: strong) — ranks take disjoint sample ranges (disjoint Philox counters), then ONE
NCCL reduce of the film inside the timed region.

The default line (N = 1, C2) also carries, as sub-records measured in the same process:
  configs.{c1,c3,c4a,c4b}  the other BASELINE configs at their full sizes, each with its own roofline and a
                           bounded-sample cpu_baseline;
  f64                      C2 in float64 — the reference's own arithmetic, the like-for-like arm of the ratio;
  cpu_baseline_reference   the UNMODIFIED reference (numba) timed in the build container by
                           tools/time_reference.py (Timing A and B) — it cannot travel.
--impl reference: the CPU oracle (C port of the reference's path, all host threads) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs.  c2 is the headline (the config the metric is quoted on); the others are reachable
# with --config and ride along as sub-records of the default line.  c5 is the strong-scaling sweep (fixed total work).
CONFIGS = {
    "c1": dict(scene="lts", width=256, height=256, spp=16, max_depth=8, semantics="reference", traversal="compat", scaling="weak"),
    "c2": dict(scene="lts", width=1024, height=1024, spp=1024, max_depth=16, semantics="reference", traversal="compat", scaling="weak"),
    "c3": dict(scene="mesh", width=1920, height=1080, spp=256, max_depth=8, semantics="iterative", traversal="lbvh2", scaling="weak"),
    "c4": dict(scene="mirror", width=512, height=512, spp=64, max_depth=256, semantics="reference", traversal="compat", scaling="weak"),
    "c4b": dict(scene="medium", width=0, height=0, spp=0, max_depth=256, semantics="implicit", traversal="none", scaling="weak",
                albedo=0.99, g=0.8, radius=20.0, photons=50_000_000),
    # NEAREST semantics (true closest hit).  On a 28-triangle scene the per-triangle-bounds scan beats both hierarchies
    "c5": dict(scene="lts", width=4096, height=4096, spp=4096, max_depth=8, semantics="reference", traversal="nearest", scaling="strong"),
}
WORKLOAD = dict(CONFIGS["c2"])
METRIC = "Mrays/s (path segments/s)"
SCENE_DESC = {"lts": "LTS example scene (Cornell box + glass cone, 28 triangles, 2000 light samples)",
              "mesh": "Cornell box + 1,000,000-triangle displaced-grid mesh (1,000,028 triangles, GPU LBVH)",
              "mirror": "closed mirror box (26 triangles, one Light; C4a)"}
RAY_ACCOUNTING = ("a ray = one intersect_bvh call of the reference's control flow (R/src/utils.py:59, R/src/light_samples.py:48), the same "
                  "count the CPU arm reports; rays_memo_fraction of them are re-traces of a ray object that has just missed (Q1) and are "
                  "answered without a traversal: traversed_mrays_per_s leaves them out")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--traversal", default=None, choices=["compat", "nearest", "lbvh", "lbvh2"])
    ap.add_argument("--scheduler", default="megakernel", choices=["megakernel", "wavefront"])
    ap.add_argument("--real", default="f32", choices=["f32", "f64"])
    ap.add_argument("--spp", type=int, default=None, help="override samples per pixel (dev only; invalidates the number)")
    ap.add_argument("--size", type=int, default=None, help="override film size (dev only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rmse", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the sub-records (other configs, float64 arm)")
    ap.add_argument("--lib", default=None, help="dev only: load a variant build of liblt_b200.so (tools/build_variant.sh)")
    a = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(CONFIGS[a.config])
    if a.traversal is None:
        a.traversal = WORKLOAD["traversal"]
    if a.spp is None:
        a.spp = WORKLOAD["spp"]
    a.width, a.height = (WORKLOAD["width"], WORKLOAD["height"]) if a.size is None else (a.size, a.size)
    return a


def scene_arrays(kind):
    from light_transport_b200 import api, compat_bvh, from_arrays, pack_scene, scenes
    if kind == "mirror":
        objs, lights, cam, depth = scenes.mirror_box(api)
    else:
        objs, lights, cam, depth = scenes.lts_scene(api)
    ps = pack_scene(objs, lights)
    nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
    if kind == "mesh":
        ps = from_arrays(scenes.heightfield(nx=1000, ny=500), base=ps)
        nodes = None
    return objs, lights, ps, nodes, cam, depth


def flops_of(c):
    """algorithmic FP32 work, SURVEY.md 8(d): F_ray = 21*N_box + 48*N_tri + F_shade, F_shade = 150 per extension ray that is
    actually shaded (memoised re-traces do no work and get no credit), 40 per shadow ray; counts from the kernel's own counters."""
    return 21.0 * c["box_tests"] + 48.0 * c["tri_tests"] + 150.0 * (c["rays_ext"] - c["rays_memo"]) + 40.0 * c["rays_shadow"]


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = []
        for k, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline(ps, nodes, cam, depth, wl, width, height, spp, traversal, sample_spp=2, pixel_stride=1):
    """the oracle (kind "port") on a bounded sample: a film of (width/stride x height/stride) pixels x `sample_spp` of the spp"""
    from oracle import oracle as O
    S = O.OracleScene(ps, nodes)
    threads = os.cpu_count() or 1
    w, h = max(1, width // pixel_stride), max(1, height // pixel_stride)
    kw = dict(width=w, height=h, spp=spp, s0=0, s1=min(sample_spp, spp), max_depth=wl["max_depth"],
              camera=cam, f_distance=depth, semantics=wl["semantics"],
              traversal="compat" if traversal == "compat" else "nearest", seed=1, n_threads=threads)
    t0 = time.perf_counter()
    _, _, c = O.render(S, **kw)
    dt = time.perf_counter() - t0
    rays = c["rays_ext"] + c["rays_shadow"]
    v = rays / dt / 1e6
    return {"value": v, "unit": "Mrays/s", "cores": threads, "value_per_core": v / threads, "kind": "port",
            "sample": f"{w}x{h} pixels x {kw['s1']} of {spp} spp ({c['paths']} paths, {rays} rays, {dt:.2f} s, "
                      f"OpenMP {threads} threads, oracle/lt_oracle.c" + (", brute force over all triangles: the reference's intersect_bvh is "
                      "undefined on this scene (Q9b)" if traversal not in ("compat",) and ps.n_tris > 1000 else "") + ")"}, dt, rays


def reference_timing():
    """the unmodified numba reference, timed in the build container (it cannot travel to the GPU box)"""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r02_reference_timing.json")))
        return {"source": "profiles/r02_reference_timing.json (tools/time_reference.py, run in the build container)", "host": j["host"],
                "workload": j["config"]["workload"],
                "timing_a": {k: j["timing_a"][k] for k in ("entry", "threads", "seconds", "paths_per_s", "mrays_per_s")},
                "timing_b": {k: j["timing_b"][k] for k in ("entry", "processes", "seconds", "paths_per_s", "mrays_per_s", "mrays_per_s_per_core")}}
    except Exception as e:
        return {"unavailable": str(e)}


def time_to_rmse(R, cam, depth, args, size=256, target=0.01, chunk=32, ref_factor=16):
    """BASELINE metric, second half: wall-time until the image's RMSE against a reference image with
    16x the spp of the config (same scene, same integrator) drops below `target`; 256x256 film."""
    import torch
    n = size * size * 3
    kw = dict(width=size, height=size, max_depth=WORKLOAD["max_depth"], camera=cam, f_distance=depth, semantics=WORKLOAD["semantics"],
              traversal=args.traversal, real=args.real, scheduler=args.scheduler)
    ref_spp = WORKLOAD["spp"] * ref_factor
    film = torch.zeros(n, dtype=torch.float64, device="cuda")
    ref = torch.empty_like(film)
    R.render(spp=ref_spp, seed=999, film=film, **kw)
    R.finalize(film, ref_spp, out=ref)
    img = torch.empty_like(film)
    spp_max = WORKLOAD["spp"] * 4
    # one untimed progressive pass (render + finalize + the RMSE reduction) so that lazily loaded kernels are resident
    R.render(spp=spp_max, s0=0, s1=chunk, seed=8, film=film, **kw)
    R.finalize(film, chunk, out=img)
    float(torch.sqrt(torch.mean((img - ref) ** 2)))
    film.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done, rmse = 0, float("inf")
    while done < spp_max:                      # progressive passes share one job (spp_max) so counters stay disjoint
        R.render(spp=spp_max, s0=done, s1=done + chunk, seed=7, film=film, **kw)
        done += chunk
        R.finalize(film, done, out=img)
        rmse = float(torch.sqrt(torch.mean((img - ref) ** 2)))
        if rmse <= target:
            break
    torch.cuda.synchronize()
    return {"target_rmse": target, "seconds": time.perf_counter() - t0, "spp": done, "rmse": rmse, "film": [size, size],
            "reference_spp": ref_spp, "reached": rmse <= target,
            "note": "GPU float32 image against a GPU float32 image of 16x the spp: a convergence time, not a correctness check"}


_PEAKS = {}


def peaks(R):
    """roofline denominators measured in this process (FP32 / FP64 FMA, L2 gathers) + the driver's MEASURED_PEAKS.json"""
    if not _PEAKS:
        try:
            _PEAKS.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
            _PEAKS["hbm_source"] = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            _PEAKS.update(hbm_gbs=6650.0, hbm_source="fallback 6650 GB/s (B200_PROFILING.md)")
        _PEAKS["fp32_tflops"] = R.measure_fp32_peak()
        _PEAKS["fp64_tflops"] = R.measure_fp64_peak()
        _PEAKS["l2_gather_gbs"] = R.measure_l2_bandwidth(48 << 20, 64, 1)
        _PEAKS["l2_stream_gbs"] = R.measure_l2_bandwidth(48 << 20, 64, 0)
    return _PEAKS


def traffic_of(key, field="bytes"):
    """dram bytes of one launch of the dominant kernel (or another field) from the committed ncu capture of the same command"""
    for f in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", f)))
            if key in tj and field in tj[key]:
                return tj[key][field], f
        except Exception:
            pass
    return None, None


def roofline_of(R, wl, traversal, real, counters_sum, kernel_ns, launches, key):
    """roofline record of the walk kernel for one config: FP pipe for the shared-memory scenes, L2 gathers for the hierarchy"""
    pk = peaks(R)
    rays = counters_sum["rays_ext"] + counters_sum["rays_shadow"]
    traced = rays - counters_sum["rays_memo"]
    flops = flops_of(counters_sum)
    achieved_tf = flops / (kernel_ns * 1e-9) / 1e12
    peak_tf = pk["fp64_tflops"] if real == "f64" else pk["fp32_tflops"]
    traffic, tsrc = traffic_of(key)
    fp = {"bound": "fp64" if real == "f64" else "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
          "traffic": traffic, "traffic_unit": f"bytes per launch (dram read+write, ncu, profiles/{tsrc})" if tsrc else None,
          "kernel": "lt::walk_megakernel", "kernel_ms_per_launch": kernel_ns / 1e6 / launches,
          "peak_source": f"{'FP64' if real == 'f64' else 'FP32'} FMA micro-benchmark run in this process (lt_measure_fp{'64' if real == 'f64' else '32'}_peak); "
                         "MEASURED_PEAKS.json has no FP32/FP64 figure (nominal FP32: 148 SM x 128 x 2 x 1.965 GHz = 74.4)",
          "flops_model": "21*box_tests + 48*tri_tests + 150*(ext_rays - memoised) + 40*shadow_rays (SURVEY.md 8d)",
          "box_tests_per_ray": counters_sum["box_tests"] / max(traced, 1), "tri_tests_per_ray": counters_sum["tri_tests"] / max(traced, 1),
          "lane_utilisation": counters_sum["lane_steps"] / max(counters_sum["lane_slots"], 1)}
    if traversal not in ("lbvh", "lbvh2"):
        return fp
    # a large mesh is bound by node/triangle bytes.  Algorithmic bytes per ray: 64 B per visited binary node
    # (box_tests / 2) or 80 B per visited 8-wide node (box_tests / 8), 64 B (TriRec) or 48 B (WTri) per tested triangle, 48 B ray in /
    # hit out.  They are served by L1/L2 (the hierarchy is L2-resident), so the physical bound is the L2 -> SM gather bandwidth
    # measured by lt_measure_l2_bandwidth (one random 64-byte granule per lane); the HBM view uses the real dram bytes of ncu.
    wide = traversal == "lbvh"
    bytes_all = ((80.0 / 8 if wide else 64.0 / 2) * counters_sum["box_tests"] + 48.0 * counters_sum["tri_tests"] + 48.0 * rays)
    gbs_alg = bytes_all / (kernel_ns * 1e-9) / 1e9
    # what the L2 really delivered per ray: lts__t_sectors_srcunit_tex_op_read x 32 B of the ncu capture of this kernel on this scene
    # (profiles/r02_traffic.json), scaled by the live ray rate; without a capture the algorithmic bytes stand in (an upper bound)
    l2_per_ray, l2src = traffic_of(key, "l2_read_bytes_per_ray")
    gbs = (l2_per_ray * rays / (kernel_ns * 1e-9) / 1e9) if l2_per_ray else gbs_alg
    out = {"bound": "l2", "achieved": gbs, "peak": pk["l2_gather_gbs"], "unit": "GB/s", "frac": gbs / pk["l2_gather_gbs"],
           "achieved_source": (f"L2 -> SM sectors per ray measured by ncu (profiles/{l2src}: {l2_per_ray:.0f} B/ray, L1 misses only) x the live ray rate"
                               if l2_per_ray else "algorithmic bytes (no ncu capture for this variant): an upper bound, L1 hits included"),
           "traffic": traffic, "traffic_unit": fp["traffic_unit"], "kernel": "lt::walk_megakernel", "kernel_ms_per_launch": fp["kernel_ms_per_launch"],
           "algorithmic": {"bytes_per_ray": bytes_all / max(rays, 1), "gbs": gbs_alg,
                           "model": ("80 B x visited 8-wide nodes (= box_tests/8)" if wide else "64 B x visited binary nodes (= box_tests/2)")
                                    + " + 48 B x tri_tests + 48 B x rays (SURVEY.md 8d); served by L1 + L2"},
           "peak_source": "lt_measure_l2_bandwidth in this process: 128-bit ld.global.cg gathers, one random 64-byte granule per lane, 48 MiB buffer "
                          f"(streaming: {pk['l2_stream_gbs']:.0f} GB/s)",
           "hbm": {"peak": pk["hbm_gbs"], "peak_source": pk["hbm_source"], "unit": "GB/s",
                   "achieved": (traffic / (kernel_ns * 1e-9 / launches) / 1e9) if traffic else None,
                   "frac": (traffic / (kernel_ns * 1e-9 / launches) / 1e9 / pk["hbm_gbs"]) if traffic else None,
                   "note": "real DRAM bytes of one launch (ncu) over the live kernel time: the kernel is not HBM-bound"},
           "fp32": {k: fp[k] for k in ("achieved", "peak", "unit", "frac", "flops_model", "peak_source")},
           "box_tests_per_ray": fp["box_tests_per_ray"], "tri_tests_per_ray": fp["tri_tests_per_ray"], "lane_utilisation": fp["lane_utilisation"]}
    return out


def run_walk(cfg, wl, R, ps, nodes, cam, depth, *, width, height, spp, traversal, real, scheduler, steps, warmup, rank=0, world=1,
             stream=None, flush=None, sampler=None):
    """W warm-up + K timed passes of one config on this rank.  Returns the device-timed record."""
    import torch
    import torch.distributed as dist
    from light_transport_b200 import parallel
    spp_total = spp if wl["scaling"] == "strong" else spp * world
    s0, s1 = parallel.sample_range(rank, world, spp_total)
    film = torch.zeros(height * width * 3, dtype=torch.float64, device="cuda")
    image = torch.empty_like(film)
    kw = dict(width=width, height=height, spp=spp_total, s0=s0, s1=s1, max_depth=wl["max_depth"], camera=cam, f_distance=depth,
              semantics=wl["semantics"], traversal=traversal, real=real, scheduler=scheduler)

    def step(i):
        """one pass: clear film, walk, (reduce), finalize.  Returns (event ms, counters)."""
        flush.fill_(i & 0xff)                         # L2 flush, outside the timed events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        film.zero_()
        _, _, c = R.render(film=film, seed=1 + i, **kw)
        if world > 1:
            image.copy_(film)
            parallel.reduce_film(image, root=0, spp_total=spp_total)     # the single collective of a pass: lt_film_reduce (NCCL) + finalize on rank 0
        elif rank == 0:
            R.finalize(film, spp_total, out=image)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1), c

    for i in range(warmup):
        step(i)
    if sampler is not None:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, kernel_ns = 0.0, 0
    cs = {k: 0 for k in ("box_tests", "tri_tests", "rays_ext", "rays_shadow", "rays_memo", "paths", "lane_steps", "lane_slots")}
    for i in range(steps):
        t, c = step(100 + i)
        for k in cs:
            cs[k] += c[k]
        ms += t
        kernel_ns += c["kernel_ns"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tr = torch.tensor([float(cs["rays_ext"] + cs["rays_shadow"]), float(cs["paths"]), float(cs["rays_memo"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(tr, op=dist.ReduceOp.SUM)
    ms_max, rays_all, paths_all, memo_all = float(tm[0]), float(tr[0]), float(tr[1]), float(tr[2])
    del film, image
    return dict(value=rays_all / (ms_max * 1e-3) / 1e6, ms_per_step=ms_max / steps, mpaths_per_s=paths_all / (ms_max * 1e-3) / 1e6,
                traversed_mrays_per_s=(rays_all - memo_all) / (ms_max * 1e-3) / 1e6, rays_per_path=rays_all / max(paths_all, 1),
                rays_memo_fraction=memo_all / max(rays_all, 1), counters=cs, kernel_ns=kernel_ns, spp_total=spp_total)


def run_e2e(wl, ps, nodes, lights, cam, depth, *, width, height, spp_total, traversal, real, scheduler, steps, world, local):
    """the reference-facing call with HOST buffers (scene objects in, image array out), `steps` calls after one untimed call"""
    import torch
    import torch.distributed as dist
    import light_transport_b200 as lt
    from light_transport_b200 import api
    sc = api.Scene(camera=cam, lights=lights, width=width, height=height, max_depth=wl["max_depth"], f_distance=depth,
                   number_of_samples=1, allocate=False)
    h2d = sum(v.nbytes for v in ps.to_npz_dict().values())
    if nodes is not None:
        h2d += sum(getattr(nodes, k).nbytes for k in ("bmin", "bmax", "prim_off", "second", "n_prims", "axis"))
    kw = dict(rng="philox", semantics=wl["semantics"], traversal=traversal, real=real, world=(True if world > 1 else None), device=local,
              scheduler=scheduler)
    lt.render_scene(sc, ps, nodes, spp=max(1, spp_total // 64), seed=499, **kw)      # untimed: context, allocator pools, pinned staging
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from light_transport_b200 import host as _host

    def timed(clear):
        """`steps` calls; clear=True drops the cached uploaded scene before every call, inside the timed region"""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e_rays = 0
        for i in range(steps):
            if clear:
                _host.clear_renderer_cache()
            img, info = lt.render_scene(sc, ps, nodes, spp=spp_total, seed=500 + i, return_info=True, **kw)
            e_rays += info["rays_ext"] + info["rays_shadow"]
        torch.cuda.synchronize()
        et = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        er = torch.tensor([float(e_rays)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            dist.all_reduce(er, op=dist.ReduceOp.SUM)
        return float(er[0]) / float(et[0]) / 1e6

    strict = timed(True)          # every call uploads the scene (and builds its hierarchy): the contract's "inputs copied every step"
    cached = timed(False)         # what a caller's loop over the same scene gets: render_scene keeps the uploaded scene
    return {"value": strict, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(height * width * 3 * 8),
            "api": "light_transport_b200.render_scene(scene, primitives, bvh, spp=, rng='philox')", "steps": steps,
            "value_scene_cached": cached,
            "note": "value: the uploaded-scene cache is cleared before every call inside the timed region (scene H2D + hierarchy build every "
                    "step); value_scene_cached: repeated calls on the same scene, H2D per step is then the launch parameters only"}


def run_medium(args, rank, world, local, photons=None, steps=None, warmup=None, emit_line=True):
    """C4b: homogeneous scattering medium, albedo 0.99, max depth 256; a ray = one flight segment."""
    import torch
    import torch.distributed as dist
    from light_transport_b200 import medium
    wl = CONFIGS["c4b"]
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    torch.cuda.set_device(local)
    if world > 1 and emit_line:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = wl["photons"] if photons is None else photons
    kw = dict(albedo=wl["albedo"], g=wl["g"], radius=wl["radius"], max_depth=wl["max_depth"], mode=wl["semantics"],
              n_rbins=32, real=args.real, device=local)
    for i in range(warmup):
        medium.walk(n // 10, seed=i, first_photon=rank * n, **kw)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    seg, ms, wall0 = 0, 0.0, time.perf_counter()
    hist = np.zeros(wl["max_depth"] + 1, np.uint64)
    for i in range(steps):
        r = medium.walk(n, seed=100 + i, first_photon=rank * n, **kw)       # disjoint photon ranges per rank
        seg += r["n_segments"]; ms += r["kernel_ms"]; hist += r["hist"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    tm = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
    ts = torch.tensor([float(seg)], dtype=torch.float64, device="cuda")
    th = torch.from_numpy(hist.astype(np.float64)).cuda()
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(ts)
        dist.all_reduce(th)                                                  # the tally reduce of a pass
    out = None
    if rank == 0:
        out = {"metric": METRIC, "value": float(ts[0]) / (float(tm[0]) * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": steps,
               "warmup": warmup, "ms_per_step": float(tm[0]) / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": args.real, "data": "synthetic",
               "config": {"workload": f"C4b: homogeneous sphere R={wl['radius']} mfp, albedo {wl['albedo']}, HG g={wl['g']}, max depth "
                                      f"{wl['max_depth']}, implicit capture + roulette, {n} photons per GPU per step; a ray = one flight segment",
                          "photons_per_gpu": n, "parallelism": f"photon-sharded x{world}"},
               "e2e": {"value": float(ts[0]) / float(tm[1]) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 80, "d2h_bytes_per_step": int(hist.nbytes + 32 * 8 + 72),
                       "api": "light_transport_b200.medium.walk(...)"},
               "gpu_launches": steps, "mean_collisions": float((th.cpu().numpy() * np.arange(len(hist))).sum() / max(float(th.sum()), 1)),
               "parity": "unpinned by the reference (it has no volume walk, only R/src/medium_samples.py:14-16); pinned by the oracle + analytic expectations",
               "roofline": None}
        try:
            # the medium walk touches no memory to speak of (tallies in shared memory): its bound is the FP32 pipe.  Algorithmic
            # work of one flight segment + collision: 75 flops (free path 2, move 6, |x|^2 5, shell 3,
            # weight/roulette 7, Henyey-Greenstein cosine 10, sin/cos 5, rotation 27, renormalisation 10; a transcendental = 1)
            if not _PEAKS:
                from light_transport_b200 import host as _h
                _o, _l, _ps, _nodes, _c, _d = scene_arrays("lts")
                _R = _h.Renderer(_ps, _nodes, device=local)
                peaks(_R)
                _R.close()
            pk = _PEAKS["fp64_tflops" if args.real == "f64" else "fp32_tflops"]
            ach = 75.0 * float(ts[0]) / (float(tm[0]) * 1e-3) / 1e12
            out["roofline"] = {"bound": "fp64" if args.real == "f64" else "fp32", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                               "traffic": None, "kernel": "lt::medium_kernel", "kernel_ms_per_launch": float(tm[0]) / steps,
                               "flops_model": "75 flops per flight segment (DESIGN.md section 4); Philox4x32-10 (about 100 integer operations per "
                                              "segment) and the tallies are not counted",
                               "peak_source": "FMA micro-benchmark run in this process (lt_measure_fp32_peak / lt_measure_fp64_peak)"}
        except Exception as e:            # noqa: BLE001
            out["roofline"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            t0 = time.perf_counter()
            o = O.medium_walk(n // 200, seed=1, **{k: v for k, v in kw.items() if k not in ("real", "device")})
            dt = time.perf_counter() - t0
            v = o["n_segments"] / dt / 1e6
            out["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": os.cpu_count(), "value_per_core": v / (os.cpu_count() or 1), "kind": "port",
                                   "sample": f"{n // 200} photons, {o['n_segments']} segments, {dt:.2f} s (oracle/lt_oracle.c lto_medium_walk)"}
        if emit_line:
            emit(out)
    if world > 1 and emit_line:
        dist.destroy_process_group()
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    objs, lights, ps, nodes, cam, depth = scene_arrays(WORKLOAD["scene"])
    stride = 32 if WORKLOAD["scene"] == "mesh" else 1          # brute force over 1 M triangles: a 60x33 film
    sspp = 1 if WORKLOAD["scene"] == "mesh" else 2
    for _ in range(max(args.warmup, 0)):
        cpu_baseline(ps, nodes, cam, depth, WORKLOAD, args.width, args.height, args.spp, args.traversal, sample_spp=1, pixel_stride=stride * 2)
    tot_t, tot_r, cb = 0.0, 0, None
    for _ in range(max(args.steps, 1)):
        cb, dt, rays = cpu_baseline(ps, nodes, cam, depth, WORKLOAD, args.width, args.height, args.spp, args.traversal, sample_spp=sspp, pixel_stride=stride)
        tot_t += dt
        tot_r += rays
    v = tot_r / tot_t / 1e6
    cb["value"] = v
    cb["value_per_core"] = v / cb["cores"]
    emit({"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / max(args.steps, 1) * 1e3,
          "higher_is_better": True, "scaling": WORKLOAD["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": config_dict(args, 1), "cpu_baseline": cb, "cpu_baseline_reference": reference_timing(),
          "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


def config_dict(args, n):
    strong = WORKLOAD["scaling"] == "strong"
    per_gpu = args.spp // n if strong else args.spp
    return {"workload": f"{args.config.upper()}: {SCENE_DESC[WORKLOAD['scene']]}, {args.width}x{args.height}, "
                        f"{per_gpu} spp per GPU, max depth {WORKLOAD['max_depth']}, {WORKLOAD['semantics']} semantics, "
                        f"{args.traversal} traversal, {args.scheduler} scheduler, Philox4x32-10",
            "film": [args.width, args.height], "spp_per_gpu": per_gpu, "spp_total": args.spp if strong else args.spp * n,
            "max_depth": WORKLOAD["max_depth"], "semantics": WORKLOAD["semantics"], "traversal": args.traversal,
            "scheduler": args.scheduler, "rng": "philox", "parallelism": f"photon-sharded x{n}",
            "l2": "flushed between timed steps (256 MiB write)"}


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line of the contract, on the process's real stdout"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(obj) + "\n").encode())


def sub_record(name, host, local, stream, flush, *, steps, warmup, real="f32", traversal=None, e2e_steps=0, cpu=True, cpu_kw=None):
    """one of the other BASELINE configs (or the float64 arm) at its full size, measured in this process"""
    wl = CONFIGS[name]
    traversal = traversal or wl["traversal"]
    objs, lights, ps, nodes, cam, depth = scene_arrays(wl["scene"])
    R = host.Renderer(ps, nodes, device=local, build_lbvh=("binary" if traversal in ("lbvh", "lbvh2") else False))
    R.set_stream(stream.cuda_stream)
    r = run_walk(name, wl, R, ps, nodes, cam, depth, width=wl["width"], height=wl["height"], spp=wl["spp"], traversal=traversal, real=real,
                 scheduler="megakernel", steps=steps, warmup=warmup, stream=stream, flush=flush)
    key = f"{name}:{traversal}:megakernel:{real}"
    out = {"workload": f"{name.upper()}: {SCENE_DESC[wl['scene']]}, {wl['width']}x{wl['height']}, {wl['spp']} spp, max depth {wl['max_depth']}, "
                       f"{wl['semantics']} semantics, {traversal} traversal, {real}",
           "value": r["value"], "unit": "Mrays/s", "ms_per_step": r["ms_per_step"], "steps": steps, "warmup": warmup, "dtype": real,
           "mpaths_per_s": r["mpaths_per_s"], "traversed_mrays_per_s": r["traversed_mrays_per_s"], "rays_per_path": r["rays_per_path"],
           "rays_memo_fraction": r["rays_memo_fraction"],
           "roofline": roofline_of(R, wl, traversal, real, r["counters"], r["kernel_ns"], steps, key)}
    if R.lbvh_stats:
        out["hierarchy"] = {k: R.lbvh_stats[k] for k in ("n_prims", "n_nodes", "max_depth", "binary_depth", "build_ms", "sah_cost", "bytes")}
    R.close()
    if e2e_steps:
        out["e2e"] = run_e2e(wl, ps, nodes, lights, cam, depth, width=wl["width"], height=wl["height"], spp_total=wl["spp"], traversal=traversal,
                             real=real, scheduler="megakernel", steps=e2e_steps, world=1, local=local)
    if cpu:
        out["cpu_baseline"], _, _ = cpu_baseline(ps, nodes, cam, depth, wl, wl["width"], wl["height"], wl["spp"], traversal, **(cpu_kw or {}))
    return out


def strong_record(host, local, stream, flush, rank, world):
    """C5 (north_star's scaling config: fixed 4096^2 x 4096 spp job split over the ranks by sample index) as a sub-record of the
    default line at EVERY N, so that the driver's scaling run records strong scaling too.  Collective: all ranks call it."""
    wl = CONFIGS["c5"]
    objs, lights, ps, nodes, cam, depth = scene_arrays(wl["scene"])
    R = host.Renderer(ps, nodes, device=local)
    R.set_stream(stream.cuda_stream)
    kw = dict(width=wl["width"], height=wl["height"], traversal=wl["traversal"], real="f32", scheduler="megakernel", rank=rank, world=world,
              stream=stream, flush=flush)
    run_walk("c5", wl, R, ps, nodes, cam, depth, spp=max(64, world), steps=1, warmup=2, **kw)          # warm-up at 1/64 of the job
    r = run_walk("c5", wl, R, ps, nodes, cam, depth, spp=wl["spp"], steps=1, warmup=0, **kw)
    R.close()
    return {"workload": f"C5: {SCENE_DESC[wl['scene']]}, {wl['width']}x{wl['height']}, {wl['spp']} spp IN TOTAL over {world} GPU(s), max depth "
                        f"{wl['max_depth']}, {wl['semantics']} semantics, {wl['traversal']} traversal, f32; one lt_film_reduce per pass",
            "value": r["value"], "unit": "Mrays/s", "n_gpus": world, "scaling": "strong", "ms_per_step": r["ms_per_step"], "steps": 1,
            "warmup": "3 passes at 1/64 of the samples", "dtype": "f32", "mpaths_per_s": r["mpaths_per_s"],
            "traversed_mrays_per_s": r["traversed_mrays_per_s"], "rays_per_path": r["rays_per_path"]}


def main():
    global _REAL_STDOUT
    args = parse()
    # libraries (NCCL's version banner, torchrun notices) must not share stdout with the JSON line:
    # keep the real stdout aside and point fd 1 at stderr for everything else
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    from light_transport_b200 import host
    if args.lib:
        host.LIB_PATH = args.lib
    if WORKLOAD["scene"] == "medium":
        run_medium(args, rank, world, local)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    objs, lights, ps, nodes, cam, depth = scene_arrays(WORKLOAD["scene"])
    R = host.Renderer(ps, nodes, device=local, build_lbvh=("binary" if args.traversal in ("lbvh", "lbvh2") else False))
    stream = torch.cuda.current_stream()
    R.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(local) if rank == 0 else None
    r = run_walk(args.config, WORKLOAD, R, ps, nodes, cam, depth, width=args.width, height=args.height, spp=args.spp, traversal=args.traversal,
                 real=args.real, scheduler=args.scheduler, steps=args.steps, warmup=args.warmup, rank=rank, world=world, stream=stream,
                 flush=flush, sampler=sampler)
    clocks = sampler.stop() if rank == 0 else None
    e2e = run_e2e(WORKLOAD, ps, nodes, lights, cam, depth, width=args.width, height=args.height, spp_total=r["spp_total"], traversal=args.traversal,
                  real=args.real, scheduler=args.scheduler, steps=max(5, args.steps) if r["ms_per_step"] < 2000 else 2, world=world, local=local)

    full = args.size is None and args.spp == CONFIGS[args.config]["spp"]
    c5_strong = None
    if args.config == "c2" and not args.no_extra and full:
        try:                                  # every rank takes part (film reduce); a failure must not cost the headline line
            c5_strong = strong_record(host, local, stream, flush, rank, world)
        except Exception as e:                # noqa: BLE001
            c5_strong = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        key = f"{args.config}:{args.traversal}:{args.scheduler}:{args.real}" if full else "-"
        out = {"metric": METRIC, "value": r["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": WORKLOAD["scaling"], "vs_baseline": None,
               "dtype": args.real, "data": "synthetic", "config": config_dict(args, world),
               "mpaths_per_s": r["mpaths_per_s"], "traversed_mrays_per_s": r["traversed_mrays_per_s"], "rays_per_path": r["rays_per_path"],
               "ray_accounting": RAY_ACCOUNTING, "rays_memo_fraction": r["rays_memo_fraction"],
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(args.steps * (2 if world == 1 else 2)),
               "roofline": roofline_of(R, WORKLOAD, args.traversal, args.real, r["counters"], r["kernel_ns"], args.steps, key),
               "cpu_baseline_reference": reference_timing()}
        if R.lbvh_stats:
            out["hierarchy"] = {k: R.lbvh_stats[k] for k in ("n_prims", "n_nodes", "max_depth", "binary_depth", "build_ms", "sah_cost", "bytes")}
        if world == 1 and not args.no_cpu_baseline:
            mesh = WORKLOAD["scene"] == "mesh"
            out["cpu_baseline"], _, _ = cpu_baseline(ps, nodes, cam, depth, WORKLOAD, args.width, args.height, args.spp, args.traversal,
                                                     sample_spp=1 if mesh else 8, pixel_stride=32 if mesh else 1)
            cb = out["cpu_baseline"]
            out["speedup"] = {"e2e_over_port": e2e["value"] / cb["value"], "e2e_over_one_port_core": e2e["value"] / cb["value_per_core"],
                              "note": "the driver computes the official ratio from the two arms; the per-core figure does not depend on how many host "
                                      "cores the box happens to have"}
        if world == 1 and not args.no_rmse and args.config == "c2":
            out["time_to_rmse"] = time_to_rmse(R, cam, depth, args)
            if "cpu_baseline" in out:
                # the CPU counterpart, DERIVED: the same number of samples (the estimator is the same, so is its variance per sample)
                # at the port's measured ray rate — not a separately timed run
                t = out["time_to_rmse"]
                cpu_rays = t["film"][0] * t["film"][1] * t["spp"] * r["rays_per_path"]
                t["cpu_port_seconds_derived"] = cpu_rays / (out["cpu_baseline"]["value"] * 1e6)
                t["cpu_port_note"] = "derived: film x spp x rays/path of this run / cpu_baseline.value (same estimator, same spp to reach the target)"
    R.close()
    if rank == 0:
        if c5_strong is not None:
            out["configs"] = {"c5_strong": c5_strong}
        if world == 1 and args.config == "c2" and not args.no_extra and full:
            # ---- sub-records: the float64 arm of the headline config and the other BASELINE configs, same process, same clocks
            def guarded(fn):          # a failing sub-record must not cost the headline line
                try:
                    return fn()
                except Exception as e:        # noqa: BLE001
                    return {"error": f"{type(e).__name__}: {e}"}

            out["f64"] = guarded(lambda: sub_record("c2", host, local, stream, flush, steps=2, warmup=1, real="f64", e2e_steps=2, cpu=False))
            if "cpu_baseline" in out and "e2e" in out["f64"]:
                out["f64"]["speedup_over_port"] = out["f64"]["e2e"]["value"] / out["cpu_baseline"]["value"]
            extra = {}
            extra["c1"] = guarded(lambda: sub_record("c1", host, local, stream, flush, steps=20, warmup=5, e2e_steps=20, cpu_kw=dict(sample_spp=16)))
            extra["c3"] = guarded(lambda: sub_record("c3", host, local, stream, flush, steps=2, warmup=1, e2e_steps=2,
                                                     cpu_kw=dict(sample_spp=1, pixel_stride=32)))
            extra["c3_lbvh8"] = guarded(lambda: sub_record("c3", host, local, stream, flush, steps=2, warmup=1, traversal="lbvh", cpu=False))
            extra["c4a"] = guarded(lambda: sub_record("c4", host, local, stream, flush, steps=10, warmup=3, e2e_steps=5, cpu_kw=dict(sample_spp=8)))

            def medium_record():
                m = run_medium(argparse.Namespace(**vars(args)), 0, 1, local, steps=3, warmup=2, emit_line=False)
                rec = {k: m[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "dtype", "e2e", "mean_collisions", "parity", "roofline") if k in m}
                rec["workload"] = m["config"]["workload"]
                if "cpu_baseline" in m:
                    rec["cpu_baseline"] = m["cpu_baseline"]
                return rec

            extra["c4b"] = guarded(medium_record)
            extra["c5_strong"] = c5_strong
            out["configs"] = extra
            out["peaks"] = {k: v for k, v in peaks(None).items() if k in ("fp32_tflops", "fp64_tflops", "l2_gather_gbs", "l2_stream_gbs", "hbm_gbs", "hbm_source")}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
