#!/usr/bin/env python
"""bench.py — the reference's headline metric (Mrays/s, path segments per second) on config C2 of
BASELINE.json: the LTS example scene (Cornell box + glass cone, analytic triangles) at 1024x1024,
1024 spp, max depth 16, REFERENCE integrator semantics, COMPAT traversal (the reference's own
intersect_bvh results), float32 arithmetic, Philox4x32-10 random numbers.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one full pass of the walk over the film (1024 spp of every pixel).  A ray = one path
segment = one closest-hit or shadow query (R/src/utils.py:59 + R/src/light_samples.py:48).
N > 1: weak scaling — every rank renders 1024 spp of every pixel of a job that has 1024*N spp
(disjoint Philox counters), then ONE NCCL all-reduce of the float64 film inside the timed region.
--impl reference: the CPU oracle (C port of the reference's path, all host threads) on a bounded
sample of the same workload; the reference itself is numba-Python and cannot travel to the GPU box.
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
# with --config for the per-config table in DESIGN.md.  c5 is the strong-scaling sweep (fixed total work).
CONFIGS = {
    "c1": dict(scene="lts", width=256, height=256, spp=16, max_depth=8, semantics="reference", traversal="compat", scaling="weak"),
    "c2": dict(scene="lts", width=1024, height=1024, spp=1024, max_depth=16, semantics="reference", traversal="compat", scaling="weak"),
    "c3": dict(scene="mesh", width=1920, height=1080, spp=256, max_depth=8, semantics="iterative", traversal="lbvh", scaling="weak"),
    "c4": dict(scene="mirror", width=512, height=512, spp=64, max_depth=256, semantics="reference", traversal="compat", scaling="weak"),
    "c4b": dict(scene="medium", width=0, height=0, spp=0, max_depth=256, semantics="implicit", traversal="none", scaling="weak",
                albedo=0.99, g=0.8, radius=20.0, photons=50_000_000),
    # NEAREST semantics (true closest hit).  On a 28-triangle scene the all-primitive scan (15.0 Grays/s) beats the LBVH (11.3)
    "c5": dict(scene="lts", width=4096, height=4096, spp=4096, max_depth=8, semantics="reference", traversal="nearest", scaling="strong"),
}
WORKLOAD = dict(CONFIGS["c2"])
METRIC = "Mrays/s (path segments/s)"
SCENE_DESC = {"lts": "LTS example scene (Cornell box + glass cone, 28 triangles, 2000 light samples)",
              "mesh": "Cornell box + 1,000,000-triangle displaced-grid mesh (1,000,028 triangles, GPU LBVH)",
              "mirror": "closed mirror box (26 triangles, one Light; C4a)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--traversal", default=None, choices=["compat", "nearest", "lbvh"])
    ap.add_argument("--scheduler", default="megakernel", choices=["megakernel", "wavefront"])
    ap.add_argument("--real", default="f32", choices=["f32", "f64"])
    ap.add_argument("--spp", type=int, default=None, help="override samples per pixel (dev only; invalidates the number)")
    ap.add_argument("--size", type=int, default=None, help="override film size (dev only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rmse", action="store_true")
    a = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(CONFIGS[a.config])
    if a.traversal is None:
        a.traversal = WORKLOAD["traversal"]
    if a.spp is None:
        a.spp = WORKLOAD["spp"]
    a.width, a.height = (WORKLOAD["width"], WORKLOAD["height"]) if a.size is None else (a.size, a.size)
    return a


def scene_arrays():
    from light_transport_b200 import api, compat_bvh, from_arrays, pack_scene, scenes
    if WORKLOAD["scene"] == "mirror":
        objs, lights, cam, depth = scenes.mirror_box(api)
    else:
        objs, lights, cam, depth = scenes.lts_scene(api)
    ps = pack_scene(objs, lights)
    nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
    if WORKLOAD["scene"] == "mesh":
        ps = from_arrays(scenes.heightfield(nx=1000, ny=500), base=ps)
        nodes = None
    return objs, lights, ps, nodes, cam, depth


def flops_per_launch(c):
    """algorithmic FP32 work, SURVEY.md 8(d): F_ray = 21*N_box + 48*N_tri + F_shade,
    F_shade = 150 per extension ray, 40 per shadow ray; counts from the kernel's own counters."""
    return 21.0 * c["box_tests"] + 48.0 * c["tri_tests"] + 150.0 * c["rays_ext"] + 40.0 * c["rays_shadow"]


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


def cpu_baseline(ps, nodes, cam, depth, args, sample_spp=2, repeats=1):
    """the oracle (kind "port") on a bounded sample: every pixel of the film x `sample_spp` of the spp"""
    from oracle import oracle as O
    S = O.OracleScene(ps, nodes)
    threads = os.cpu_count() or 1
    kw = dict(width=args.width, height=args.height, spp=args.spp, s0=0, s1=min(sample_spp, args.spp), max_depth=WORKLOAD["max_depth"],
              camera=cam, f_distance=depth, semantics=WORKLOAD["semantics"],
              traversal="compat" if args.traversal == "compat" else "nearest", seed=1, n_threads=threads)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, _, c = O.render(S, **kw)
        dt = time.perf_counter() - t0
        rays = c["rays_ext"] + c["rays_shadow"]
        if best is None or dt < best[0]:
            best = (dt, rays, c)
    dt, rays, c = best
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"{args.width}x{args.height} pixels x {kw['s1']} of {args.spp} spp ({c['paths']} paths, {rays} rays, {dt:.2f} s, "
                      f"OpenMP {threads} threads, oracle/lt_oracle.c)"}, dt, rays


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
            "reference_spp": ref_spp, "reached": rmse <= target}


def run_medium(args, rank, world, local):
    """C4b: homogeneous scattering medium, albedo 0.99, max depth 256; a ray = one flight segment."""
    import torch
    import torch.distributed as dist
    from light_transport_b200 import medium
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = WORKLOAD["photons"]
    kw = dict(albedo=WORKLOAD["albedo"], g=WORKLOAD["g"], radius=WORKLOAD["radius"], max_depth=WORKLOAD["max_depth"], mode=WORKLOAD["semantics"],
              n_rbins=32, real=args.real, device=local)
    for i in range(args.warmup):
        medium.walk(n // 10, seed=i, first_photon=rank * n, **kw)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    seg, ms, wall0 = 0, 0.0, time.perf_counter()
    hist = np.zeros(WORKLOAD["max_depth"] + 1, np.uint64)
    for i in range(args.steps):
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
    if rank == 0:
        out = {"metric": METRIC, "value": float(ts[0]) / (float(tm[0]) * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": float(tm[0]) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": args.real, "data": "synthetic",
               "config": {"workload": f"C4b: homogeneous sphere R={WORKLOAD['radius']} mfp, albedo {WORKLOAD['albedo']}, HG g={WORKLOAD['g']}, max depth "
                                      f"{WORKLOAD['max_depth']}, implicit capture + roulette, {n} photons per GPU per step; a ray = one flight segment",
                          "photons_per_gpu": n, "parallelism": f"photon-sharded x{world}"},
               "e2e": {"value": float(ts[0]) / float(tm[1]) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 80, "d2h_bytes_per_step": int(hist.nbytes + 32 * 8 + 72),
                       "api": "light_transport_b200.medium.walk(...)"},
               "gpu_launches": args.steps, "mean_collisions": float((th.cpu().numpy() * np.arange(len(hist))).sum() / max(float(th.sum()), 1))}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            t0 = time.perf_counter()
            o = O.medium_walk(n // 100, seed=1, **{k: v for k, v in kw.items() if k not in ("real", "device")})
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": o["n_segments"] / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"{n // 100} photons, {o['n_segments']} segments, {dt:.2f} s (oracle/lt_oracle.c lto_medium_walk)"}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args, rank):
    if rank != 0:
        return
    objs, lights, ps, nodes, cam, depth = scene_arrays()
    for _ in range(max(args.warmup, 0)):
        cpu_baseline(ps, nodes, cam, depth, args, sample_spp=1)
    tot_t, tot_r, cb = 0.0, 0, None
    for _ in range(max(args.steps, 1)):
        cb, dt, rays = cpu_baseline(ps, nodes, cam, depth, args, sample_spp=2)
        tot_t += dt
        tot_r += rays
    v = tot_r / tot_t / 1e6
    cb["value"] = v
    emit({"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / max(args.steps, 1) * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": config_dict(args, 1), "cpu_baseline": cb,
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
    if WORKLOAD["scene"] == "medium":
        run_medium(args, rank, world, local)
        return

    import torch
    import torch.distributed as dist
    from light_transport_b200 import host, parallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    objs, lights, ps, nodes, cam, depth = scene_arrays()
    R = host.Renderer(ps, nodes, device=local, build_lbvh=(args.traversal == "lbvh"))
    stream = torch.cuda.current_stream()
    R.set_stream(stream.cuda_stream)
    W, H = args.width, args.height
    spp_total = args.spp if WORKLOAD["scaling"] == "strong" else args.spp * world
    s0, s1 = parallel.sample_range(rank, world, spp_total)
    film = torch.zeros(H * W * 3, dtype=torch.float64, device="cuda")
    image = torch.empty_like(film)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    kw = dict(width=W, height=H, spp=spp_total, s0=s0, s1=s1, max_depth=WORKLOAD["max_depth"], camera=cam, f_distance=depth,
              semantics=WORKLOAD["semantics"], traversal=args.traversal, real=args.real, scheduler=args.scheduler)

    def step(i, timed):
        """one pass: clear film, walk, (reduce), finalize.  Returns (event ms, counters)."""
        flush.fill_(i & 0xff)                         # L2 flush, outside the timed events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        film.zero_()
        _, _, c = R.render(film=film, seed=1 + i, **kw)
        if world > 1:
            parallel.reduce_film(film)                # the single collective of a pass
        R.finalize(film, spp_total, out=image)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1), c

    for i in range(args.warmup):
        step(i, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, rays, paths, kernel_ns, flops, last, memo = 0.0, 0, 0, 0, 0.0, None, 0
    counters_sum = {"box_tests": 0, "tri_tests": 0, "rays_ext": 0, "rays_shadow": 0}
    for i in range(args.steps):
        t, c = step(100 + i, True)
        for k in counters_sum:
            counters_sum[k] += c[k]
        ms += t
        rays += c["rays_ext"] + c["rays_shadow"]
        paths += c["paths"]
        memo += c["rays_memo"]
        kernel_ns += c["kernel_ns"]
        flops += flops_per_launch(c)
        last = c
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tr = torch.tensor([float(rays), float(paths)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(tr, op=dist.ReduceOp.SUM)
    ms_max, rays_all, paths_all = float(tm[0]), float(tr[0]), float(tr[1])
    value = rays_all / (ms_max * 1e-3) / 1e6

    # ---- e2e: the reference-facing call with HOST buffers (scene objects in, image array out)
    import light_transport_b200 as lt
    from light_transport_b200 import api
    e2e_steps = max(1, min(args.steps, 2))
    sc = api.Scene(camera=cam, lights=lights, width=W, height=H, max_depth=WORKLOAD["max_depth"], f_distance=depth,
                   number_of_samples=1, allocate=False)
    h2d = sum(v.nbytes for v in ps.to_npz_dict().values())
    if nodes is not None:
        h2d += sum(getattr(nodes, k).nbytes for k in ("bmin", "bmax", "prim_off", "second", "n_prims", "axis"))
    d2h = H * W * 3 * 8
    # one untimed call first (context cache, allocator pools, pinned staging), like the warm-up steps of the device-timed loop
    lt.render_scene(sc, ps, nodes, spp=max(1, spp_total // 64), rng="philox", seed=499, semantics=WORKLOAD["semantics"],
                    traversal=args.traversal, real=args.real, world=(True if world > 1 else None), device=local, scheduler=args.scheduler)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e_rays = 0
    for i in range(e2e_steps):
        img, info = lt.render_scene(sc, ps, nodes, spp=spp_total, rng="philox", seed=500 + i, semantics=WORKLOAD["semantics"],
                                    traversal=args.traversal, real=args.real, world=(True if world > 1 else None), device=local,
                                    return_info=True, scheduler=args.scheduler)
        e_rays += info["rays_ext"] + info["rays_shadow"]
    torch.cuda.synchronize()
    et = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    er = torch.tensor([float(e_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dist.all_reduce(er, op=dist.ReduceOp.SUM)
    e2e_value = float(er[0]) / float(et[0]) / 1e6

    if rank == 0:
        peak_tf = R.measure_fp32_peak()
        traffic = None
        try:      # dram bytes of one launch of this kernel from the committed ncu capture of the same command
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            key = f"{args.config}:{args.traversal}:{args.scheduler}:{args.real}"
            if key in tj and args.size is None and args.spp == CONFIGS[args.config]["spp"]:
                traffic = tj[key]["bytes"]
        except Exception:
            traffic = None
        achieved_tf = flops / (kernel_ns * 1e-9) / 1e12
        out = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": WORKLOAD["scaling"], "vs_baseline": None,
               "dtype": args.real, "data": "synthetic", "config": config_dict(args, world),
               "mpaths_per_s": paths_all / (ms_max * 1e-3) / 1e6, "rays_per_path": rays_all / max(paths_all, 1),
               "ray_accounting": "a ray = one intersect_bvh call of the reference's control flow (R/src/utils.py:59, "
                                 "R/src/light_samples.py:48), the same count the CPU arm reports; rays_memo_fraction of them are "
                                 "re-traces of a ray object that has just missed (Q1) and are answered without a traversal",
               "rays_memo_fraction": memo / max(rays, 1),
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "api": "light_transport_b200.render_scene(scene, primitives, bvh, spp=, rng='philox')", "steps": e2e_steps},
               "gpu_launches": int(args.steps * 2),
               "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                            "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu, profiles/r01_traffic.json)", "kernel": "lt::walk_megakernel",
                            "peak_source": "FP32 FMA micro-benchmark run in this process (lt_measure_fp32_peak); MEASURED_PEAKS.json has "
                                           "no FP32 figure (nominal 148 SM x 128 x 2 x 1.965 GHz = 74.4)",
                            "flops_model": "21*box_tests + 48*tri_tests + 150*ext_rays + 40*shadow_rays (SURVEY.md 8d)",
                            "kernel_ms_per_launch": kernel_ns / 1e6 / args.steps,
                            "box_tests_per_ray": last["box_tests"] / max(last["rays_ext"] + last["rays_shadow"], 1),
                            "tri_tests_per_ray": last["tri_tests"] / max(last["rays_ext"] + last["rays_shadow"], 1),
                            "lane_utilisation": last["lane_steps"] / max(last["lane_slots"], 1)}}
        if args.traversal == "lbvh" and WORKLOAD["scene"] == "mesh":
            # a large mesh is bound by node/triangle bytes, not flops.  B_ray = 64 B per visited
            # binary node (both child boxes + links; box_tests counts 2 per visit) + 48 B per tested triangle +
            # 48 B ray in / hit out.  Served mostly from L2 (the tree is L2-resident), so this is an upper bound
            # on what HBM had to deliver; the denominator is still the measured HBM copy bandwidth.
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            hbm = float(peaks.get("hbm_gbs", 6556.5))
            bytes_all = (32.0 * counters_sum["box_tests"] + 48.0 * counters_sum["tri_tests"]
                         + 48.0 * (counters_sum["rays_ext"] + counters_sum["rays_shadow"]))
            gbs = bytes_all / (kernel_ns * 1e-9) / 1e9
            rf = out["roofline"]
            rf["fp32"] = {"achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                          "flops_model": rf.pop("flops_model"), "peak_source": rf.pop("peak_source")}
            rf.update({"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                       "bytes_model": "64 B x visited nodes (= box_tests/2) + 48 B x tri_tests + 48 B x rays (SURVEY.md 8d); "
                                      "algorithmic bytes, mostly served by the 126 MB L2",
                       "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6556.5 GB/s (B200_PROFILING.md)",
                       "bytes_per_ray": bytes_all / max(counters_sum["rays_ext"] + counters_sum["rays_shadow"], 1)})
        if world == 1 and not args.no_cpu_baseline and WORKLOAD["scene"] != "mesh":
            out["cpu_baseline"], _, _ = cpu_baseline(ps, nodes, cam, depth, args, sample_spp=8)
        if world == 1 and not args.no_rmse and args.config == "c2":
            out["time_to_rmse"] = time_to_rmse(R, cam, depth, args)
        emit(out)
    R.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
