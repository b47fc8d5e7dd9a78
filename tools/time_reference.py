#!/usr/bin/env python
"""Time the UNMODIFIED reference's own CPU path in the build container.

    python tools/time_reference.py [--size 256] [--spp 16] [--depth 8] [--out profiles/r02_reference_timing.json]

Needs /root/reference + numba, so it cannot run on the GPU box; its JSON output is committed under
profiles/ and quoted by bench.py as `cpu_baseline_reference`.

  Timing A  path_tracing_old.render_scene(scene, prims, linear_bvh)        R/src/path_tracing_old.py:142-171
            the reference's own parallel entry point (numba.prange threads), config C1.
  Timing B  a harness @njit driver that restates R/src/path_tracing.py:263-288 (ray generation) around the
            reference's path_tracing.trace_path (R/src/path_tracing.py:23-159, calculate_grad=False), fanned
            out over os.cpu_count() PROCESSES by pixel rows (numba threads scale poorly on this code).

Rays (= intersect_bvh calls) are counted by the CPU oracle replaying the SAME rand_0/rand_1: its control
flow is bit-identical to trace_path's (tests/test_oracle_golden.py), and the number of shadow / extension
queries does not depend on which Light a shadow ray picks.
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time
import types

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

import importlib  # noqa: E402

import numba  # noqa: E402

P = "LightTransportSimulator.light_transport.src."
r_scene = importlib.import_module(P + "scene")
r_prim = importlib.import_module(P + "primitives")
r_mat = importlib.import_module(P + "material")
r_const = importlib.import_module(P + "constants")
r_bvh = importlib.import_module(P + "bvh_new")
r_ls = importlib.import_module(P + "light_samples")
r_rays = importlib.import_module(P + "rays")
r_vec = importlib.import_module(P + "vectors")
r_pt = importlib.import_module(P + "path_tracing")
r_old = importlib.import_module(P + "path_tracing_old")

from light_transport_b200 import scenes, _pack, compat_bvh  # noqa: E402

ref_ns = types.SimpleNamespace(
    Color=r_mat.Color, Material=r_mat.Material, PreComputedTriangle=r_prim.PreComputedTriangle,
    Light=r_scene.Light, generate_area_light_samples=r_ls.generate_area_light_samples,
    WHITE=r_const.WHITE, WHITE_2=r_const.WHITE_2, RED=r_const.RED, GREEN=r_const.GREEN,
    GLASS_MAT=r_const.GLASS_MAT)

Ray = r_rays.Ray
ref_trace = r_pt.trace_path
normalize = r_vec.normalize


def ref_bvh(objects):
    boxes = [r_bvh.BoundedBox(objects[i], i) for i in range(len(objects))]
    root, boxes, ordered, total = r_bvh.build_bvh(objects, boxes, 0, len(boxes), [], 0)
    lin = numba.typed.List()
    for _ in range(total):
        lin.append(r_bvh.LinearBVHNode())
    lin, _ = r_bvh.flatten_bvh(lin, root, 0)
    return lin


@numba.njit
def drive_rows(scene, prims, bvh, row0, row1, out):
    # restates R/src/path_tracing.py:263-288 (ray generation) around the reference's trace_path
    H, W, S = scene.height, scene.width, scene.number_of_samples
    top_bottom = np.linspace(scene.top, scene.bottom, H)
    left_right = np.linspace(scene.left, scene.right, W)
    for i in range(row0, row1):
        y = top_bottom[i]
        for j in range(W):
            x = left_right[j]
            for s in range(S):
                r0 = scene.rand_0[i, j, s]
                r1 = scene.rand_1[i, j, s]
                end = np.array([x, y, scene.f_distance, 1], dtype=np.float64)
                end[0] += r0[0] / W
                end[1] += r0[0] / H
                origin = np.array([scene.camera[0], scene.camera[1], scene.camera[2], 1], dtype=np.float64)
                direction = normalize(end - origin)
                ray = Ray(origin, direction)
                L, _lp, _ll = ref_trace(scene, prims, bvh, ray, 0, r0, r1, (i, j, s), False)
                out[i, j, s, :] = L


def make_scene(objects, lights, camera, depth, W, H, spp, D, seed):
    sc = r_scene.Scene(camera=camera, lights=lights, width=W, height=H, max_depth=D, f_distance=depth, number_of_samples=spp)
    np.random.seed(seed)          # Scene.__init__ draws from numba's private generator: replace by a reproducible stream
    sc.rand_0 = np.random.rand(H, W, spp, D)
    sc.rand_1 = np.random.rand(H, W, spp, D)
    return sc


_G = {}


def _worker(args):
    row0, row1 = args
    sc, objects, lin = _G["sc"], _G["objects"], _G["lin"]
    out = np.zeros((row1 - row0 + 0, sc.width, sc.number_of_samples, 3))
    full = np.zeros((sc.height, sc.width, sc.number_of_samples, 3))
    t0 = time.perf_counter()
    drive_rows(sc, objects, lin, row0, row1, full)
    dt = time.perf_counter() - t0
    out[...] = full[row0:row1]
    return row0, row1, dt, float(out.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r02_reference_timing.json"))
    a = ap.parse_args()
    W = H = a.size
    cores = os.cpu_count() or 1
    t_imp = time.perf_counter()
    objects, lights, camera, depth = scenes.lts_scene(ref_ns, new_list=numba.typed.List)
    lin = ref_bvh(objects)
    packed = _pack.pack_scene(objects, lights)
    nodes = compat_bvh.from_linear_bvh(lin)
    print(f"scene: {len(objects)} prims, {len(lin)} nodes, {len(lights)} lights ({time.perf_counter() - t_imp:.1f} s host setup)", flush=True)

    # ---- ray count of the job from the oracle on the same random numbers
    from oracle import oracle as O
    sc = make_scene(objects, lights, camera, depth, W, H, a.spp, a.depth, seed=1)
    S = O.OracleScene(packed, nodes)
    _, _, oc = O.render(S, width=W, height=H, spp=a.spp, max_depth=a.depth, camera=camera, f_distance=depth, semantics="reference",
                        traversal="compat", rand0=np.asarray(sc.rand_0), rand1=np.asarray(sc.rand_1), seed=0)
    paths = W * H * a.spp
    rays = oc["rays_ext"] + oc["rays_shadow"]
    print(f"job: {paths} paths, {rays} rays ({rays / paths:.3f} per path; oracle replay of the same rand_0/rand_1)", flush=True)

    # ---- Timing A: path_tracing_old.render_scene (parallel=True prange)
    t0 = time.perf_counter()
    warm = make_scene(objects, lights, camera, depth, 24, 16, 2, a.depth, seed=7)
    r_old.render_scene(warm, objects, lin)
    jit_a = time.perf_counter() - t0
    sc_a = make_scene(objects, lights, camera, depth, W, H, a.spp, a.depth, seed=1)     # render_scene scribbles inf into rand_0
    t0 = time.perf_counter()
    img_a = r_old.render_scene(sc_a, objects, lin)
    dt_a = time.perf_counter() - t0
    thr = numba.get_num_threads()
    print(f"Timing A: path_tracing_old.render_scene {W}x{H}x{a.spp}, D={a.depth}: {dt_a:.2f} s ({jit_a:.1f} s JIT+warm-up excluded), "
          f"{thr} numba threads, {paths / dt_a:.0f} paths/s, {rays / dt_a / 1e6:.4f} Mrays/s", flush=True)

    # ---- Timing B: trace_path harness over processes (compile once, then fork)
    t0 = time.perf_counter()
    tmp = np.zeros((16, 24, 2, 3))
    drive_rows(warm, objects, lin, 0, 2, tmp)
    jit_b = time.perf_counter() - t0
    _G.update(sc=sc, objects=objects, lin=lin)
    stripes = [(r * H // (cores * 4), (r + 1) * H // (cores * 4)) for r in range(cores * 4)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_worker, stripes, chunksize=1)
    dt_b = time.perf_counter() - t0
    cpu_b = sum(r[2] for r in res)
    print(f"Timing B: trace_path harness, {cores} processes: {dt_b:.2f} s wall ({cpu_b:.1f} core-seconds, {jit_b:.1f} s JIT excluded), "
          f"{paths / dt_b:.0f} paths/s, {rays / dt_b / 1e6:.4f} Mrays/s; per core {rays / cpu_b / 1e6:.4f} Mrays/s", flush=True)

    out = {
        "what": "the unmodified reference (numba) timed in the BUILD container; BASELINE.md section 3",
        "host": {"cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t"),
                 "cores": cores, "numba_threads": thr, "numba": numba.__version__, "numpy": np.__version__, "python": platform.python_version()},
        "config": {"workload": f"C1: LTS scene {W}x{H}, {a.spp} spp, max depth {a.depth}", "paths": paths, "rays": rays,
                   "rays_per_path": rays / paths, "ray_count_source": "oracle replay of the same rand_0/rand_1 (REFERENCE semantics, COMPAT)"},
        "timing_a": {"entry": "path_tracing_old.render_scene (R/src/path_tracing_old.py:142-171), numba.prange", "seconds": dt_a,
                     "seconds_jit_and_warmup": jit_a, "threads": thr, "paths_per_s": paths / dt_a, "mrays_per_s": rays / dt_a / 1e6,
                     "mrays_per_s_incl_jit": rays / (dt_a + jit_a) / 1e6, "image_mean": float(np.asarray(img_a).mean())},
        "timing_b": {"entry": "harness over path_tracing.trace_path (R/src/path_tracing.py:23-159), pixel rows over processes",
                     "seconds": dt_b, "seconds_jit": jit_b, "processes": cores, "core_seconds": cpu_b, "paths_per_s": paths / dt_b,
                     "mrays_per_s": rays / dt_b / 1e6, "mrays_per_s_per_core": rays / cpu_b / 1e6,
                     "mrays_per_s_incl_jit": rays / (dt_b + jit_b) / 1e6, "radiance_sum": sum(r[3] for r in res)},
    }
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
