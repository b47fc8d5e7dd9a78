"""Dev probe (GPU box): throughput of the walk kernel variants on the LTS scene + fp32 peak."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from light_transport_b200 import api, host, pack_scene, scenes, compat_bvh, _pack

objs, lights, cam, depth = scenes.lts_scene(api)
ps = pack_scene(objs, lights)
nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
R = host.Renderer(ps, nodes, build_lbvh=True)
print("device", R.device_info(), "fp32 peak TFLOP/s", R.measure_fp32_peak())
W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
import torch
film = torch.zeros(H * W * 3, dtype=torch.float64, device="cuda")
for real in ("f32", "f64"):
    for sem in ("reference", "iterative"):
        for trav in ("compat", "nearest", "lbvh"):
            for ppw in (0,):
                kw = dict(width=W, height=H, spp=spp, max_depth=16, camera=cam, f_distance=depth, seed=1, semantics=sem,
                          traversal=trav, real=real, film=film, paths_per_warp=ppw)
                R.render(**kw)
                best = None
                for _ in range(3):
                    _, _, c = R.render(**kw)
                    if best is None or c["kernel_ns"] < best["kernel_ns"]:
                        best = c
                rays = best["rays_ext"] + best["rays_shadow"]
                ms = best["kernel_ns"] / 1e6
                print(f"{real} {sem:9s} {trav:7s} ppw={ppw}: {ms:8.2f} ms  {rays/ms/1e3:8.1f} Mrays/s  {best['paths']/ms/1e3:7.1f} Mpaths/s  "
                      f"rays/path {rays/best['paths']:.2f} box/ray {best['box_tests']/rays:.1f} tri/ray {best['tri_tests']/rays:.2f} "
                      f"lane-eff {best['lane_steps']/max(best['lane_slots'],1):.3f}")
# C3-like mesh
base = ps
mesh = _pack.from_arrays(scenes.heightfield(nx=1000, ny=500), base=base)
t0 = time.time()
Rm = host.Renderer(mesh, None, build_lbvh=True)
print("mesh scene", mesh.n_tris, "tris; create+build wall", time.time() - t0, "s; lbvh", Rm.lbvh_stats)
film = torch.zeros(1080 * 1920 * 3, dtype=torch.float64, device="cuda")
for sem in ("iterative", "reference"):
    kw = dict(width=1920, height=1080, spp=8, max_depth=8, camera=cam, f_distance=depth, seed=3, semantics=sem,
              traversal="lbvh", real="f32", film=film)
    Rm.render(**kw)
    _, _, c = Rm.render(**kw)
    rays = c["rays_ext"] + c["rays_shadow"]
    ms = c["kernel_ns"] / 1e6
    print(f"mesh f32 {sem} lbvh: {ms:.2f} ms {rays/ms/1e3:.1f} Mrays/s rays/path {rays/c['paths']:.2f} box/ray {c['box_tests']/rays:.1f} "
          f"tri/ray {c['tri_tests']/rays:.2f} lane-eff {c['lane_steps']/max(c['lane_slots'],1):.3f}")
