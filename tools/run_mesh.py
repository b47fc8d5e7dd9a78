"""Dev: renders of the C3 mesh scene (1 M-triangle heightfield in the Cornell box) for A/B runs and profiling.
    run_mesh.py [--lib path] [--spp 8] [--chunks 0,4,8] [--ctas 0,4,5] [--binary] [--nx 1000 --ny 500] [--l2]"""
import argparse
import sys
sys.path.insert(0, ".")
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--chunks", default="0")
ap.add_argument("--ctas", default="0")
ap.add_argument("--binary", action="store_true", help="keep the binary nodes too (LT_WIDE=0 variant builds need them)")
ap.add_argument("--nx", type=int, default=1000)
ap.add_argument("--ny", type=int, default=500)
ap.add_argument("--l2", action="store_true", help="also run the L2 / FP peaks micro-benchmarks")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--leaf", type=int, default=0)
ap.add_argument("--bleaf", type=int, default=0, help="triangles per leaf of the binary tree (0 = library default)")
ap.add_argument("--l2win", action="store_true")
ap.add_argument("--trav", default="lbvh2", choices=["lbvh", "lbvh2"])
a = ap.parse_args()
import torch
from light_transport_b200 import api, host, pack_scene, scenes, from_arrays
if a.lib:
    host.LIB_PATH = a.lib
objs, lights, cam, depth = scenes.lts_scene(api)
ps = from_arrays(scenes.heightfield(nx=a.nx, ny=a.ny), base=pack_scene(objs, lights))
R = host.Renderer(ps, None)
print("lbvh:", R.build_lbvh(keep_binary=True, leaf_max=a.leaf, binary_leaf=a.bleaf))
film = torch.zeros(1080 * 1920 * 3, dtype=torch.float64, device="cuda")
kw = dict(width=1920, height=1080, spp=a.spp, max_depth=8, camera=cam, f_distance=depth, seed=3, semantics="iterative", traversal=a.trav,
          real="f32", film=film)
for ct in [int(x) for x in a.ctas.split(",")]:
    for ch in [int(x) for x in a.chunks.split(",")]:
        for _ in range(a.reps):
            c = R.render(lbvh_chunk=ch, ctas_per_sm=ct, l2_window=a.l2win, **kw)[2]
        rays = c["rays_ext"] + c["rays_shadow"]
        print(f"ctas {ct} chunk {ch:3d}: {c['kernel_ns']/1e6:8.2f} ms  {rays / c['kernel_ns'] * 1e3:8.1f} Mrays/s  box/ray {c['box_tests']/rays:.1f}  "
              f"tri/ray {c['tri_tests']/rays:.2f}  rays/path {rays / c['paths']:.2f}  lane util {c['lane_steps'] / max(c['lane_slots'], 1):.2f}  "
              f"film sum {float(film.sum()):.6g}")
        film.zero_()
if a.l2:
    print("fp32 peak TFLOP/s", R.measure_fp32_peak(), " fp64 peak", R.measure_fp64_peak())
    for mb in (16, 48, 96, 256, 1024):
        print(f"L2 gather, {mb} MiB buffer:", {f"{g}B/{'rand' if m else 'seq'}": round(R.measure_l2_bandwidth(mb << 20, g, m)) for g in (16, 64, 80, 128) for m in (0, 1)})
