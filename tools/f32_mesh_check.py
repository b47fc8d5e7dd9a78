"""Dev: fraction of float32 paths on the C3 mesh whose radiance differs from the float64 walk of the same Philox streams.
    f32_mesh_check.py [--lib path] [--binary]"""
import argparse
import sys
sys.path.insert(0, ".")
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--binary", action="store_true")
ap.add_argument("--no-mesh", action="store_true")
ap.add_argument("--no-cone", action="store_true")
a = ap.parse_args()
import numpy as np
from light_transport_b200 import api, host, pack_scene, scenes, from_arrays
if a.lib:
    host.LIB_PATH = a.lib
objs, lights, cam, depth = scenes.lts_scene(api)
if a.no_cone:
    objs = [o for o in objs if o.material.transmission <= 0 or o.material.is_diffuse]
ps = pack_scene(objs, lights)
print("base triangles", ps.n_tris)
if not a.no_mesh:
    ps = from_arrays(scenes.heightfield(nx=1000, ny=500), base=ps)
R = host.Renderer(ps, None)
R.build_lbvh(keep_binary=a.binary)
for D in (1, 2, 3, 8):
    kw = dict(width=128, height=72, spp=2, max_depth=D, camera=cam, f_distance=depth, semantics="iterative", seed=3, traversal="lbvh", return_samples=True)
    _, s64, c64 = R.render(real="f64", **kw)
    _, s32, c32 = R.render(real="f32", **kw)
    err = np.abs(s32 - s64) / np.maximum(np.abs(s64), 1e-3)
    print(f"D={D}: paths differing by more than " + ", ".join(f"{tol:g}: {100 * (err > tol).any(axis=-1).mean():.2f} %" for tol in (1e-5, 1e-3, 1e-1)))
    print("   rays f64", c64["rays_ext"], c64["rays_shadow"], " f32", c32["rays_ext"], c32["rays_shadow"], " mean radiance", s64.mean(), s32.mean())
    if D == 1:
        bad = (err > 1e-3).any(axis=-1)
        idx = np.argwhere(bad)[:12]
        for i, j, k in idx:
            print("   pixel", i, j, k, "f64", s64[i, j, k], "f32", s32[i, j, k])
