"""Dev: share of float32 paths on the LTS scene whose radiance differs from the float64 walk of the same Philox streams."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from conftest import load_scene
from light_transport_b200 import host
if "--lib" in sys.argv:
    k = sys.argv.index("--lib")
    host.LIB_PATH = sys.argv[k + 1]
    del sys.argv[k:k + 2]
ps, nodes, cam, f = load_scene("lts_scene.npz")
R = host.Renderer(ps, nodes, build_lbvh="binary")
for sem in ("reference", "iterative"):
    for trav in ("compat", "nearest", "lbvh", "lbvh2"):
        for D in (2, 8, 16):
            kw = dict(width=96, height=96, spp=4, max_depth=D, camera=cam, f_distance=f, semantics=sem, seed=5, traversal=trav, return_samples=True)
            _, s64, c64 = R.render(real="f64", **kw)
            _, s32, c32 = R.render(real="f32", **kw)
            err = np.abs(s32 - s64) / np.maximum(np.abs(s64), 1e-3)
            print(f"{sem:9s} {trav:7s} D={D:2d}: paths differing by > 1e-3: {100 * (err > 1e-3).any(axis=-1).mean():6.2f} %   rays f64 {c64['rays_ext']}+{c64['rays_shadow']}  f32 {c32['rays_ext']}+{c32['rays_shadow']}")
