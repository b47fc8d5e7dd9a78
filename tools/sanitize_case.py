"""Small coverage case for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from light_transport_b200 import api, host, pack_scene, scenes, compat_bvh, from_arrays, medium

objs, lights, cam, depth = scenes.lts_scene(api)
ps = pack_scene(objs, lights)
nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
R = host.Renderer(ps, nodes, build_lbvh=True)
kw = dict(width=24, height=16, spp=3, max_depth=20, camera=cam, f_distance=depth, seed=1)
for real in ("f32", "f64"):
    for sched in ("megakernel", "wavefront"):
        for trav in ("compat", "nearest", "lbvh"):
            for sem in ("reference", "iterative"):
                f, s, c = R.render(traversal=trav, real=real, scheduler=sched, semantics=sem, return_samples=True, **kw)
                assert c["paths"] == 24 * 16 * 3
rng = np.random.default_rng(0)
r0, r1 = rng.random((8, 8, 2, 6)), rng.random((8, 8, 2, 6))
o = R.render_records(width=8, height=8, spp=2, max_depth=6, camera=cam, f_distance=depth, rand0=r0, rand1=r1, gradients=True)
prim, t, _ = R.intersect(rng.uniform(-7, 7, (500, 3)), rng.normal(size=(500, 3)), "lbvh", "f32")
R.close()
mesh = from_arrays(scenes.heightfield(nx=40, ny=20), base=ps)
Rm = host.Renderer(mesh, None, build_lbvh=True)
Rm.render(traversal="lbvh", real="f32", **kw)
Rm.close()
m = medium.walk(20000, albedo=0.99, g=0.8, radius=5.0, max_depth=64, mode="implicit", real="f32")
m = medium.walk(20000, albedo=0.9, g=0.0, radius=0.0, max_depth=32, mode="analog", real="f64")
print("sanitize case done", int(m["hist"].sum()))
