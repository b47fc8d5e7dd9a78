"""Dev: where does a small render_scene call spend its time?  e2e_profile.py [c1|c4]"""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import numpy as np
import light_transport_b200 as lt
from light_transport_b200 import api, scenes, pack_scene, compat_bvh, host
objs, lights, cam, depth = scenes.lts_scene(api)
ps = pack_scene(objs, lights)
nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
sc = api.Scene(camera=cam, lights=lights, width=256, height=256, max_depth=8, f_distance=depth, number_of_samples=1, allocate=False)
kw = dict(spp=16, rng="philox", semantics="reference", traversal="compat", real="f32")
for _ in range(3):
    lt.render_scene(sc, ps, nodes, seed=1, **kw)
t0 = time.perf_counter()
for i in range(50):
    lt.render_scene(sc, ps, nodes, seed=2 + i, **kw)
print(f"render_scene: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per call")
R = host.Renderer(ps, nodes)
img = np.empty((256, 256, 3))
rk = dict(width=256, height=256, spp=16, max_depth=8, camera=cam, f_distance=depth, semantics="reference", traversal="compat", real="f32")
for _ in range(3):
    R.render(image=img, seed=1, **rk)
t0 = time.perf_counter()
for i in range(50):
    _, _, c = R.render(image=img, seed=2 + i, **rk)
print(f"Renderer.render(image=): {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per call, kernel {c['kernel_ns'] / 1e6:.3f} ms")
t0 = time.perf_counter()
for i in range(50):
    r2 = host.Renderer(ps, nodes); r2.close()
print(f"Renderer create+close: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for i in range(50):
    lt.render_scene(sc, ps, nodes, seed=100 + i, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
