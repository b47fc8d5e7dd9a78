"""Dev: where does the end-to-end call spend its time?"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import light_transport_b200 as lt
from light_transport_b200 import api, host, scenes, compat_bvh, pack_scene
objs, lights, cam, depth = scenes.lts_scene(api)
ps = pack_scene(objs, lights)
nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
sc = api.Scene(camera=cam, lights=lights, width=1024, height=1024, max_depth=16, f_distance=depth, number_of_samples=1, allocate=False)
for it in range(4):
    t0 = time.perf_counter()
    R = host.Renderer(ps, nodes)
    t1 = time.perf_counter()
    img = np.empty((1024, 1024, 3))
    _, _, c = R.render(width=1024, height=1024, spp=256, max_depth=16, camera=cam, f_distance=depth, seed=it, real="f32", image=img)
    t2 = time.perf_counter()
    R.close()
    t3 = time.perf_counter()
    img2, info = lt.render_scene(sc, ps, nodes, spp=256, rng="philox", seed=it, real="f32", return_info=True)
    t4 = time.perf_counter()
    img3, info = lt.render_scene(sc, objs, None, spp=256, rng="philox", seed=it, real="f32", return_info=True)
    t5 = time.perf_counter()
    print(f"iter {it}: Renderer() {1e3*(t1-t0):.1f} ms, render {1e3*(t2-t1):.1f} ms (kernel {c['kernel_ns']/1e6:.1f}), close {1e3*(t3-t2):.1f} ms, "
          f"render_scene(packed) {1e3*(t4-t3):.1f} ms, render_scene(objects) {1e3*(t5-t4):.1f} ms")
