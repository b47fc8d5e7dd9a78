"""Dev: where does the end-to-end call spend its time on the 1 M-triangle scene?"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import light_transport_b200 as lt
from light_transport_b200 import api, host, scenes, pack_scene, from_arrays
objs, lights, cam, depth = scenes.lts_scene(api)
ps = from_arrays(scenes.heightfield(nx=1000, ny=500), base=pack_scene(objs, lights))
sc = api.Scene(camera=cam, lights=lights, width=1920, height=1080, max_depth=8, f_distance=depth, number_of_samples=1, allocate=False)
for it in range(4):
    t0 = time.perf_counter()
    R = host.Renderer(ps, None, build_lbvh=True)
    t1 = time.perf_counter()
    img = np.empty((1080, 1920, 3))
    _, _, c = R.render(width=1920, height=1080, spp=4, max_depth=8, camera=cam, f_distance=depth, seed=it, real="f32", image=img, semantics="iterative", traversal="lbvh")
    t2 = time.perf_counter()
    R.close()
    t3 = time.perf_counter()
    img2, info = lt.render_scene(sc, ps, None, spp=4, rng="philox", seed=it, real="f32", return_info=True, semantics="iterative", traversal="lbvh")
    t4 = time.perf_counter()
    print(f"iter {it}: Renderer() {1e3*(t1-t0):.1f} ms, render {1e3*(t2-t1):.1f} ms (kernel {c['kernel_ns']/1e6:.1f}), close {1e3*(t3-t2):.1f} ms, "
          f"render_scene(packed) {1e3*(t4-t3):.1f} ms")
