"""Dev: completion / finiteness stress over seeds, scenes, semantics, traversals, precisions and schedulers."""
import sys, time, itertools
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from conftest import load_scene
from light_transport_b200 import host
t0 = time.time()
n = 0
for sf in ("lts_scene.npz", "mirror_scene.npz"):
    ps, nodes, cam, f = load_scene(sf)
    R = host.Renderer(ps, nodes, build_lbvh=True)
    for sem, trav, real, sched in itertools.product(("reference", "iterative"), ("compat", "nearest", "lbvh"), ("f32", "f64"), ("megakernel", "wavefront")):
        for seed, (W, H, spp, D) in enumerate(((33, 17, 5, 16), (64, 64, 3, 256 if sf.startswith("mirror") else 32), (7, 3, 64, 8))):
            film, _, c = R.render(width=W, height=H, spp=spp, max_depth=D, camera=cam, f_distance=f, seed=100 + seed, semantics=sem,
                                  traversal=trav, real=real, scheduler=sched)
            assert np.isfinite(film).all() and c["paths"] == W * H * spp and c["frame_overflow"] == 0, (sf, sem, trav, real, sched, c)
            n += 1
    R.close()
print(f"{n} renders ok in {time.time() - t0:.1f} s")
