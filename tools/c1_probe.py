import sys
sys.path.insert(0, ".")
import numpy as np, torch
from light_transport_b200 import api, host, pack_scene, scenes, compat_bvh
objs, lights, cam, depth = scenes.lts_scene(api)
ps = pack_scene(objs, lights)
nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
R = host.Renderer(ps, nodes)
film = torch.zeros(256 * 256 * 3, dtype=torch.float64, device="cuda")
kw = dict(width=256, height=256, spp=16, max_depth=8, camera=cam, f_distance=depth, semantics="reference", traversal="compat", real="f32", film=film)
for ppw in (0, 16, 32, 64, 128, 256, 512, 1024):
    for ct in (0, 3, 4):
        ts = []
        for i in range(12):
            c = R.render(seed=i, paths_per_warp=ppw, ctas_per_sm=ct, **kw)[2]
            ts.append(c["kernel_ns"])
        rays = c["rays_ext"] + c["rays_shadow"]
        print(f"ppw {ppw:5d} ctas {ct}: kernel {np.median(ts[2:]) / 1e3:7.1f} us  {rays / np.median(ts[2:]) * 1e3:8.0f} Mrays/s")
