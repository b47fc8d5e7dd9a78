"""Dev: one render of the LTS scene for profiling: run_once.py <scheduler> <traversal> <real> [size spp]"""
import sys
sys.path.insert(0, ".")
import torch
from light_transport_b200 import api, host, pack_scene, scenes, compat_bvh
if "--lib" in sys.argv:
    k = sys.argv.index("--lib")
    host.LIB_PATH = sys.argv[k + 1]
    del sys.argv[k:k + 2]
objs, lights, cam, depth = scenes.lts_scene(api)
ps = pack_scene(objs, lights)
nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
R = host.Renderer(ps, nodes, build_lbvh=True)
sched, trav, real = sys.argv[1:4]
W = H = int(sys.argv[4]) if len(sys.argv) > 4 else 512
spp = int(sys.argv[5]) if len(sys.argv) > 5 else 64
film = torch.zeros(H * W * 3, dtype=torch.float64, device="cuda")
kw = dict(width=W, height=H, spp=spp, max_depth=16, camera=cam, f_distance=depth, seed=1, semantics="reference", traversal=trav,
          real=real, film=film, scheduler=sched)
for _ in range(2):
    c = R.render(**kw)[2]
print(c)
print("Grays/s %.3f" % ((c["rays_ext"] + c["rays_shadow"]) / c["kernel_ns"]))
