"""Dev: one render of the C3 mesh scene for profiling: run_mesh.py [scheduler] [spp]"""
import sys
sys.path.insert(0, ".")
import torch
from light_transport_b200 import api, host, pack_scene, scenes, from_arrays
objs, lights, cam, depth = scenes.lts_scene(api)
ps = from_arrays(scenes.heightfield(nx=1000, ny=500), base=pack_scene(objs, lights))
R = host.Renderer(ps, None, build_lbvh=True)
sched = sys.argv[1] if len(sys.argv) > 1 else "megakernel"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
film = torch.zeros(1080 * 1920 * 3, dtype=torch.float64, device="cuda")
kw = dict(width=1920, height=1080, spp=spp, max_depth=8, camera=cam, f_distance=depth, seed=3, semantics="iterative", traversal="lbvh",
          real="f32", film=film, scheduler=sched)
for _ in range(2):
    c = R.render(**kw)[2]
rays = c["rays_ext"] + c["rays_shadow"]
print(R.lbvh_stats, c, rays / c["kernel_ns"] * 1e3, "Mrays/s")
