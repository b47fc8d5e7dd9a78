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
chunks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
for ch in chunks:
    for _ in range(2):
        c = R.render(lbvh_chunk=ch, **kw)[2]
    rays = c["rays_ext"] + c["rays_shadow"]
    print(f"chunk {ch:3d}: {c['kernel_ns']/1e6:8.2f} ms  {rays / c['kernel_ns'] * 1e3:8.1f} Mrays/s  box/ray {c['box_tests']/rays:.1f}")
