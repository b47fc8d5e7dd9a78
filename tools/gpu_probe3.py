"""Dev probe: megakernel vs wavefront scheduler, f32, LTS scene, REFERENCE semantics, D=16."""
import sys
sys.path.insert(0, ".")
import torch
from light_transport_b200 import api, host, pack_scene, scenes, compat_bvh
objs, lights, cam, depth = scenes.lts_scene(api)
ps = pack_scene(objs, lights)
nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
R = host.Renderer(ps, nodes, build_lbvh=True)
W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
film = torch.zeros(H * W * 3, dtype=torch.float64, device="cuda")
for real in ("f32", "f64"):
    for trav in ("compat", "nearest", "lbvh"):
        for sched in ("megakernel", "wavefront"):
            kw = dict(width=W, height=H, spp=spp, max_depth=16, camera=cam, f_distance=depth, seed=1, semantics="reference",
                      traversal=trav, real=real, film=film, scheduler=sched)
            R.render(**kw)
            best = min((R.render(**kw)[2] for _ in range(3)), key=lambda c: c["kernel_ns"])
            rays = best["rays_ext"] + best["rays_shadow"]
            ms = best["kernel_ns"] / 1e6
            print(f"{real} {trav:7s} {sched:10s}: {ms:8.2f} ms {rays/ms/1e3:8.1f} Mrays/s slot-util {best['lane_steps']/max(best['lane_slots'],1):.3f}")
