"""Dev: throughput of lt_render_records (side outputs + the 4*D gradient replays) on the LTS scene."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from light_transport_b200 import api, host, pack_scene, scenes, compat_bvh
objs, lights, cam, depth = scenes.lts_scene(api)
ps = pack_scene(objs, lights)
nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
R = host.Renderer(ps, nodes)
W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
spp, D = 4, 8
rng = np.random.default_rng(0)
r0, r1 = rng.random((H, W, spp, D)), rng.random((H, W, spp, D))
li = rng.integers(0, len(lights), (H, W, spp, 40)).astype(np.int32)
for real in ("f64", "f32"):
    for grad in (False, True):
        for it in range(2):
            t0 = time.perf_counter()
            out = R.render_records(width=W, height=H, spp=spp, max_depth=D, camera=cam, f_distance=depth, rand0=r0, rand1=r1,
                                   light_idx=li, gradients=grad, real=real)
            dt = time.perf_counter() - t0
        c = out["counters"] if "counters" in out else {}
        n = H * W * spp * (1 + 4 * D if grad else 1)
        print(f"{real} gradients={grad}: {n/1e6:.2f} M walks, wall {dt*1e3:.1f} ms, kernel {c.get('kernel_ns', 0)/1e6:.2f} ms, "
              f"{(c.get('rays_ext',0)+c.get('rays_shadow',0))/max(c.get('kernel_ns',1),1)*1e3:.0f} Mrays/s")
