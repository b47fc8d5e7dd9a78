"""Dev: throughput of lt_render_records (per-sample outputs + the 4*D finite-difference replays per sample, R/src/path_tracing.py:203-303)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from conftest import load_scene
from light_transport_b200 import host
ps, nodes, cam, f = load_scene("lts_scene.npz")
R = host.Renderer(ps, nodes)
W = H = 256; spp = 16; D = 8
rng = np.random.default_rng(1)
r0, r1 = rng.random((H, W, spp, D)), rng.random((H, W, spp, D))
li = rng.integers(0, ps.n_lights, (H, W, spp, 32), dtype=np.int32)
for grads in (False, True):
    for real in ("f64", "f32"):
        kw = dict(width=W, height=H, spp=spp, max_depth=D, camera=cam, f_distance=f, rand0=r0, rand1=r1, light_idx=li, real=real, gradients=grads)
        R.render_records(**kw)
        t0 = time.perf_counter()
        out = R.render_records(**kw)
        dt = time.perf_counter() - t0
        c = out.get("counters", {})
        walks = H * W * spp * (1 + (4 * D if grads else 0))
        print(f"gradients={grads} {real}: {dt*1e3:8.1f} ms wall, kernel {c.get('kernel_ns', 0)/1e6:8.1f} ms, {walks/1e6:.1f} M walks (upper bound: gated replays are skipped), "
              f"rays {c.get('rays_ext',0)+c.get('rays_shadow',0)}, {(c.get('rays_ext',0)+c.get('rays_shadow',0))/max(c.get('kernel_ns',1),1)*1e3:.0f} Mrays/s")
