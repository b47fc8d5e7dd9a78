"""Dev: what can a LEAN trace kernel do on the C3 mesh, and what is ray coherence worth?  Times lt_intersect (one thread per ray,
whole-query voted traversal) on camera rays and on the first and second generation of diffuse bounce rays, in path order and
sorted by (origin cell, direction octant).   trace_probe.py [--n 2000000]"""
import argparse, sys
sys.path.insert(0, ".")
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=2_000_000); ap.add_argument("--lib", default=None); ap.add_argument("--off", type=float, default=1e-4)
a = ap.parse_args()
import numpy as np
from light_transport_b200 import api, host, pack_scene, scenes, from_arrays
if a.lib: host.LIB_PATH = a.lib
objs, lights, cam, depth = scenes.lts_scene(api)
ps = from_arrays(scenes.heightfield(nx=1000, ny=500), base=pack_scene(objs, lights))
R = host.Renderer(ps, None, build_lbvh="binary")
rng = np.random.default_rng(1)
W, H = 1920, 1080
n = a.n
pix = np.sort(rng.choice(W * H, n // 4, replace=False)).repeat(4)[:n]          # 4 samples per pixel, pixel order (like a warp's lanes)
i, j = pix // W, pix % W
top = 1 / (W / H)
x = -1 + j * (2 / (W - 1)); y = top - i * (2 * top / (H - 1))
u = rng.random(n)
end = np.stack([x + u / W, y + u / H, np.full(n, depth)], 1)
o = np.tile(np.asarray(cam[:3], float), (n, 1)); d = end - o; d /= np.linalg.norm(d, axis=1, keepdims=True)

def bounce(o, d, prim, t):
    hit = prim >= 0
    p = o + t[:, None] * d
    nn = ps.tri_n[np.maximum(prim, 0)]
    flip = (nn * d).sum(1) > 0
    nn = np.where(flip[:, None], -nn, nn)
    # cosine hemisphere around nn
    r1, r2 = rng.random(len(o)), rng.random(len(o))
    phi = 2 * np.pi * r1; rr = np.sqrt(r2); z = np.sqrt(1 - r2)
    a_ = np.where(np.abs(nn[:, :1]) > 0.9, np.array([[0, 1.0, 0]]), np.array([[1.0, 0, 0]]))
    tx = np.cross(nn, a_); tx /= np.linalg.norm(tx, axis=1, keepdims=True); ty = np.cross(nn, tx)
    nd = rr[:, None] * np.cos(phi)[:, None] * tx + rr[:, None] * np.sin(phi)[:, None] * ty + z[:, None] * nn
    no = p + a.off * nn
    return no[hit], nd[hit]

def sort_rays(o, d, bits=5):
    lo, hi = o.min(0), o.max(0)
    q = np.minimum(((o - lo) / np.maximum(hi - lo, 1e-9) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    def spread(v):
        r = np.zeros_like(v)
        for b in range(bits): r |= ((v >> b) & 1) << (3 * b)
        return r
    key = (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])
    octant = (d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2)
    order = np.argsort((key << 3) | octant, kind="stable")
    return o[order], d[order]

def timeit(name, o, d):
    for trav in ("lbvh2", "lbvh"):
        for srt in (False, True):
            oo, dd = sort_rays(o, d) if srt else (o, d)
            best = None
            for _ in range(3):
                p, t, c = R.intersect(oo, dd, trav, "f32")
                best = c["kernel_ns"] if best is None else min(best, c["kernel_ns"])
            print(f"{name:10s} {trav:6s} {'sorted  ' if srt else 'in order'}: {len(oo) / best * 1e3:8.0f} Mrays/s  box/ray {c['box_tests'] / len(oo):5.1f} tri/ray {c['tri_tests'] / len(oo):.2f}  hit {100 * (p >= 0).mean():.0f} %")
    return R.intersect(o, d, "lbvh2", "f32")[:2]

p0, t0 = timeit("camera", o, d)
o1, d1 = bounce(o, d, p0, t0)
p1, t1 = timeit("bounce 1", o1, d1)
o2, d2 = bounce(o1, d1, p1, t1)
timeit("bounce 2", o2, d2)
