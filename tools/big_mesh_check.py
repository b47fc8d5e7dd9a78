import sys
sys.path.insert(0, ".")
import numpy as np, time
from light_transport_b200 import api, host, pack_scene, scenes, from_arrays
objs, lights, cam, depth = scenes.lts_scene(api)
base = pack_scene(objs, lights)
for nx, ny in ((2000, 1000), (4000, 2000)):
    t0 = time.time()
    ps = from_arrays(scenes.heightfield(nx=nx, ny=ny), base=base)
    R = host.Renderer(ps, None)
    st = R.build_lbvh(keep_binary=True)
    print(f"{ps.n_tris} triangles: host {time.time()-t0:.1f} s, build {st['build_ms']:.1f} ms, wide nodes {st['n_nodes']}, depth {st['max_depth']}/{st['binary_depth']}, bytes {st['bytes']/1e6:.0f}/{st['binary_bytes']/1e6:.0f} MB")
    rng = np.random.default_rng(1)
    n = 200000
    o = np.concatenate([np.tile(np.asarray(cam, dtype=np.float64)[:3], (n // 2, 1)), rng.uniform(-7, 7, (n // 2, 3))])
    d = rng.normal(size=(n, 3)); d[: n // 2, 2] = -np.abs(d[: n // 2, 2]) - 0.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p8, t8, _ = R.intersect(o, d, "lbvh", "f64")
    p2, t2, _ = R.intersect(o, d, "lbvh2", "f64")
    k = 2000
    pn, tn, _ = R.intersect(o[:k], d[:k], "nearest", "f64")
    print("  lbvh == lbvh2:", np.array_equal(p8, p2) and np.array_equal(t8, t2), " == scan on", k, "rays:", np.array_equal(p8[:k], pn) and np.array_equal(t8[:k], tn), " mesh hits", float((p8 >= 28).mean()))
    kw = dict(width=960, height=540, spp=8, max_depth=8, camera=cam, f_distance=depth, seed=3, semantics="iterative", real="f32")
    for trav in ("lbvh2", "lbvh"):
        for _ in range(2):
            c = R.render(traversal=trav, **kw)[2]
        rays = c["rays_ext"] + c["rays_shadow"]
        print(f"  {trav}: {rays / c['kernel_ns'] * 1e3:8.1f} Mrays/s  box/ray {c['box_tests']/rays:.1f} tri/ray {c['tri_tests']/rays:.2f}")
    R.close()
