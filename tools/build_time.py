import sys
sys.path.insert(0, ".")
from light_transport_b200 import api, host, pack_scene, scenes, from_arrays
objs, lights, cam, depth = scenes.lts_scene(api)
base = pack_scene(objs, lights)
for nx, ny in ((1000, 500), (2000, 1000), (2828, 1414), (4000, 2000)):
    ps = from_arrays(scenes.heightfield(nx=nx, ny=ny), base=base)
    R = host.Renderer(ps, None)
    for kb in (False, True):
        st = R.build_lbvh(keep_binary=kb)
        st = R.build_lbvh(keep_binary=kb)
        print(ps.n_tris, "binary" if kb else "wide  ", {k: round(st[k], 2) for k in ("build_ms", "sort_ms", "collapse_ms")})
    R.close()
