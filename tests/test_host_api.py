"""CPU: host-side logic — the API mirrors reproduce the reference's objects, the packer, the
compat-BVH builder, and the C-ABI library loads and exports what include/lt_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import REPO, load_scene


def test_mirrors_reproduce_reference_scene():
    """scenes.lts_scene built from the plain-Python mirrors == the arrays read off the reference's
    jitclass objects (golden lts_scene.npz), bit for bit — including generate_area_light_samples'
    Q14 quirk and the normals of PreComputedTriangle.__init__."""
    from light_transport_b200 import api, scenes, _pack
    ps_ref, _, cam_ref, f_ref = load_scene("lts_scene.npz")
    objs, lights, cam, depth = scenes.lts_scene(api)
    ps = _pack.pack_scene(objs, lights)
    for k, v in ps_ref.to_npz_dict().items():
        assert np.array_equal(getattr(ps, k), v), k
    assert np.array_equal(cam[:3], cam_ref) and depth == f_ref
    assert ps.n_tris == 28 and ps.n_lights == 2000          # LTS.ipynb cell 19: 28 primitives


def test_mirror_box_matches_golden():
    from light_transport_b200 import api, scenes, _pack
    ps_ref, _, cam_ref, f_ref = load_scene("mirror_scene.npz")
    objs, lights, cam, f = scenes.mirror_box(api)
    ps = _pack.pack_scene(objs, lights)
    for k, v in ps_ref.to_npz_dict().items():
        assert np.array_equal(getattr(ps, k), v), k


@pytest.mark.parametrize("scene_file", ["lts_scene.npz", "mirror_scene.npz"])
def test_compat_builder_equals_reference_node_list(scene_file):
    """compat_bvh.build_compat_nodes == build_bvh + flatten_bvh of the reference (Q7, Q8)."""
    from light_transport_b200 import compat_bvh
    ps, nodes, _, _ = load_scene(scene_file)
    mine = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
    for k in ("bmin", "bmax", "prim_off", "second", "n_prims", "axis"):
        assert np.array_equal(getattr(mine, k), getattr(nodes, k)), k
    assert int(mine.n_prims.sum()) == ps.n_tris            # LTS.ipynb cell 23's sanity check
    assert compat_bvh.compat_defined(mine)


def test_scene_mirror_semantics():
    from light_transport_b200 import api
    np.random.seed(5)
    sc = api.Scene(camera=np.array([0, 0, 8, 1.0]), lights=[], width=6, height=4, max_depth=3, f_distance=7.5,
                   number_of_samples=2)
    np.random.seed(5)
    assert np.array_equal(sc.rand_0, np.random.rand(4, 6, 2, 3))      # R/src/scene.py:68-69 order
    assert np.array_equal(sc.rand_1, np.random.rand(4, 6, 2, 3))
    assert sc.top == 1 / (6 / 4) and sc.bottom == -sc.top and sc.image.shape == (4, 6, 3)
    # R/src/scene.py:70-73: logits of the random numbers and the two side-effect arrays trace_path writes
    assert np.array_equal(sc.rand_0_logit, np.log(sc.rand_0 / (1 - sc.rand_0))) and sc.rand_1_logit.shape == sc.rand_1.shape
    assert sc.bounce_record.dtype == np.int8 and sc.bounce_record.shape == (4, 6, 2, 3) and (sc.bounce_record == 1).all()
    assert sc.record_log_pdf.dtype == np.float64 and (sc.record_log_pdf == 0).all()
    big = api.Scene(camera=np.array([0, 0, 8, 1.0]), lights=[], width=1024, height=1024, max_depth=16,
                    number_of_samples=1024, allocate=False)           # Q15: 704 GB otherwise
    assert big.rand_0 is None


def test_pack_rejects_bad_w():
    from light_transport_b200 import api, _pack
    m = api.Material(color=api.WHITE, shininess=1, reflection=0, ior=1.5)
    t = api.PreComputedTriangle(np.array([0, 0, 0, 1.0]), np.array([1, 0, 0, 1.0]), np.array([0, 1, 0, 0.5]), m)
    with pytest.raises(ValueError):
        _pack.pack_scene([t], [])


def test_heightfield_ingest():
    from light_transport_b200 import scenes, _pack
    base, _, _, _ = load_scene("lts_scene.npz")
    v = scenes.heightfield(nx=20, ny=10)
    ps = _pack.from_arrays(v, base=base)
    assert ps.n_tris == 28 + 400 and ps.tri_mat[-1] == len(base.mat_flags)
    assert np.allclose(np.linalg.norm(ps.tri_n, axis=1), 1)


def test_cabi_exports_every_declared_symbol():
    """liblt_b200.so loads without a GPU and exports each function include/lt_b200.h declares."""
    from light_transport_b200 import host
    hdr = open(os.path.join(REPO, "include", "lt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(lt_[a-z0-9_]+)\s*\(", hdr)))
    assert {"lt_init", "lt_scene_create", "lt_intersect", "lt_render", "lt_bvh_build", "lt_film_finalize"} <= set(names)
    L = host.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in lt_b200.h but not exported"
    assert b"sm_100a" in L.lt_version()


def test_no_cpu_fallback():
    """Without a device the product path fails loudly (never routes through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from light_transport_b200 import host
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    with pytest.raises(host.LtError) as e:
        host.Renderer(ps, nodes)
    assert "no CPU fallback" in str(e.value)
    src = open(os.path.join(REPO, "light_transport_b200", "host.py")).read()
    assert "oracle" not in src.replace("never routes through the oracle", "")


def test_sample_range_partition():
    from light_transport_b200 import parallel
    for spp in (1, 7, 16, 1024, 4096):
        for R in (1, 2, 3, 4, 8):
            rs = [parallel.sample_range(r, R, spp) for r in range(R)]
            assert rs[0][0] == 0 and rs[-1][1] == spp
            assert all(rs[i][1] == rs[i + 1][0] for i in range(R - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_load_obj(tmp_path):
    import light_transport_b200 as lt
    obj = tmp_path / "cube.obj"
    obj.write_text("# unit cube, quads + a degenerate face\n"
                   "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\nv 1 1 1\nv 0 1 1\nvn 0 0 1\n"
                   "f 1 2 3 4\nf 5/1/1 6/2/1 7/3/1 8/4/1\nf 1 2 6 5\nf 2 3 7 6\nf 3 4 8 7\nf -5 -8 -4 -1\nf 1 1 2\n")
    v = lt.load_obj(str(obj), scale=2.0, translate=(0, -5, 0))
    assert v.shape == (12, 3, 3) and v.min() == -5 and v.max() == 2
    base, _, _, _ = load_scene("lts_scene.npz")
    ps = lt.from_arrays(v, base=base, kd=(0.3, 0.4, 0.5))
    assert ps.n_tris == 28 + 12 and np.allclose(ps.mat_kd[-1], (0.3, 0.4, 0.5))
    with pytest.raises(ValueError):
        (tmp_path / "empty.obj").write_text("v 0 0 0\n")
        lt.load_obj(str(tmp_path / "empty.obj"))


def test_film_reduce_rejects_bad_arguments_without_a_gpu():
    """The collective's entry points validate before they touch CUDA or NCCL: null handles come back as LT_E_INVALID."""
    import ctypes as C
    from light_transport_b200 import host
    L = host.lib()
    for name in ("lt_comm_unique_id", "lt_comm_init", "lt_film_reduce", "lt_nccl_version"):
        getattr(L, name).restype = C.c_int
    assert L.lt_film_reduce(None, None, None, C.c_int64(0), C.c_int(0), C.c_int64(0)) == -1
    assert L.lt_comm_init(None, C.c_int(2), C.c_int(0), None, None) == -1
    assert L.lt_comm_unique_id(None) == -1
    assert b"lt_" in L.lt_last_error() or len(L.lt_last_error()) > 0


def test_renderer_close_is_idempotent_and_safe_after_context_loss():
    """a Renderer that outlives its context must not touch freed memory; close() twice is fine (no GPU needed:
    exercised on a stand-in object that skips __init__)."""
    import ctypes as C
    from light_transport_b200 import host
    r = host.Renderer.__new__(host.Renderer)
    r._device, r._ctx, r._scene, r._stream = 0, C.c_void_p(1234), C.c_void_p(5678), None      # a context that is NOT in the cache = gone
    assert r._ctx_gone()
    r.close()                       # must not call into the library with the dangling handles
    assert not r._scene
    r.close()
    with pytest.raises(host.LtError):
        r._bind()


def test_renderer_cache_key_is_the_content():
    """the key render_scene() caches uploaded scenes under: equal arrays -> equal key, any changed value or shape -> another"""
    import copy
    from light_transport_b200 import host
    ps, nodes, _, _ = load_scene("lts_scene.npz")
    k = host._content_key(ps, nodes)
    assert k is not None and k == host._content_key(copy.deepcopy(ps), copy.deepcopy(nodes))
    assert k != host._content_key(ps, None)
    ps2 = copy.deepcopy(ps)
    ps2.tri_v0[3, 1] += 1e-9
    assert k != host._content_key(ps2, nodes)
    n2 = copy.deepcopy(nodes)
    n2.second[0] += 1
    assert k != host._content_key(ps, n2)
    keep = host.RENDERER_CACHE_HASH_BYTES
    host.RENDERER_CACHE_HASH_BYTES = 64
    try:
        assert host._content_key(ps, nodes) is None          # large scenes are keyed by object identity instead
    finally:
        host.RENDERER_CACHE_HASH_BYTES = keep


def test_pin_array_declines_what_it_cannot_pin():
    """arrays too small to matter, too large, or not contiguous are left alone without touching the device library"""
    from light_transport_b200 import host
    assert host.pin_array(np.zeros(10)) is False
    assert host.pin_array(np.zeros((512, 512))[:, ::2]) is False
    assert host.pin_array([1.0, 2.0]) is False
    assert not host._pinned
