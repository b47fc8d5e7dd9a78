"""CPU: the oracle (oracle/lt_oracle.c) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import PATH_SETS, RAY_SETS, load_paths, load_scene, GOLDEN
import os


@pytest.mark.parametrize("scene_file,ray_file", RAY_SETS)
def test_intersect_bit_exact(oracle, scene_file, ray_file):
    ps, nodes, _, _ = load_scene(scene_file)
    S = oracle.OracleScene(ps, nodes)
    r = np.load(os.path.join(GOLDEN, ray_file))
    # COMPAT = the reference's intersect_bvh (R/src/bvh_new.py:413-482): ids and t bit-exact
    p, t, _ = oracle.intersect(S, r["o"], r["d"], "compat")
    assert np.array_equal(p, r["compat_prim"])
    assert np.array_equal(t, r["compat_t"])
    # NEAREST = brute force over the reference's triangle_intersect (R/src/intersects.py:45-104)
    p, t, _ = oracle.intersect(S, r["o"], r["d"], "nearest")
    assert np.array_equal(p, r["nearest_prim"])
    assert np.array_equal(t, r["nearest_t"])
    # the two semantics do differ (Q8/Q9) — the fixture must exercise that
    assert (r["compat_prim"] != r["nearest_prim"]).sum() > 0


@pytest.mark.parametrize("scene_file,path_file", PATH_SETS)
def test_trace_path_replay_bit_exact(oracle, scene_file, path_file):
    ps, nodes, _, _ = load_scene(scene_file)
    S = oracle.OracleScene(ps, nodes)
    g = load_paths(path_file)
    film, samples, cnt = oracle.render(S, width=g["W"], height=g["H"], spp=g["spp"], max_depth=g["D"], camera=g["camera"],
                                       f_distance=g["f_distance"], semantics=g["semantics"], traversal="compat",
                                       rand0=g["rand0"], rand1=g["rand1"], light_idx=g["light_idx"], return_samples=True)
    assert cnt["light_idx_overflow"] == 0
    assert np.array_equal(samples, g["radiance"])      # bit-exact, not merely 1e-5
    assert np.allclose(film, samples.sum(axis=2), rtol=1e-13, atol=1e-13)
    img = np.clip(film / g["spp"], 0, 1)
    assert np.allclose(img, g["image"], rtol=1e-12, atol=1e-12)
    assert cnt["paths"] == g["W"] * g["H"] * g["spp"]


def test_compat_undefined_raises(oracle):
    """Q9b: a leaf holding a primitive index >= node count makes the reference raise IndexError."""
    from light_transport_b200 import api, scenes, _pack, compat_bvh
    # 2-triangle walls share an AABB centroid -> land in one leaf -> 14 prims / fewer nodes
    d = 5.0
    ns = api
    mat = ns.Material(color=ns.WHITE_2, shininess=30, reflection=0.1, ior=1.5)
    objs = []
    A, B, Cc, D = [-d, -d, -d], [-d, -d, d], [d, -d, d], [d, -d, -d]
    E, F, G, H = [-d, d, -d], [-d, d, d], [d, d, d], [d, d, -d]
    for q in ([D, Cc, G, H], [F, B, A, E], [E, A, D, H], [A, B, Cc, D], [H, G, F, E], [B, F, G, Cc]):
        scenes._quad(ns, objs, q, mat)
    lamp = ns.Material(color=ns.WHITE, shininess=1, reflection=0.9, ior=1.5, emission=200)
    objs.append(scenes._tri(ns, [-1, 4.5, -1], [1, 4.5, 1], [-1, 4.5, 1], lamp, True))
    objs.append(scenes._tri(ns, [-1, 4.5, -1], [1, 4.5, -1], [1, 4.5, 1], lamp, True))
    lights = [ns.Light(source=np.array([0, 4.5, 0, 1.0]), material=lamp, normal=objs[-1].normal, total_area=4)]
    ps = _pack.pack_scene(objs, lights)
    nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
    assert len(nodes) < ps.n_tris and not compat_bvh.compat_defined(nodes)
    S = oracle.OracleScene(ps, nodes)
    o = np.array([[0.0, 0.0, 0.0]])
    dd = np.array([[0.0, 1.0, 0.0]])
    with pytest.raises(IndexError):
        oracle.intersect(S, o, dd, "compat")


def test_philox_known_answers(oracle):
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    L = oracle.lib()
    def ph(ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        L.lto_philox4x32_10(c, k, o)
        return [int(x) for x in o]
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_oracle_properties(oracle):
    """size-independent properties: sample-range additivity, Philox determinism, film = sum of samples."""
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    S = oracle.OracleScene(ps, nodes)
    kw = dict(width=24, height=16, spp=8, max_depth=8, camera=cam, f_distance=f, seed=7)
    full, s_full, c_full = oracle.render(S, return_samples=True, **kw)
    a, _, ca = oracle.render(S, s0=0, s1=3, **kw)
    b, _, cb = oracle.render(S, s0=3, s1=8, **kw)
    assert np.allclose(a + b, full, rtol=1e-12, atol=1e-12)
    assert ca["rays_ext"] + cb["rays_ext"] == c_full["rays_ext"]
    again, s_again, _ = oracle.render(S, return_samples=True, **kw)
    assert np.array_equal(s_again, s_full)
    other, _, _ = oracle.render(S, **{**kw, "seed": 8})
    assert not np.array_equal(other, full)
    # empty sample range and zero depth are legal and give black
    z, _, cz = oracle.render(S, s0=4, s1=4, **kw)
    assert not z.any() and cz["paths"] == 0
    z, _, _ = oracle.render(S, **{**kw, "max_depth": 0})
    assert not z.any()


@pytest.mark.parametrize("name", ["records_lts_d4.npz", "records_lts_d8.npz"])
def test_records_and_gradients_bit_exact(oracle, name):
    """per-sample side outputs of trace_path (returned record_log_pdf, the arrays it
    writes into Scene, the root frame's direct/indirect lists) and calculate_gradients' 4*D replays
    (R/src/path_tracing.py:203-249), against the unmodified reference."""
    ps, nodes, _, _ = load_scene("lts_scene.npz")
    g = np.load(os.path.join(GOLDEN, name))
    W, H, spp, D = int(g["W"]), int(g["H"]), int(g["spp"]), int(g["D"])
    np.random.seed(int(g["rand_seed"]))
    r0, r1 = np.random.rand(H, W, spp, D), np.random.rand(H, W, spp, D)
    out = oracle.render_records(oracle.OracleScene(ps, nodes), width=W, height=H, spp=spp, max_depth=D, camera=g["camera"],
                                f_distance=float(g["f_distance"]), rand0=r0, rand1=r1, light_idx=g["light_idx"].astype(np.int32),
                                gradients=True)
    for k, gk in (("samples", "radiance"), ("record_log_pdf", "record_log_pdf"), ("scene_record_log_pdf", "scene_record_log_pdf"),
                  ("bounce_record", "bounce_record"), ("direct_list", "direct_list"), ("indirect_list", "indirect_list"),
                  ("rlp_minus", "rlp_minus"), ("rlp_plus", "rlp_plus"), ("record_s", "record_s")):
        assert np.array_equal(out[k], g[gk]), k
    assert (g["rlp_minus"] != g["rlp_plus"]).any()        # the replays do differ: the fixture exercises them


def test_more_samples_block_bit_exact(oracle):
    """R/src/path_tracing.py:310-364: extra samples at fixed pixels, each with 4*D UNGATED replays whose "logits" are the
    uniforms themselves (:347) — golden vectors from the reference's own trace_path + calculate_gradients."""
    from conftest import load_scene
    g = np.load(os.path.join(GOLDEN, "more_samples_lts_d4.npz"))
    ps, nodes, _, _ = load_scene("lts_scene.npz")
    S = oracle.OracleScene(ps, nodes)
    W, H, n, D = int(g["W"]), int(g["H"]), int(g["n"]), int(g["D"])
    for k, (i, j) in enumerate(g["pixels"]):
        pix = int(i) * W + int(j)
        r0, r1 = g["rand_0_2"][k:k + 1], g["rand_1_2"][k:k + 1]
        o = oracle.render_records(S, width=W, height=H, spp=n, max_depth=D, camera=g["camera"], f_distance=float(g["f_distance"]), rand0=r0,
                                  rand1=r1, light_idx=g["light_idx"][k:k + 1].astype(np.int32), gradients=True, pixel_range=(pix, pix + 1),
                                  ungated=True, logits=(r0, r1))
        assert np.array_equal(o["samples"][0], g["samples_2"][k])
        assert np.array_equal(o["record_log_pdf"][0], g["record_log_pdf_2"][k])
        assert np.array_equal(o["rlp_minus"][0], g["rlp_minus_2"][k]) and np.array_equal(o["rlp_plus"][0], g["rlp_plus_2"][k])
        assert np.array_equal(o["record_s"][0], g["record_s_2"][k])
    assert np.abs(g["rlp_minus_2"]).sum() > 0 and not np.array_equal(g["rlp_minus_2"], g["rlp_plus_2"])


def test_render_pixels_equals_render_on_a_subset(oracle):
    """lto_render_pixels (per-pixel sums and sums of squares for single pixels of a big film) == lto_render on those pixels."""
    from conftest import load_scene
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    S = oracle.OracleScene(ps, nodes)
    kw = dict(width=40, height=28, spp=6, max_depth=8, camera=cam, f_distance=f, seed=4)
    px = np.array([5 * 40 + 7, 100, 1119, 0], np.int32)
    a, b, c = oracle.render_pixels(S, px, **kw)
    film, samples, _ = oracle.render(S, return_samples=True, **kw)
    sub = samples.reshape(-1, 6, 3)[px]
    assert np.allclose(a, sub.sum(1), rtol=1e-14, atol=1e-14) and np.allclose(b, (sub ** 2).sum(1), rtol=1e-14, atol=1e-14)
    assert c["paths"] == len(px) * 6
