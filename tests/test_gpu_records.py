"""GPU (B200): the per-sample side outputs of render_scene_samples and calculate_gradients' replays
 through lt_render_records, against golden vectors of the unmodified reference."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_scene, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R_lts():
    from light_transport_b200 import host
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    r = host.Renderer(ps, nodes, build_lbvh=True)
    yield r
    r.close()


def _inputs(name):
    g = np.load(os.path.join(GOLDEN, name))
    W, H, spp, D = int(g["W"]), int(g["H"]), int(g["spp"]), int(g["D"])
    np.random.seed(int(g["rand_seed"]))
    return g, dict(width=W, height=H, spp=spp, max_depth=D, camera=g["camera"], f_distance=float(g["f_distance"]),
                   rand0=np.random.rand(H, W, spp, D), rand1=np.random.rand(H, W, spp, D), light_idx=g["light_idx"].astype(np.int32))


@pytest.mark.parametrize("name", ["records_lts_d4.npz", "records_lts_d8.npz"])
def test_records_match_reference(R_lts, name):
    g, kw = _inputs(name)
    o = R_lts.render_records(traversal="compat", real="f64", gradients=True, **kw)
    assert np.array_equal(o["bounce_record"], g["bounce_record"])                 # integer output: exact
    assert (rel_err(o["samples"], g["radiance"]) > 1e-5).sum() == 0
    for k, gk in (("record_log_pdf", "record_log_pdf"), ("scene_record_log_pdf", "scene_record_log_pdf"),
                  ("rlp_minus", "rlp_minus"), ("rlp_plus", "rlp_plus")):
        assert np.array_equal(o[k] != 0, g[gk] != 0), k                          # same structure: which bounces recursed
        assert np.allclose(o[k], g[gk], rtol=1e-9, atol=1e-12), k
    assert np.allclose(o["direct_list"], g["direct_list"], rtol=1e-9, atol=1e-12)
    assert np.allclose(o["indirect_list"], g["indirect_list"], rtol=1e-7, atol=1e-10)
    assert np.array_equal(o["record_s"], g["record_s"])
    # finite-difference gradient as the notebook forms it (LTS.ipynb cell 32) agrees too
    D = kw["max_depth"]
    gl_gpu = ((o["rlp_plus"] - o["rlp_minus"]) / (2 * o["record_s"][..., None])).sum(axis=-1)
    gl_ref = ((g["rlp_plus"] - g["rlp_minus"]) / (2 * g["record_s"][..., None])).sum(axis=-1)
    assert np.allclose(gl_gpu, gl_ref, rtol=1e-6, atol=1e-8) and gl_ref.shape[-1] == 2 * D


def test_records_equal_oracle_other_traversals(R_lts, oracle):
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    g, kw = _inputs("records_lts_d8.npz")
    S = oracle.OracleScene(ps, nodes)
    for trav, otrav in (("nearest", "nearest"), ("lbvh", "nearest")):
        o = R_lts.render_records(traversal=trav, real="f64", gradients=True, **kw)
        ref = oracle.render_records(S, traversal=otrav, gradients=True, **kw)
        assert np.array_equal(o["bounce_record"], ref["bounce_record"])
        for k in ("record_log_pdf", "scene_record_log_pdf", "rlp_minus", "rlp_plus", "direct_list"):
            assert np.allclose(o[k], ref[k], rtol=1e-9, atol=1e-12), (trav, k)
        assert o["counters"]["paths"] == ref["counters"]["paths"] * (1 + 4 * kw["max_depth"])


def test_render_scene_full_drop_in():
    """the reference's current render_scene: 10-tuple + Scene side effects, via the mirrors"""
    import light_transport_b200 as lt
    from light_transport_b200 import api, scenes
    objs, lights, cam, depth = scenes.lts_scene(api)
    g, kw = _inputs("records_lts_d4.npz")
    sc = api.Scene(camera=cam, lights=lights, width=kw["width"], height=kw["height"], max_depth=kw["max_depth"], f_distance=depth,
                   number_of_samples=kw["spp"], allocate=False)
    sc.rand_0, sc.rand_1 = kw["rand0"], kw["rand1"]
    sc.bounce_record = np.ones(kw["rand0"].shape, np.int8)
    sc.record_log_pdf = np.zeros(kw["rand0"].shape)
    out = lt.render_scene_full(sc, objs, None, light_idx=kw["light_idx"])
    assert len(out) == 10 and out[8] is None
    image_ver0, image_ver1, image_ver2, samples, rlp, minus, plus, s_set, _, (dl, il) = out
    assert not image_ver0.any() and np.allclose(image_ver1, g["image"], rtol=1e-9, atol=1e-9) and np.array_equal(sc.image, image_ver1)
    assert np.array_equal(sc.bounce_record, g["bounce_record"])
    assert np.allclose(sc.record_log_pdf, g["scene_record_log_pdf"], rtol=1e-9, atol=1e-12)
    assert np.allclose(rlp, g["record_log_pdf"], rtol=1e-9, atol=1e-12) and minus.shape == g["rlp_minus"].shape


def test_records_argument_checks(R_lts):
    from light_transport_b200 import host
    g, kw = _inputs("records_lts_d4.npz")
    with pytest.raises(ValueError):
        R_lts.render_records(**{**kw, "rand1": kw["rand1"][:, :, :1]})
    big = dict(kw, max_depth=40, rand0=np.random.rand(4, 4, 1, 40), rand1=np.random.rand(4, 4, 1, 40), width=4, height=4, spp=1,
               light_idx=None)
    with pytest.raises(host.LtError):
        R_lts.render_records(**big)


def test_more_samples_block_matches_reference(R_lts):
    """render_more_samples (R/src/path_tracing.py:310-364) against golden vectors produced by the reference's own trace_path and
    calculate_gradients: per-sample radiance, the returned record_log_pdf arrays and the 4*D ungated replays of every sample."""
    import os
    import light_transport_b200 as lt
    from light_transport_b200 import api
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "more_samples_lts_d4.npz"))
    W, H, n, D = int(g["W"]), int(g["H"]), int(g["n"]), int(g["D"])
    sc = api.Scene(camera=np.append(g["camera"], 1.0), lights=[], width=W, height=H, max_depth=D, f_distance=float(g["f_distance"]),
                   number_of_samples=1, allocate=False)
    pixels = [tuple(int(v) for v in p) for p in g["pixels"]]
    s2, rlp2, m2, p2, ss2 = lt.render_more_samples(sc, None, None, pixels=pixels, n=n, rand0=g["rand_0_2"], rand1=g["rand_1_2"],
                                                   light_idx=g["light_idx"].astype(np.int32), renderer=R_lts)
    assert s2.shape == (len(pixels), n, 3) and m2.shape == (len(pixels), n, 2 * D, D)
    assert np.allclose(s2, g["samples_2"], rtol=1e-9, atol=1e-12)
    for mine, ref in ((rlp2, g["record_log_pdf_2"]), (m2, g["rlp_minus_2"]), (p2, g["rlp_plus_2"])):
        assert np.array_equal(mine == 0, ref == 0)
        assert np.allclose(mine, ref, rtol=1e-9, atol=1e-12)
    assert np.array_equal(ss2, g["record_s_2"])
