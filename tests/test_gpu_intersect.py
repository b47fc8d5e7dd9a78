"""GPU (B200): ray/scene intersection through the C-ABI (lt_intersect) against the golden vectors
of the reference and against the oracle.  Hit primitive ids must be bit-exact in float64; float32
disagreements are counted, reported and bounded (north_star: 'grazing-edge disagreements are counted
and reported')."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, RAY_SETS, load_scene, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderers():
    from light_transport_b200 import host
    out = {}
    for sf, _ in RAY_SETS:
        ps, nodes, cam, f = load_scene(sf)
        out[sf] = host.Renderer(ps, nodes, build_lbvh="binary")
    yield out
    for r in out.values():
        r.close()


@pytest.mark.parametrize("scene_file,ray_file", RAY_SETS)
def test_compat_f64_bit_exact(renderers, scene_file, ray_file):
    r = np.load(os.path.join(GOLDEN, ray_file))
    prim, t, cnt = renderers[scene_file].intersect(r["o"], r["d"], "compat", "f64")
    assert np.array_equal(prim, r["compat_prim"]), f"{(prim != r['compat_prim']).sum()} ids differ"
    hit = prim >= 0
    assert np.array_equal(np.isinf(t), ~hit)
    assert np.array_equal(t[hit], r["compat_t"][hit])         # same operations, same roundings
    assert cnt["tri_tests"] > 0 and cnt["box_tests"] > 0


@pytest.mark.parametrize("trav", ["nearest", "lbvh", "lbvh2"])
@pytest.mark.parametrize("scene_file,ray_file", RAY_SETS)
def test_nearest_f64_bit_exact(renderers, scene_file, ray_file, trav):
    r = np.load(os.path.join(GOLDEN, ray_file))
    prim, t, _ = renderers[scene_file].intersect(r["o"], r["d"], trav, "f64")
    assert np.array_equal(prim, r["nearest_prim"]), f"{(prim != r['nearest_prim']).sum()} ids differ"
    hit = prim >= 0
    assert np.array_equal(t[hit], r["nearest_t"][hit])


@pytest.mark.parametrize("trav,key", [("compat", "compat"), ("nearest", "nearest"), ("lbvh", "nearest"), ("lbvh2", "nearest")])
@pytest.mark.parametrize("scene_file,ray_file", RAY_SETS)
def test_f32_disagreements_counted(renderers, scene_file, ray_file, trav, key, record_property):
    r = np.load(os.path.join(GOLDEN, ray_file))
    prim, t, _ = renderers[scene_file].intersect(r["o"], r["d"], trav, "f32")
    ref_p, ref_t = r[key + "_prim"], r[key + "_t"]
    bad = prim != ref_p
    record_property("f32_disagreements", int(bad.sum()))
    print(f"{scene_file} {trav} f32: {bad.sum()} / {len(prim)} primitive ids differ from the float64 reference")
    assert bad.mean() < 0.02                      # only rays that graze edges/ties in float32
    ok = ~bad & (prim >= 0)
    assert np.allclose(t[ok], ref_t[ok], rtol=2e-4, atol=1e-4)
    # f32 LBVH and f32 all-primitive scan implement the same semantics in the same arithmetic
    if trav in ("lbvh", "lbvh2"):
        p2, t2, _ = renderers[scene_file].intersect(r["o"], r["d"], "nearest", "f32")
        assert (p2 != prim).mean() < 2e-3


def test_edge_cases(renderers):
    R = renderers["lts_scene.npz"]
    # empty batch
    p, t, _ = R.intersect(np.zeros((0, 3)), np.zeros((0, 3)), "compat", "f64")
    assert len(p) == 0
    # rays that leave through the open front of the box miss; axis-parallel rays (inf in 1/d) work
    o = np.array([[5, 5, 0.0], [5, 5, 0], [5, 5, 0], [0, 0, 8.0]])
    d = np.array([[0, 0, 1.0], [0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    for trav in ("compat", "nearest", "lbvh", "lbvh2"):
        p, t, _ = R.intersect(o, d, trav, "f64")
        assert p[0] == -1 and np.isinf(t[0]) and p[3] == -1
        assert p[1] in (6, 7) and abs(t[1] - 12.5) < 1e-12          # floor
        assert p[2] in (0, 1) and abs(t[2] - 2.5) < 1e-12           # right wall
    with pytest.raises(ValueError):
        R.intersect(np.zeros((2, 3)), np.zeros((3, 3)))


def test_compat_undefined_is_an_error():
    """Q9b: where the reference raises IndexError, lt_scene_set_compat_bvh refuses the node list."""
    from light_transport_b200 import host, compat_bvh
    ps, nodes, _, _ = load_scene("lts_scene.npz")
    bad = compat_bvh.CompatNodes(nodes.bmin[:20].copy(), nodes.bmax[:20].copy(), nodes.prim_off[:20].copy(),
                                 nodes.second[:20].copy(), nodes.n_prims[:20].copy(), nodes.axis[:20].copy())
    bad.n_prims[19] = 1; bad.prim_off[19] = 27; bad.second[:] = np.minimum(bad.second, 19)
    with pytest.raises(host.LtError) as e:
        host.Renderer(ps, bad)
    assert e.value.code in (-3, -1)
    with pytest.raises(host.LtError) as e:      # traversal without its structure
        R = host.Renderer(ps)
        R.intersect(np.zeros((1, 3)), np.ones((1, 3)), "compat")
    assert e.value.code == -4


def test_lbvh_mesh_vs_oracle(oracle):
    """C3-class scene (displaced-grid mesh, here 2 x 240 x 120 = 57.6k triangles + the Cornell box):
    GPU LBVH build + traversal == oracle brute force (float64, ids and t bit-exact) on mixed rays,
    and == the GPU's own all-primitive scan on many more."""
    from light_transport_b200 import host, scenes, _pack
    base, _, cam, f = load_scene("lts_scene.npz")
    ps = _pack.from_arrays(scenes.heightfield(nx=240, ny=120), base=base)
    R = host.Renderer(ps, None, build_lbvh=True)
    st = R.lbvh_stats
    assert st["n_prims"] == ps.n_tris and 0 < st["n_nodes"] < ps.n_tris // 2 and 3 <= st["max_depth"] <= 28
    assert 10 < st["binary_depth"] <= 96 and st["sah_cost"] > 0
    rng = np.random.default_rng(5)
    n = 60000
    o = np.concatenate([np.tile(cam, (n // 2, 1)), rng.uniform(-7, 7, (n // 2, 3))])
    d = rng.normal(size=(n, 3)); d[: n // 2, 2] = -np.abs(d[: n // 2, 2]) - 0.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p_l, t_l, c_l = R.intersect(o, d, "lbvh", "f64")
    p_n, t_n, c_n = R.intersect(o, d, "nearest", "f64")
    assert np.array_equal(p_l, p_n) and np.array_equal(t_l, t_n)
    assert c_l["tri_tests"] < c_n["tri_tests"] / 100          # the hierarchy actually culls
    S = oracle.OracleScene(ps, None)
    k = 3000
    p_o, t_o, _ = oracle.intersect(S, o[:k], d[:k], "nearest")
    assert np.array_equal(p_l[:k], p_o) and np.array_equal(t_l[:k][p_o >= 0], t_o[p_o >= 0])
    assert (p_l >= 28).mean() > 0.2                            # the mesh is actually being hit
    R.close()


@pytest.mark.parametrize("binary_leaf,leaf_max", [(1, 3), (2, 3), (3, 3), (2, 2), (3, 1)])
def test_binary_tree_leaf_sizes(binary_leaf, leaf_max):
    """The binary tree kept for `lbvh2` holds only nodes with more than `binary_leaf` triangles below them (densely renumbered)
    and multi-triangle leaf references (position | count) into the 8-wide hierarchy's leaf order: every leaf size, including
    the cap at the slot size, gives the all-primitive scan's answer bit for bit — closest hits and any-hit shadow walks."""
    from light_transport_b200 import host, scenes, _pack
    base, _, cam, f = load_scene("lts_scene.npz")
    ps = _pack.from_arrays(scenes.heightfield(nx=60, ny=40), base=base)
    R = host.Renderer(ps, None)
    st = R.build_lbvh(keep_binary=True, leaf_max=leaf_max, binary_leaf=binary_leaf)
    eff = min(binary_leaf, leaf_max)
    n_nodes = (st["binary_bytes"] - 48 * ps.n_tris) // 64
    assert (n_nodes == ps.n_tris - 1) if eff == 1 else (ps.n_tris / (2 * eff) < n_nodes < ps.n_tris - 1)
    rng = np.random.default_rng(11)
    n = 20000
    o = np.concatenate([np.tile(cam, (n // 2, 1)), rng.uniform(-7, 7, (n // 2, 3))])
    d = rng.normal(size=(n, 3)); d[: n // 2, 2] = -np.abs(d[: n // 2, 2]) - 0.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p_b, t_b, _ = R.intersect(o, d, "lbvh2", "f64")
    p_n, t_n, _ = R.intersect(o, d, "nearest", "f64")
    assert np.array_equal(p_b, p_n) and np.array_equal(t_b, t_n) and (p_b >= 28).mean() > 0.2
    kw = dict(width=48, height=32, spp=4, max_depth=6, camera=cam, f_distance=f, semantics="iterative", seed=9, real="f64", return_samples=True)
    _, s_b, c_b = R.render(traversal="lbvh2", **kw)
    _, s_n, c_n = R.render(traversal="lbvh", **kw)             # the 8-wide hierarchy: same semantics, different tree
    assert np.array_equal(s_b, s_n) and c_b["rays_shadow"] == c_n["rays_shadow"] and c_b["rays_ext"] == c_n["rays_ext"]
    R.close()


@pytest.mark.parametrize("seed,n_soup", [(1, 300), (2, 3000), (3, 20000)])
def test_hierarchies_on_triangle_soup_with_ties(seed, n_soup):
    """Unstructured input for the hierarchy walks: random triangles of mixed sizes, exact duplicates (tie on t: the lowest
    primitive index must win whatever leaf either copy sits in), zero-area triangles (never hit) and a few scene-sized ones.
    Both hierarchies, every binary leaf size: ids and t of the all-primitive scan, bit for bit in float64."""
    from light_transport_b200 import host, _pack
    base, _, cam, f = load_scene("lts_scene.npz")
    rng = np.random.default_rng(seed)
    c = rng.uniform(-6, 6, (n_soup, 1, 3))
    size = np.where(rng.random((n_soup, 1, 1)) < 0.02, 6.0, rng.choice([0.05, 0.3, 1.0], (n_soup, 1, 1)))
    v = c + size * rng.normal(size=(n_soup, 3, 3))
    n_dup, n_deg = n_soup // 10, n_soup // 50
    v[rng.choice(n_soup, n_dup, replace=False)] = v[rng.choice(n_soup, n_dup)]            # exact duplicates
    deg = rng.choice(n_soup, n_deg, replace=False)
    v[deg, 2] = v[deg, 1]                                                                   # zero area
    with np.errstate(invalid="ignore", divide="ignore"):
        ps = _pack.from_arrays(v, base=base)
    R = host.Renderer(ps, None)
    n = 30000
    o = rng.uniform(-7, 7, (n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    o[: n // 3] = np.asarray(cam, dtype=np.float64)[:3]
    p_n, t_n, _ = R.intersect(o, d, "nearest", "f64")
    assert (p_n >= 28).mean() > 0.15 and not np.isin(p_n, 28 + deg).any()
    for leaf_max, binary_leaf in ((3, 2), (3, 3), (1, 1), (2, 1)):
        R.build_lbvh(keep_binary=True, leaf_max=leaf_max, binary_leaf=binary_leaf)
        for trav in ("lbvh", "lbvh2"):
            p, t, _ = R.intersect(o, d, trav, "f64")
            assert np.array_equal(p, p_n), f"{trav} leaf {leaf_max}/{binary_leaf}: {(p != p_n).sum()} ids differ"
            assert np.array_equal(t, t_n)
    R.close()


def test_compat_beyond_the_small_scene_paths(oracle):
    """COMPAT on a scene with more than 32 leaves / referenced primitives (the Cornell scene + a 4x4 displaced grid = 60
    triangles): the kernels' small-scene tables (32-bit masks, two tests per round, descent chains) do not apply and the
    general 64-bit leaf walk runs.  Ids and t bit-exact against the oracle in float64, walk replayed on Philox streams."""
    from light_transport_b200 import host, scenes, _pack, compat_bvh
    base, _, cam, f = load_scene("lts_scene.npz")
    ps = _pack.from_arrays(scenes.heightfield(nx=4, ny=4), base=base)
    nodes = compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
    n_leaves = int((nodes.n_prims > 0).sum())
    assert ps.n_tris == 60 and 32 < n_leaves <= 64 and len(nodes.n_prims) <= 128
    R = host.Renderer(ps, nodes)
    S = oracle.OracleScene(ps, nodes)
    rng = np.random.default_rng(11)
    n = 20000
    o = np.concatenate([np.tile(cam, (n // 2, 1)), rng.uniform(-6, 6, (n // 2, 3))])
    d = rng.normal(size=(n, 3)); d[: n // 2, 2] = -np.abs(d[: n // 2, 2]) - 0.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p_o, t_o, _ = oracle.intersect(S, o, d, "compat")
    p_g, t_g, _ = R.intersect(o, d, "compat", "f64")
    assert np.array_equal(p_g, p_o) and np.array_equal(t_g[p_o >= 0], t_o[p_o >= 0])
    assert (p_o >= 28).mean() > 0.05                           # the added triangles are being hit
    p_f, _, _ = R.intersect(o, d, "compat", "f32")
    assert (p_f != p_o).mean() < 0.02
    kw = dict(width=48, height=32, spp=4, max_depth=8, camera=cam, f_distance=f, seed=5, semantics="reference", traversal="compat")
    for sched in ("megakernel", "wavefront"):
        film, samples, cnt = R.render(real="f64", return_samples=True, scheduler=sched, **kw)
        o_film, o_samples, o_cnt = oracle.render(S, return_samples=True, **kw)
        assert (rel_err(samples, o_samples) > 1e-9).sum() == 0
        assert cnt["rays_ext"] == o_cnt["rays_ext"] and cnt["rays_shadow"] == o_cnt["rays_shadow"]
    R.close()


# ---------------------------------------------------------------------------------------------------------------------
# C3 at its BASELINE size: 1,000,028 triangles

@pytest.fixture(scope="module")
def c3_scene():
    from light_transport_b200 import host, scenes, _pack
    base, _, cam, f = load_scene("lts_scene.npz")
    ps = _pack.from_arrays(scenes.heightfield(nx=1000, ny=500), base=base)
    R = host.Renderer(ps, None, build_lbvh="binary")
    yield R, ps, cam, f
    R.close()


def _c3_rays(ps, cam, n, seed):
    """camera rays, rays from random interior points, rays aimed at Light samples"""
    rng = np.random.default_rng(seed)
    k = n // 3
    o = np.concatenate([np.tile(cam, (k, 1)), rng.uniform(-7, 7, (n - k, 3))])
    d = rng.normal(size=(n, 3))
    d[:k, 2] = -np.abs(d[:k, 2]) - 0.5
    tgt = ps.light_pos[rng.integers(0, len(ps.light_pos), n - 2 * k)]
    d[2 * k:] = tgt - o[2 * k:]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d


def test_c3_full_size_hierarchy(c3_scene):
    R, ps, cam, f = c3_scene
    st = R.lbvh_stats
    print("C3 hierarchy:", st)
    assert st["n_prims"] == ps.n_tris == 1_000_028
    assert st["max_depth"] <= 28 and st["binary_depth"] <= 96
    assert st["bytes"] <= 64 << 20                      # nodes + float32 triangle records stay well inside the L2
    assert st["sah_cost"] > 0


def test_c3_full_size_f64_equals_oracle_brute_force(c3_scene, oracle):
    """lt_intersect through the 8-wide hierarchy, float64: primitive ids AND t bit-exact against the oracle's brute force
    over all 1,000,028 triangles (R/src/intersects.py:45-104 on every primitive, EPSILON < t, lowest index wins ties)."""
    R, ps, cam, f = c3_scene
    o, d = _c3_rays(ps, cam, 3000, seed=21)
    p_l, t_l, c_l = R.intersect(o, d, "lbvh", "f64")
    p_o, t_o, _ = oracle.intersect(oracle.OracleScene(ps, None), o, d, "nearest")
    assert np.array_equal(p_l, p_o), f"{(p_l != p_o).sum()} ids differ"
    assert np.array_equal(t_l[p_o >= 0], t_o[p_o >= 0])
    assert (p_o >= 28).mean() > 0.2 and (p_o < 0).mean() < 0.5
    assert c_l["tri_tests"] < 64 * len(o)               # the hierarchy culls: a handful of triangles per ray, not a million


def test_c3_full_size_f64_equals_gpu_scan_and_f32_counted(c3_scene, record_property):
    """100k rays: the hierarchy == the GPU's own all-primitive scan (float64, ids and t); float32 disagreements with the
    float64 result are COUNTED, classified and bounded (north_star: 'grazing-edge disagreements are counted and reported')."""
    R, ps, cam, f = c3_scene
    o, d = _c3_rays(ps, cam, 100_000, seed=22)
    p_l, t_l, _ = R.intersect(o, d, "lbvh", "f64")
    p_n, t_n, _ = R.intersect(o, d, "nearest", "f64")
    assert np.array_equal(p_l, p_n) and np.array_equal(t_l, t_n)
    p_2, t_2, _ = R.intersect(o, d, "lbvh2", "f64")                          # the binary tree the hierarchy is collapsed from
    assert np.array_equal(p_2, p_n) and np.array_equal(t_2, t_n)
    p_f, t_f, _ = R.intersect(o, d, "lbvh", "f32")
    bad = p_f != p_l
    both = bad & (p_f >= 0) & (p_l >= 0)
    same_t = np.isclose(t_f[both], t_l[both], rtol=1e-4, atol=1e-4)          # a neighbouring triangle at the same distance: an edge
    n_edge, n_other = int(same_t.sum()), int(bad.sum() - same_t.sum())
    record_property("c3_f32_disagreements", int(bad.sum()))
    print(f"C3 1M triangles, f32 vs f64 over {len(o)} rays: {bad.sum()} ids differ ({n_edge} shared-edge neighbours at the same t, "
          f"{n_other} hit/miss or occluder flips)")
    assert bad.mean() < 5e-3 and n_other <= max(20, n_edge)
    ok = ~bad & (p_l >= 0)
    assert np.allclose(t_f[ok], t_l[ok], rtol=2e-4, atol=1e-4)


def test_c3_walk_f64_replays_oracle(c3_scene, oracle):
    """ITERATIVE semantics (R/src/path_tracing_fix1.py:17-136) on the 1 M-triangle scene through the GPU hierarchy,
    float64, Philox: every path equals the oracle's (which intersects by brute force) and so do the ray counts."""
    R, ps, cam, f = c3_scene
    kw = dict(width=64, height=36, spp=1, max_depth=8, camera=cam, f_distance=f, semantics="iterative", seed=3)
    film, samples, cnt = R.render(traversal="lbvh", real="f64", return_samples=True, **kw)
    _, o_samples, o_cnt = oracle.render(oracle.OracleScene(ps, None), traversal="nearest", return_samples=True, **kw)
    assert (rel_err(samples, o_samples) > 1e-5).sum() == 0
    assert cnt["rays_ext"] == o_cnt["rays_ext"] and cnt["rays_shadow"] == o_cnt["rays_shadow"]
    # float32 walk of the same streams: a few paths branch differently at edges; counted and bounded
    for trav in ("lbvh", "lbvh2"):
        _, s32, _ = R.render(traversal=trav, real="f32", return_samples=True, **kw)
        differ = (rel_err(s32, o_samples) > 1e-3).any(axis=-1).mean()
        print(f"C3 walk, {trav}: {100 * differ:.2f} % of float32 paths differ from the float64 oracle by more than 1e-3")
        assert differ < 0.02
