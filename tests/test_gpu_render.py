"""GPU (B200): the path walk through the C-ABI (lt_render).
 * replayed-RNG mode: per-path radiance within 1e-5 relative of the reference's trace_path
   (golden vectors) — north_star's second correctness check;
 * Philox mode: the oracle draws the same counter-based streams, so float64 GPU paths replay it
   path by path; float32 agrees within 3 sigma of the Monte Carlo standard error per pixel;
 * size-independent properties at larger sizes (sample-range additivity = virtual ranks, film = sum of
   samples, determinism)."""
import numpy as np
import pytest

from conftest import PATH_SETS, load_paths, load_scene, rel_err

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5          # BASELINE.json north_star: "per-path radiance must agree within 1e-5 relative"


@pytest.fixture(scope="module")
def R_lts():
    from light_transport_b200 import host
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    r = host.Renderer(ps, nodes, build_lbvh="binary")
    yield r
    r.close()


@pytest.fixture(scope="module")
def R_mirror():
    from light_transport_b200 import host
    ps, nodes, cam, f = load_scene("mirror_scene.npz")
    r = host.Renderer(ps, nodes, build_lbvh=True)
    yield r
    r.close()


def _renderer(scene_file, R_lts, R_mirror):
    return R_lts if scene_file.startswith("lts") else R_mirror


@pytest.mark.parametrize("scene_file,path_file", PATH_SETS)
def test_replay_f64_matches_reference(R_lts, R_mirror, scene_file, path_file):
    R = _renderer(scene_file, R_lts, R_mirror)
    g = load_paths(path_file)
    film, samples, cnt = R.render(width=g["W"], height=g["H"], spp=g["spp"], max_depth=g["D"], camera=g["camera"],
                                  f_distance=g["f_distance"], semantics=g["semantics"], traversal="compat", real="f64",
                                  rand0=g["rand0"], rand1=g["rand1"], light_idx=g["light_idx"], return_samples=True)
    err = rel_err(samples, g["radiance"])
    n_bad = int((err > REL_TOL).any(axis=-1).sum())
    print(f"{path_file}: max rel err {err.max():.3g}, paths beyond 1e-5: {n_bad} / {err.shape[0]*err.shape[1]*err.shape[2]}")
    assert n_bad == 0
    assert cnt["light_idx_overflow"] == 0 and cnt["frame_overflow"] == 0
    assert cnt["paths"] == g["W"] * g["H"] * g["spp"]
    assert np.allclose(film, samples.sum(axis=2), rtol=1e-12, atol=1e-12)
    img = R.finalize(film, g["spp"])
    assert np.allclose(img, g["image"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("scene_file,path_file", PATH_SETS)
def test_replay_control_flow_equals_oracle(R_lts, R_mirror, oracle, scene_file, path_file):
    """identical ray counts => identical control flow (every roulette, every recursion) on every path"""
    R = _renderer(scene_file, R_lts, R_mirror)
    ps, nodes, _, _ = load_scene(scene_file)
    g = load_paths(path_file)
    kw = dict(width=g["W"], height=g["H"], spp=g["spp"], max_depth=g["D"], camera=g["camera"], f_distance=g["f_distance"],
              semantics=g["semantics"], traversal="compat", rand0=g["rand0"], rand1=g["rand1"], light_idx=g["light_idx"])
    _, _, c_gpu = R.render(real="f64", **kw)
    _, _, c_cpu = oracle.render(oracle.OracleScene(ps, nodes), **kw)
    for k in ("paths", "rays_ext", "rays_shadow", "max_frames"):
        assert c_gpu[k] == c_cpu[k], k


@pytest.mark.parametrize("semantics", ["reference", "iterative"])
@pytest.mark.parametrize("trav", ["compat", "nearest", "lbvh", "lbvh2"])
def test_philox_f64_replays_oracle(R_lts, oracle, trav, semantics):
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    kw = dict(width=40, height=24, spp=6, max_depth=8, camera=cam, f_distance=f, semantics=semantics, seed=11)
    film, samples, cnt = R_lts.render(traversal=trav, real="f64", return_samples=True, **kw)
    S = oracle.OracleScene(ps, nodes)
    o_film, o_samples, o_cnt = oracle.render(S, traversal="compat" if trav == "compat" else "nearest", return_samples=True, **kw)
    err = rel_err(samples, o_samples)
    assert (err > REL_TOL).sum() == 0, f"max {err.max()}"
    assert cnt["rays_ext"] == o_cnt["rays_ext"] and cnt["rays_shadow"] == o_cnt["rays_shadow"]


def test_deep_walk_frame_arena(R_mirror, oracle):
    """C4a: max depth 256 — frames beyond the in-kernel stack spill to the arena and still replay."""
    ps, nodes, cam, f = load_scene("mirror_scene.npz")
    kw = dict(width=32, height=32, spp=4, max_depth=256, camera=cam, f_distance=f, semantics="reference", seed=5)
    film, samples, cnt = R_mirror.render(traversal="compat", real="f64", return_samples=True, **kw)
    _, o_samples, o_cnt = oracle.render(oracle.OracleScene(ps, nodes), traversal="compat", return_samples=True, **kw)
    assert (rel_err(samples, o_samples) > REL_TOL).sum() == 0
    assert cnt["max_frames"] == o_cnt["max_frames"] and cnt["rays_ext"] == o_cnt["rays_ext"]


def _z_stats(a, b, spp, mask=None):
    am, bm = a.mean(axis=2), b.mean(axis=2)
    se = np.sqrt(a.var(axis=2, ddof=1) / spp + b.var(axis=2, ddof=1) / spp)
    valid = se > 1e-9
    if mask is not None:
        valid &= mask[:, :, None]
    z = np.abs(am - bm)[valid] / se[valid]
    return z, float(am[valid].mean()), float(bm[valid].mean()), float(np.sqrt((se[valid] ** 2).mean() / valid.sum()))


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("trav", ["compat", "lbvh"])
def test_philox_independent_streams_within_3_sigma(R_lts, oracle, trav, real):
    """north_star's third check: with INDEPENDENT Philox streams per-pixel estimates agree within 3 sigma
    of the Monte Carlo standard error.  GPU (seed 101) vs oracle float64 (seed 202); the radiance
    distribution is heavy-tailed (lamp hits are worth 200), so the admissible fraction beyond 3 sigma is
    calibrated by the same statistic between two oracle runs (seeds 202 vs 303).

    Grazing-edge pixels are counted separately: Q13 jitters x and y by the SAME number, so every sample
    of a pixel on the image's ANTI-diagonal (x == y) travels in the plane x = y and meets the box on a
    wall/wall edge — a tie that float64 resolves by last-bit rounding and float32 by different rounding.
    Those W-of-W^2 pixels are excluded from the float32 bound and reported."""
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    W, H, spp = 32, 32, 256
    kw = dict(width=W, height=H, spp=spp, max_depth=8, camera=cam, f_distance=f, semantics="reference", return_samples=True)
    otrav = "compat" if trav == "compat" else "nearest"
    S = oracle.OracleScene(ps, nodes)
    _, g_s, _ = R_lts.render(traversal=trav, real=real, seed=101, **kw)
    _, o_s, _ = oracle.render(S, traversal=otrav, seed=202, **kw)
    _, o2_s, _ = oracle.render(S, traversal=otrav, seed=303, **kw)
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    diag = ii + jj == W - 1            # the exact-tie pixels: x == y only on the ANTI-diagonal (on the main one y = -x + 2r/W)
    mask = ~diag if real == "f32" else None
    z, gm, om, se_mean = _z_stats(g_s, o_s, spp, mask)
    z0, _, _, _ = _z_stats(o2_s, o_s, spp, mask)
    zd, _, _, _ = _z_stats(g_s, o_s, spp, diag)
    frac, frac0 = (z > 3).mean(), (z0 > 3).mean()
    print(f"{trav}/{real}: {len(z)} pixel-channels; beyond 3 sigma: GPU-vs-oracle {frac*100:.2f}%, oracle-vs-oracle {frac0*100:.2f}%; "
          f"image means {gm:.5f} vs {om:.5f} (se {se_mean:.5f}); grazing-edge (diagonal) pixel-channels: {len(zd)}, "
          f"{(zd > 3).sum()} beyond 3 sigma")
    assert frac <= frac0 + 0.006
    assert abs(gm - om) < 4 * se_mean


def test_f32_same_stream_close_to_f64(R_lts):
    """float32 walks the same Philox numbers (top 24 bits); most paths take the same decisions."""
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    kw = dict(width=48, height=48, spp=4, max_depth=8, camera=cam, f_distance=f, seed=3, traversal="compat", return_samples=True)
    _, s32, c32 = R_lts.render(real="f32", **kw)
    _, s64, c64 = R_lts.render(real="f64", **kw)
    close = (rel_err(s32, s64, floor=1e-2) < 1e-3).all(axis=-1)
    print(f"f32 vs f64 same stream: {close.mean()*100:.2f}% of paths within 1e-3")
    assert close.mean() > 0.97
    assert abs(c32["rays_ext"] - c64["rays_ext"]) / c64["rays_ext"] < 0.01


def test_virtual_ranks_equal_full(R_lts):
    """multi-GPU semantics on one GPU: disjoint sample ranges rendered separately and summed == one render
    (SURVEY.md §4).  Also: deterministic per-path results, order-independent film within fp-sum tolerance."""
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    from light_transport_b200 import parallel
    kw = dict(width=64, height=40, spp=16, max_depth=8, camera=cam, f_distance=f, seed=9, traversal="lbvh", real="f32")
    full, s_full, c_full = R_lts.render(return_samples=True, **kw)
    for world in (2, 4, 8):
        acc = np.zeros_like(full)
        rays = 0
        for r in range(world):
            s0, s1 = parallel.sample_range(r, world, 16)
            part, s_part, c = R_lts.render(s0=s0, s1=s1, return_samples=True, **kw)
            assert np.array_equal(s_part, s_full[:, :, s0:s1])          # same Philox counters => same paths
            acc += part
            rays += c["rays_ext"] + c["rays_shadow"]
        assert np.allclose(acc, full, rtol=1e-12, atol=1e-12)
        assert rays == c_full["rays_ext"] + c_full["rays_shadow"]
    again, s_again, _ = R_lts.render(return_samples=True, **kw)
    assert np.array_equal(s_again, s_full)


def test_film_accumulates_and_pixel_ranges(R_lts):
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    kw = dict(width=32, height=20, spp=8, max_depth=8, camera=cam, f_distance=f, seed=2, traversal="nearest", real="f32")
    full, _, _ = R_lts.render(**kw)
    film = np.zeros_like(full)
    R_lts.render(film=film, pixel_range=(0, 300), **kw)
    R_lts.render(film=film, pixel_range=(300, 640), **kw)
    assert np.allclose(film, full, rtol=1e-12, atol=1e-12)
    z, _, c = R_lts.render(s0=3, s1=3, **kw)                   # empty sample range
    assert not z.any() and c["paths"] == 0
    z, _, _ = R_lts.render(**{**kw, "max_depth": 0})
    assert not z.any()
    one, _, c = R_lts.render(**{**kw, "width": 1, "height": 1, "spp": 1})   # 1x1 film: linspace degenerates to start
    assert c["paths"] == 1


def test_render_scene_drop_in(oracle):
    """the reference-facing call: render_scene(scene, primitives, bvh) with the mirrors, replay mode"""
    import light_transport_b200 as lt
    from light_transport_b200 import api, scenes
    objs, lights, cam, depth = scenes.lts_scene(api)
    np.random.seed(1)
    sc = api.Scene(camera=cam, lights=lights, width=24, height=24, max_depth=4, f_distance=depth, number_of_samples=2)
    img, samples, info = lt.render_scene(sc, objs, None, light_idx=np.zeros((24, 24, 2, 32), np.int32), return_samples=True,
                                         return_info=True)
    assert img.shape == (24, 24, 3) and np.array_equal(sc.image, img) and img.min() >= 0 and img.max() <= 1
    ps = lt.pack_scene(objs, lights)
    S = oracle.OracleScene(ps, lt.compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2))
    _, o_s, _ = oracle.render(S, width=24, height=24, spp=2, max_depth=4, camera=cam, f_distance=depth, rand0=sc.rand_0,
                              rand1=sc.rand_1, light_idx=np.zeros((24, 24, 2, 32), np.int32), return_samples=True)
    assert (rel_err(samples, o_s) > REL_TOL).sum() == 0
    # philox path for configs whose Scene arrays cannot exist (Q15)
    big = api.Scene(camera=cam, lights=lights, width=64, height=64, max_depth=8, f_distance=depth, number_of_samples=1,
                    allocate=False)
    img2 = lt.render_scene(big, objs, None, spp=32, rng="philox", seed=4, traversal="lbvh")
    assert img2.shape == (64, 64, 3) and 0.05 < img2.mean() < 0.9


# ---------------------------------------------------------------- wavefront scheduler (LT_SCHED_WAVEFRONT)
@pytest.mark.parametrize("scene_file,path_file", PATH_SETS)
def test_wavefront_replay_f64_matches_reference(R_lts, R_mirror, scene_file, path_file):
    """the CTA-resident wavefront scheduler runs the same per-path arithmetic: replay parity holds"""
    R = _renderer(scene_file, R_lts, R_mirror)
    g = load_paths(path_file)
    kw = dict(width=g["W"], height=g["H"], spp=g["spp"], max_depth=g["D"], camera=g["camera"], f_distance=g["f_distance"],
              semantics=g["semantics"], traversal="compat", real="f64", rand0=g["rand0"], rand1=g["rand1"],
              light_idx=g["light_idx"], return_samples=True)
    film, samples, cnt = R.render(scheduler="wavefront", **kw)
    assert (rel_err(samples, g["radiance"]) > REL_TOL).sum() == 0
    assert cnt["paths"] == g["W"] * g["H"] * g["spp"] and cnt["frame_overflow"] == 0
    assert np.allclose(film, samples.sum(axis=2), rtol=1e-12, atol=1e-12)
    _, s_mk, c_mk = R.render(scheduler="megakernel", **kw)
    assert np.array_equal(samples, s_mk)                      # identical per-path results, bit for bit
    # (tri_tests is not compared: a shadow query stops testing at its first hit, and how many tests are already in
    #  flight at that point differs between the schedulers — two per round, or a pooled warp-wide round)
    for k in ("rays_ext", "rays_shadow", "max_frames"):
        assert cnt[k] == c_mk[k], k


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("semantics", ["reference", "iterative"])
@pytest.mark.parametrize("trav", ["compat", "nearest", "lbvh"])
def test_wavefront_equals_megakernel(R_lts, trav, semantics, real):
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    kw = dict(width=96, height=64, spp=12, max_depth=16, camera=cam, f_distance=f, semantics=semantics, seed=21,
              traversal=trav, real=real, return_samples=True)
    f_w, s_w, c_w = R_lts.render(scheduler="wavefront", **kw)
    f_m, s_m, c_m = R_lts.render(scheduler="megakernel", **kw)

    def same(a, b):
        if real == "f64":            # built with -fmad=false: both kernels round identically
            return np.array_equal(a, b)
        # the fast-math float32 unit may contract/reorder differently per kernel: equal up to rare decision flips
        return (rel_err(a, b, floor=1e-2) < 1e-3).all(axis=-1).mean() > 0.995

    assert same(s_w, s_m)
    assert c_w["paths"] == c_m["paths"]
    if real == "f64":
        assert np.allclose(f_w, f_m, rtol=1e-12, atol=1e-12)
        assert c_w["rays_ext"] == c_m["rays_ext"] and c_w["rays_shadow"] == c_m["rays_shadow"]
    else:
        assert abs(c_w["rays_ext"] - c_m["rays_ext"]) < 2e-3 * c_m["rays_ext"]
    # ragged job sizes: fewer paths than one CTA pool, sample sub-range, pixel sub-range
    for extra in (dict(width=5, height=3, spp=2), dict(s0=5, s1=7), dict(pixel_range=(100, 1000))):
        a = R_lts.render(scheduler="wavefront", **{**kw, **extra})
        b = R_lts.render(scheduler="megakernel", **{**kw, **extra})
        assert same(a[1], b[1]) and a[2]["paths"] == b[2]["paths"]


def test_full_size_properties(R_lts):
    """BASELINE's C2 film (1024x1024) at a bounded sample count: the size-independent properties hold at full
    film size — sample-range additivity (the multi-GPU decomposition), determinism, counters add up, the image
    is a clipped mean, both schedulers agree statistically."""
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    W = H = 1024
    kw = dict(width=W, height=H, spp=1024, max_depth=16, camera=cam, f_distance=f, seed=5, traversal="compat", real="f32")
    full, _, c_full = R_lts.render(s0=0, s1=8, **kw)
    acc = np.zeros_like(full)
    rays = 0
    for a, b in ((0, 3), (3, 4), (4, 8)):
        part, _, c = R_lts.render(s0=a, s1=b, **kw)
        acc += part
        rays += c["rays_ext"] + c["rays_shadow"]
    assert np.allclose(acc, full, rtol=1e-12, atol=1e-9)
    assert rays == c_full["rays_ext"] + c_full["rays_shadow"] and c_full["paths"] == W * H * 8
    again, _, c2 = R_lts.render(s0=0, s1=8, **kw)
    assert np.allclose(again, full, rtol=1e-12, atol=1e-9) and c2["rays_ext"] == c_full["rays_ext"]
    img = R_lts.finalize(full, 8)
    assert img.min() >= 0 and img.max() <= 1 and np.allclose(img, np.clip(full / 8, 0, 1))
    wf, _, c_wf = R_lts.render(s0=0, s1=8, scheduler="wavefront", **kw)
    assert c_wf["paths"] == c_full["paths"] and abs(wf.mean() - full.mean()) < 2e-3 * full.mean()
    assert 7.0 < (c_full["rays_ext"] + c_full["rays_shadow"]) / c_full["paths"] < 7.7       # SURVEY: ~8 rays per path


def test_ragged_and_degenerate_inputs(oracle):
    """non-square films, a single-triangle scene (LBVH without a hierarchy), one light (no light RNG), odd sizes"""
    from light_transport_b200 import api, host, _pack, compat_bvh
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    R = host.Renderer(ps, nodes, build_lbvh=True)
    S = oracle.OracleScene(ps, nodes)
    for (W, H, spp, D) in ((37, 5, 3, 4), (1, 64, 2, 8), (64, 1, 5, 2), (3, 3, 1, 1)):
        kw = dict(width=W, height=H, spp=spp, max_depth=D, camera=cam, f_distance=f, seed=W * 131 + H)
        _, s_g, c_g = R.render(traversal="compat", real="f64", return_samples=True, **kw)
        _, s_o, c_o = oracle.render(S, traversal="compat", return_samples=True, **kw)
        assert (rel_err(s_g, s_o) > REL_TOL).sum() == 0, (W, H)
        assert c_g["rays_ext"] == c_o["rays_ext"] and c_g["rays_shadow"] == c_o["rays_shadow"]
    R.close()
    # one emissive triangle facing the camera + one Light: every traversal agrees with the oracle
    lamp = api.Material(color=api.WHITE, shininess=1, reflection=0.9, ior=1.5, emission=5.0)
    tri = api.PreComputedTriangle(np.array([-4, -4, 0, 1.0]), np.array([4, -4, 0, 1.0]), np.array([0, 4, 0, 1.0]), lamp, True)
    light = api.Light(source=np.array([0, 0, 3, 1.0]), material=lamp, normal=tri.normal, total_area=1.0)
    one = _pack.pack_scene([tri], [light])
    n1 = compat_bvh.build_compat_nodes(one.tri_v0, one.tri_v1, one.tri_v2)
    assert len(n1) == 1 and compat_bvh.compat_defined(n1)
    R1 = host.Renderer(one, n1, build_lbvh=True)
    S1 = oracle.OracleScene(one, n1)
    kw = dict(width=16, height=16, spp=4, max_depth=4, camera=np.array([0, 0, 8.0]), f_distance=7.5, seed=3)
    _, s_o, _ = oracle.render(S1, traversal="compat", return_samples=True, **kw)
    assert s_o.max() > 0
    for trav in ("compat", "nearest", "lbvh"):
        for sched in ("megakernel", "wavefront"):
            _, s_g, _ = R1.render(traversal=trav, real="f64", scheduler=sched, return_samples=True, **kw)
            assert (rel_err(s_g, s_o) > REL_TOL).sum() == 0, (trav, sched)
    R1.close()


@pytest.mark.parametrize("scene_file", ["lts_scene.npz", "mirror_scene.npz"])
@pytest.mark.parametrize("trav", ["compat", "nearest"])
def test_scheduling_policies_do_not_change_paths(R_lts, R_mirror, scene_file, trav):
    """Phase-aligned iterations, granule size and scheduler only decide WHEN a lane traces and shades: every path's
    float64 radiance and the ray counts are identical whichever policy runs (the host picks the phase policy from the
    scene's diffuse share of surface area; here both are forced)."""
    R = _renderer(scene_file, R_lts, R_mirror)
    ps, nodes, cam, f = load_scene(scene_file)
    kw = dict(width=40, height=24, spp=6, max_depth=32, camera=cam, f_distance=f, seed=13, semantics="reference", traversal=trav,
              real="f64", return_samples=True)
    _, s_on, c_on = R.render(phased=True, **kw)
    _, s_off, c_off = R.render(phased=False, **kw)
    _, s_small, c_small = R.render(paths_per_warp=64, **kw)
    assert np.array_equal(s_on, s_off) and np.array_equal(s_on, s_small)
    for k in ("paths", "rays_ext", "rays_shadow", "rays_memo", "max_frames"):
        assert c_on[k] == c_off[k] == c_small[k], k


@pytest.mark.parametrize("semantics", ["reference", "iterative"])
def test_zero_pdf_sample_still_casts_its_shadow_ray(R_lts, oracle, semantics):
    """rand_0 == 0 puts the hemisphere sample on the disk's rim: pdf == 0.  The reference has already cast the shadow ray by
    then (R/src/path_tracing.py:70-80; path_tracing_fix1.py:56-67 has even added the direct light), so the query counts and,
    for ITERATIVE, its light is kept.  Replayed against the oracle in float64; float32 keeps the same ray counts."""
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    W, H, spp, D = 12, 10, 2, 8
    rng = np.random.default_rng(17)
    r0, r1 = rng.random((H, W, spp, D)), rng.random((H, W, spp, D))
    r0[:, :, 0, 1] = 0.0; r1[:, :, 0, 1] = 0.5            # sample 0: zero pdf at bounce 1
    r0[:, :, 1, 0] = 0.0; r1[:, :, 1, 0] = 0.5            # sample 1: at bounce 0 (jitter 0 as well)
    kw = dict(width=W, height=H, spp=spp, max_depth=D, camera=cam, f_distance=f, semantics=semantics, traversal="compat",
              rand0=r0, rand1=r1, seed=9, return_samples=True)
    _, s_gpu, c_gpu = R_lts.render(real="f64", **kw)
    _, s_cpu, c_cpu = oracle.render(oracle.OracleScene(ps, nodes), **kw)
    assert (rel_err(s_gpu, s_cpu) > 1e-9).sum() == 0
    for k in ("rays_ext", "rays_shadow"):
        assert c_gpu[k] == c_cpu[k], k
    _, _, c32 = R_lts.render(real="f32", **kw)
    assert abs(c32["rays_shadow"] - c_cpu["rays_shadow"]) <= 2 and abs(c32["rays_ext"] - c_cpu["rays_ext"]) <= 4


def test_c2_float32_acceptance_at_size(R_lts, oracle, record_property):
    """SURVEY.md 8(d) row C2 / north_star's third check on the BENCHMARKED arm: config C2 itself (1024 x 1024 film, 1024 spp,
    D = 16, REFERENCE semantics, COMPAT traversal), float32 GPU against the float64 oracle on 4096 random pixels with
    INDEPENDENT Philox streams: per-pixel means within 3 sigma of the Monte Carlo standard error.  The radiance is heavy
    tailed (lamp hits are worth 200), so the admissible fraction beyond 3 sigma is calibrated by the same statistic between
    two oracle runs.  Pixels on the image's anti-diagonal (Q13: x == y there, every sample meets the box on a wall/wall edge, and
    even the float64 reference decides those hits by last-bit rounding: 12 % of them differ between its own COMPAT and brute-force
    answers and 2 % escape through the edge) are reported separately and held to a looser bound."""
    import torch
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    W = H = 1024
    spp, D, n_pix = 1024, 16, 4096
    rng = np.random.default_rng(2024)
    pixels = np.sort(rng.choice(W * H, n_pix, replace=False)).astype(np.int32)
    i, j = pixels // W, pixels % W
    diag = i + j == W - 1             # exact ties exist only on the ANTI-diagonal: there x == y, so d.x == d.y (the main diagonal has y = -x + 2r/W)
    extra = np.array([k * W + (W - 1 - k) for k in range(0, W, 8)], dtype=np.int32)   # 128 anti-diagonal pixels
    S = oracle.OracleScene(ps, nodes)
    kw = dict(width=W, height=H, spp=spp, max_depth=D, camera=cam, f_distance=f, semantics="reference", traversal="compat")
    allpix = np.concatenate([pixels, extra])
    a1, b1, _ = oracle.render_pixels(S, allpix, seed=202, **kw)
    a2, b2, _ = oracle.render_pixels(S, allpix, seed=303, **kw)
    film = torch.zeros(H * W * 3, dtype=torch.float64, device="cuda")
    for p in allpix:
        R_lts.render(real="f32", seed=101, film=film, pixel_range=(int(p), int(p) + 1), **kw)
    g = film.cpu().numpy().reshape(-1, 3)[allpix]

    def z(am, bm, asq):      # difference of two means over the standard error (variance from the float64 run)
        var = np.maximum(asq / spp - (am / spp) ** 2, 0) * spp / (spp - 1)
        se = np.sqrt(2 * var / spp)
        ok = se > 1e-9
        return (np.abs(am - bm) / spp)[ok] / se[ok]

    n = n_pix
    nd = ~diag
    z_gpu = z(a1[:n][nd], g[:n][nd], b1[:n][nd])
    z_cal = z(a1[:n][nd], a2[:n][nd], b1[:n][nd])
    frac, frac0 = float((z_gpu > 3).mean()), float((z_cal > 3).mean())
    dm = np.concatenate([np.arange(n)[diag], np.arange(n, n + len(extra))])
    z_dg, z_dc = z(a1[dm], g[dm], b1[dm]), z(a1[dm], a2[dm], b1[dm])
    fd, fd0 = float((z_dg > 3).mean()), float((z_dc > 3).mean())
    mean_gpu, mean_cpu = g[:n].sum() / (n * spp * 3), a1[:n].sum() / (n * spp * 3)
    print(f"C2 float32 vs float64 oracle, {nd.sum()} off-diagonal pixels x {spp} spp: beyond 3 sigma {100 * frac:.2f} % "
          f"(oracle vs oracle: {100 * frac0:.2f} %), max z {z_gpu.max():.1f} ({z_cal.max():.1f}); mean radiance {mean_gpu:.5f} vs {mean_cpu:.5f}; "
          f"{len(dm)} diagonal pixels: {100 * fd:.2f} % ({100 * fd0:.2f} %)")
    record_property("c2_f32_beyond_3sigma", frac)
    assert frac <= frac0 + 0.004 and frac < 0.02
    assert abs(mean_gpu - mean_cpu) < 4 * np.sqrt((b1[:n].sum() / (n * spp * 3) - mean_cpu ** 2) / (n * spp * 3)) + 1e-4
    # exact-tie pixels (W of W^2): every sample of an anti-diagonal pixel meets the box on a wall/wall edge; float64 resolves the
    # tie like the reference (same last-bit roundings), float32 rounds the candidates apart.  Counted and reported, not hidden.
    record_property("c2_f32_diagonal_beyond_3sigma", fd)
    assert fd < 0.8


def test_image_rmse_float32_vs_float64(R_lts):
    """Image-level check at 256 x 256: the float32 image of a seed is far closer to the float64 image of the SAME seed than two
    float64 images of different seeds are to each other (float32 error is well below the Monte Carlo noise)."""
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    kw = dict(width=256, height=256, spp=256, max_depth=16, camera=cam, f_distance=f, semantics="reference", traversal="compat")
    img = {}
    for name, real, seed in (("f32", "f32", 5), ("f64", "f64", 5), ("f64b", "f64", 6)):
        film, _, _ = R_lts.render(real=real, seed=seed, **kw)
        img[name] = R_lts.finalize(film, 256)
    ii, jj = np.mgrid[0:256, 0:256]
    diag = ii + jj == 255                           # exact-tie pixels (Q13: x == y), 0.4 % of this film: float32 splits their samples between the two walls
    rmse = lambda a, b, m: float(np.sqrt(np.mean((a - b)[m] ** 2)))
    e_prec, e_noise = rmse(img["f32"], img["f64"], ~diag), rmse(img["f64"], img["f64b"], ~diag)
    e_diag, e_diag_noise = rmse(img["f32"], img["f64"], diag), rmse(img["f64"], img["f64b"], diag)
    print(f"256x256x256spp, off the anti-diagonal: RMSE(f32, f64 same seed) = {e_prec:.2e}, RMSE(f64 seed a, f64 seed b) = {e_noise:.2e}; "
          f"on its {diag.sum()} pixels: {e_diag:.2e} vs {e_diag_noise:.2e}")
    assert e_prec < 0.25 * e_noise
    assert e_diag < 0.3                             # reported: a wall/wall blend instead of the reference's one-wall tie break


def test_fix1_film_blend(oracle):
    """R/src/path_tracing_fix1.py:165-166: scene.image accumulates 0.25 * clip(mean, 0, 1) per call (ITERATIVE semantics)."""
    import light_transport_b200 as lt
    from light_transport_b200 import api, scenes
    objs, lights, cam, depth = scenes.lts_scene(api)
    np.random.seed(3)
    sc = api.Scene(camera=cam, lights=lights, width=20, height=14, max_depth=6, f_distance=depth, number_of_samples=3)
    plain = lt.render_scene(sc, objs, None, semantics="iterative").copy()           # overwrites scene.image (path_tracing_old style)
    sc.image[...] = 0.125
    out = lt.render_scene(sc, objs, None, semantics="iterative", film_blend=0.25)
    assert np.allclose(out, 0.125 + 0.25 * plain, rtol=0, atol=1e-15) and np.array_equal(sc.image, out)
    out2 = lt.render_scene(sc, objs, None, semantics="iterative", film_blend=0.25)  # the notebook calls it repeatedly
    assert np.allclose(out2, 0.125 + 0.5 * plain, rtol=0, atol=1e-15)


def test_empty_pixel_share_renders_nothing(R_lts):
    """Only pixel_range (0, 0) means the whole film; an empty share (k, k) renders nothing, a reversed one is refused."""
    from light_transport_b200 import host
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    kw = dict(width=16, height=12, spp=2, max_depth=4, camera=cam, f_distance=f, seed=2, real="f64")
    film, _, cnt = R_lts.render(pixel_range=(5, 5), **kw)
    assert cnt["paths"] == 0 and not film.any()
    film, _, cnt = R_lts.render(pixel_range=(0, 0), **kw)
    assert cnt["paths"] == 16 * 12 * 2
    with pytest.raises(host.LtError):
        R_lts.render(pixel_range=(7, 3), **kw)


def test_render_scene_keeps_the_uploaded_scene_and_pins_the_image(R_lts):
    """render_scene() re-uses the uploaded scene for equal CONTENT (round-1 verdict 6), re-uploads when a value changes,
    and reads the image back into scene.image page-locked — with the same pixels as a fresh Renderer's read-back."""
    import copy
    import gc
    import light_transport_b200 as lt
    from light_transport_b200 import api, host, scenes
    objs, lights, cam, depth = scenes.lts_scene(api)
    ps = lt.pack_scene(objs, lights)
    nodes = lt.compat_bvh.build_compat_nodes(ps.tri_v0, ps.tri_v1, ps.tri_v2)
    host.clear_renderer_cache()
    sc = api.Scene(camera=cam, lights=lights, width=96, height=96, max_depth=4, f_distance=depth, number_of_samples=1, allocate=False)
    kw = dict(spp=8, rng="philox", seed=3, real="f64")
    a = lt.render_scene(sc, ps, nodes, **kw).copy()
    assert a is not sc.image and len(host._renderer_cache) == 1
    r0 = next(iter(host._renderer_cache.values()))[0]
    assert sc.image.ctypes.data in host._pinned                                  # 96*96*3*8 = 221 KB >= PIN_MIN_BYTES
    # same content through a different object -> the same uploaded scene
    b = lt.render_scene(sc, copy.deepcopy(ps), copy.deepcopy(nodes), **kw)
    assert b is sc.image and np.array_equal(a, b)
    assert len(host._renderer_cache) == 1 and next(iter(host._renderer_cache.values()))[0] is r0
    # reference: an explicit Renderer writing into an ordinary (pageable) array
    img = np.empty((96, 96, 3))
    R_lts.render(width=96, height=96, spp=8, max_depth=4, camera=cam, f_distance=depth, seed=3, real="f64", image=img)
    assert np.array_equal(a, img)
    # a changed value is a different scene
    ps2 = copy.deepcopy(ps)
    ps2.mat_kd[:] = ps2.mat_kd * 0.5
    c = lt.render_scene(sc, ps2, nodes, **kw)
    assert len(host._renderer_cache) == 2 and not np.array_equal(a, c) and c.mean() < a.mean()
    # the LRU closes what it evicts
    host.RENDERER_CACHE_SIZE, keep = 1, host.RENDERER_CACHE_SIZE
    try:
        lt.render_scene(sc, ps, nodes, traversal="nearest", **kw)
        assert len(host._renderer_cache) == 1 and not r0._scene
    finally:
        host.RENDERER_CACHE_SIZE = keep
    addr = sc.image.ctypes.data
    del sc, b, c
    gc.collect()
    assert addr not in host._pinned                                              # unregistered with the array
    host.clear_renderer_cache()
    assert not host._renderer_cache


def test_recoverable_cuda_error_does_not_poison_later_calls(R_lts):
    """a refused call (registering the same host range twice) reports its error once; the next launch check must not see it"""
    import ctypes as C
    from light_transport_b200 import host
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    a = np.zeros(1 << 16)
    L, ctx = host.lib(), host._get_ctx(0)
    ptr, n = C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes)
    assert L.lt_host_register(ctx, ptr, n) == 0
    try:
        assert L.lt_host_register(ctx, ptr, n) != 0 and b"cudaHostRegister" in L.lt_last_error()
        _, s, c = R_lts.render(width=16, height=16, spp=2, max_depth=4, camera=cam, f_distance=f, seed=1, real="f32", return_samples=True)
        assert c["paths"] == 16 * 16 * 2 and np.isfinite(s).all()
    finally:
        assert L.lt_host_unregister(ctx, ptr) == 0
