#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference + numba); the fixtures it writes
are committed so that the CPU and GPU test-suites never read /root/reference.

    python tests/golden/make_golden.py            # ~3-5 min (numba JIT dominates)

What is pinned (SURVEY.md §4, §8c — the reference has no tests or golden vectors of its own):
  lts_scene.npz       packed scene arrays read off the reference's jitclass objects + the
                      node list produced by the reference's build_bvh/flatten_bvh
  mirror_scene.npz    same for the C4a mirror box
  rays_*.npz          ray sets → (prim id, t) from the reference's intersect_bvh (COMPAT) and
                      from brute-force triangle_intersect (NEAREST semantics)
  more_samples_*.npz  the "more samples" block (R/src/path_tracing.py:310-364) around the reference's trace_path + calculate_gradients
  hg_table.npz        the reference's henyey_greenstein(cosTheta, g) on a grid (R/src/medium_samples.py:14-16)
  paths_*.npz         per-(pixel, sample) radiance from the reference's trace_path with
                      replayed rand_0/rand_1 and replayed light indices
                      (semantics REFERENCE = path_tracing.trace_path,
                                 ITERATIVE = path_tracing_fix1.trace_path)
"""
import os
import sys
import time
import types

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

import numba  # noqa: E402

P = "LightTransportSimulator.light_transport.src."
import importlib  # noqa: E402

r_scene = importlib.import_module(P + "scene")
r_prim = importlib.import_module(P + "primitives")
r_mat = importlib.import_module(P + "material")
r_const = importlib.import_module(P + "constants")
r_bvh = importlib.import_module(P + "bvh_new")
r_ls = importlib.import_module(P + "light_samples")
r_rays = importlib.import_module(P + "rays")
r_vec = importlib.import_module(P + "vectors")
r_isect = importlib.import_module(P + "intersects")
r_pt = importlib.import_module(P + "path_tracing")
r_fix1 = importlib.import_module(P + "path_tracing_fix1")

from light_transport_b200 import scenes, _pack, compat_bvh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
K_LIGHT = 48  # replayed light indices per path

ref_ns = types.SimpleNamespace(
    Color=r_mat.Color, Material=r_mat.Material, PreComputedTriangle=r_prim.PreComputedTriangle,
    Light=r_scene.Light, generate_area_light_samples=r_ls.generate_area_light_samples,
    WHITE=r_const.WHITE, WHITE_2=r_const.WHITE_2, RED=r_const.RED, GREEN=r_const.GREEN,
    GLASS_MAT=r_const.GLASS_MAT)


def ref_bvh(objects):
    boxes = [r_bvh.BoundedBox(objects[i], i) for i in range(len(objects))]
    root, boxes, ordered, total = r_bvh.build_bvh(objects, boxes, 0, len(boxes), [], 0)
    lin = numba.typed.List()
    for _ in range(total):
        lin.append(r_bvh.LinearBVHNode())
    lin, _ = r_bvh.flatten_bvh(lin, root, 0)
    # Q7: ordered_prims is the input order
    assert all(ordered[i] is objects[i] or np.array_equal(ordered[i].vertex_1, objects[i].vertex_1)
               for i in range(len(objects)))
    return lin


Ray = r_rays.Ray
intersect_bvh = r_bvh.intersect_bvh
triangle_intersect = r_isect.triangle_intersect
EPS = r_const.EPSILON


@numba.njit
def _prim_index(prims, tri):
    for k in range(len(prims)):
        p = prims[k]
        if (p.vertex_1 == tri.vertex_1).all() and (p.vertex_2 == tri.vertex_2).all() and \
                (p.vertex_3 == tri.vertex_3).all():
            return k
    return -2


@numba.njit
def run_intersect(prims, bvh, o, d, out_prim, out_t, bf_prim, bf_t):
    for r in range(o.shape[0]):
        ray = Ray(o[r].copy(), d[r].copy())
        tri, t = intersect_bvh(ray, prims, bvh)
        if tri is None:
            out_prim[r] = -1
            out_t[r] = np.inf
        else:
            out_prim[r] = _prim_index(prims, tri)
            out_t[r] = t
        best = -1
        bt = np.inf
        for k in range(len(prims)):
            tt = triangle_intersect(o[r], d[r], prims[k])
            if tt is not None and EPS < tt < bt:
                bt = tt
                best = k
        bf_prim[r] = best
        bf_t[r] = bt


ref_trace = r_pt.trace_path
fix_trace = r_fix1.trace_path
normalize = r_vec.normalize


@numba.njit
def draw_light_idx(seed, n_lights, out):
    np.random.seed(seed)
    for k in range(out.shape[0]):
        out[k] = np.random.randint(0, n_lights)


@numba.njit
def check_choice(seed, n_lights, K):
    np.random.seed(seed)
    a = np.empty(K, np.int64)
    for k in range(K):
        a[k] = np.random.choice(n_lights, 1)[0]
    return a


@numba.njit
def drive_reference(scene, prims, bvh, seed_base, out):
    # restates R/src/path_tracing.py:263-288 (ray generation) around the reference's trace_path
    H, W, S = scene.height, scene.width, scene.number_of_samples
    top_bottom = np.linspace(scene.top, scene.bottom, H)
    left_right = np.linspace(scene.left, scene.right, W)
    for i in range(H):
        y = top_bottom[i]
        for j in range(W):
            x = left_right[j]
            for s in range(S):
                r0 = scene.rand_0[i, j, s]
                r1 = scene.rand_1[i, j, s]
                end = np.array([x, y, scene.f_distance, 1], dtype=np.float64)
                end[0] += r0[0] / W
                end[1] += r0[0] / H
                origin = np.array([scene.camera[0], scene.camera[1], scene.camera[2], 1], dtype=np.float64)
                direction = normalize(end - origin)
                ray = Ray(origin, direction)
                np.random.seed(seed_base + (i * W + j) * S + s)
                L, _lp, _ll = ref_trace(scene, prims, bvh, ray, 0, r0, r1, (i, j, s), False)
                out[i, j, s, :] = L


@numba.njit
def drive_fix1(scene, prims, bvh, seed_base, out):
    # restates R/src/path_tracing_fix1.py:141-162 around its trace_path
    H, W, S = scene.height, scene.width, scene.number_of_samples
    top_bottom = np.linspace(scene.top, scene.bottom, H)
    left_right = np.linspace(scene.left, scene.right, W)
    for i in range(H):
        y = top_bottom[i]
        for j in range(W):
            x = left_right[j]
            for s in range(S):
                r00 = scene.rand_0[i, j, s, 0]
                end = np.array([x, y, scene.f_distance, 1], dtype=np.float64)
                end[0] += r00 / W
                end[1] += r00 / H
                origin = np.array([scene.camera[0], scene.camera[1], scene.camera[2], 1], dtype=np.float64)
                direction = normalize(end - origin)
                ray = Ray(origin, direction)
                np.random.seed(seed_base + (i * W + j) * S + s)
                out[i, j, s, :] = fix_trace(scene, prims, bvh, ray, 0, [i, j, s])


ref_grads = r_pt.calculate_gradients


@numba.njit
def drive_records(scene, prims, bvh, seed_base, out, out_rlp, out_dl, out_il, out_minus, out_plus, out_s):
    # restates the per-sample body of render_scene_samples, R/src/path_tracing.py:270-303
    H, W, S = scene.height, scene.width, scene.number_of_samples
    top_bottom = np.linspace(scene.top, scene.bottom, H)
    left_right = np.linspace(scene.left, scene.right, W)
    for i in range(H):
        y = top_bottom[i]
        for j in range(W):
            x = left_right[j]
            for s in range(S):
                r0 = scene.rand_0[i, j, s]
                r1 = scene.rand_1[i, j, s]
                end = np.array([x, y, scene.f_distance, 1], dtype=np.float64)
                end[0] += r0[0] / W
                end[1] += r0[0] / H
                origin = np.array([scene.camera[0], scene.camera[1], scene.camera[2], 1], dtype=np.float64)
                direction = normalize(end - origin)
                ray = Ray(origin, direction)
                np.random.seed(seed_base + (i * W + j) * S + s)
                L, rlp, lists = ref_trace(scene, prims, bvh, ray, 0, r0, r1, (i, j, s), False)
                out[i, j, s, :] = L
                out_rlp[i, j, s, :] = rlp
                out_dl[i, j, s, :, :] = lists[0]
                out_il[i, j, s, :, :] = lists[1]
                m, p, ss = ref_grads(scene, prims, bvh, origin, direction, scene.rand_0_logit[i, j, s], scene.rand_1_logit[i, j, s], (i, j, s))
                out_minus[i, j, s, :, :] = m
                out_plus[i, j, s, :, :] = p
                out_s[i, j, s, :] = ss


def golden_records(name, objects, lin, lights, camera, fdist, W, H, spp, D, rand_seed, seed_base):
    sc = r_scene.Scene(camera=camera, lights=lights, width=W, height=H, max_depth=D, f_distance=fdist, number_of_samples=spp)
    np.random.seed(rand_seed)
    rand_0 = np.random.rand(H, W, spp, D)
    rand_1 = np.random.rand(H, W, spp, D)
    sc.rand_0 = rand_0.copy()
    sc.rand_1 = rand_1.copy()
    sc.rand_0_logit = np.log(rand_0 / (1 - rand_0))          # R/src/scene.py:70-71
    sc.rand_1_logit = np.log(rand_1 / (1 - rand_1))
    out = np.zeros((H, W, spp, 3)); rlp = np.zeros((H, W, spp, D))
    dl = np.zeros((H, W, spp, 3, D)); il = np.zeros((H, W, spp, 3, D))
    minus = np.zeros((H, W, spp, 2 * D, D)); plus = np.zeros((H, W, spp, 2 * D, D)); ss = np.zeros((H, W, spp, 2 * D))
    t0 = time.time()
    drive_records(sc, objects, lin, seed_base, out, rlp, dl, il, minus, plus, ss)
    nL = len(lights)
    lidx = np.zeros((H * W * spp, K_LIGHT), np.int16)
    tmp = np.zeros(K_LIGHT, np.int64)
    for p in range(H * W * spp):
        draw_light_idx(seed_base + p, nL, tmp)
        lidx[p] = tmp
    print(f"{name}: {H*W*spp} samples with records + {4*D} gradient replays each in {time.time()-t0:.1f}s; "
          f"bounce_record mean {sc.bounce_record.mean():.3f}")
    np.savez_compressed(os.path.join(OUT, name), radiance=out, record_log_pdf=rlp, direct_list=dl, indirect_list=il,
                        rlp_minus=minus, rlp_plus=plus, record_s=ss, scene_record_log_pdf=np.asarray(sc.record_log_pdf),
                        bounce_record=np.asarray(sc.bounce_record), light_idx=lidx.reshape(H, W, spp, K_LIGHT),
                        W=W, H=H, spp=spp, D=D, rand_seed=rand_seed, seed_base=seed_base, camera=camera[:3], f_distance=fdist,
                        semantics="reference", image=np.clip(out.mean(axis=2), 0, 1))


def make_rays(rng, packed, camera, fdist, n_cam=64, n_int=6000, n_light=6000):
    ys = np.linspace(1, -1, n_cam)
    xs = np.linspace(-1, 1, n_cam)
    X, Y = np.meshgrid(xs, ys)
    jit = rng.random(X.shape)
    end = np.stack([X + jit / n_cam, Y + jit / n_cam, np.full_like(X, fdist)], -1).reshape(-1, 3)
    end_nojit = np.stack([X, Y, np.full_like(X, fdist)], -1).reshape(-1, 3)
    o_cam = np.broadcast_to(camera[:3], end.shape)
    sets = [(o_cam, end - o_cam), (o_cam, end_nojit - o_cam)]
    lo = np.minimum(np.minimum(packed.tri_v0, packed.tri_v1), packed.tri_v2).min(0)
    hi = np.maximum(np.maximum(packed.tri_v0, packed.tri_v1), packed.tri_v2).max(0)
    o_int = lo + (hi - lo) * (0.02 + 0.96 * rng.random((n_int, 3)))
    d_int = rng.normal(size=(n_int, 3))
    sets.append((o_int, d_int))
    o_l = lo + (hi - lo) * (0.02 + 0.96 * rng.random((n_light, 3)))
    tgt = packed.light_pos[rng.integers(0, len(packed.light_pos), n_light)]
    sets.append((o_l, tgt - o_l))
    o = np.concatenate([s[0] for s in sets])
    d = np.concatenate([s[1] for s in sets])
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    o4 = np.concatenate([o, np.ones((len(o), 1))], 1)
    d4 = np.concatenate([d, np.zeros((len(d), 1))], 1)
    return np.ascontiguousarray(o4), np.ascontiguousarray(d4)


def golden_rays(name, objects, lin, packed, camera, fdist, seed):
    rng = np.random.default_rng(seed)
    o4, d4 = make_rays(rng, packed, camera, fdist)
    n = len(o4)
    cp, ct = np.zeros(n, np.int32), np.zeros(n)
    bp, bt = np.zeros(n, np.int32), np.zeros(n)
    t0 = time.time()
    run_intersect(objects, lin, o4, d4, cp, ct, bp, bt)
    print(f"{name}: {n} rays, compat!=nearest on {(cp != bp).sum()} ({time.time()-t0:.1f}s)")
    assert (cp != -2).all()
    np.savez_compressed(os.path.join(OUT, name), o=o4[:, :3], d=d4[:, :3], compat_prim=cp, compat_t=ct,
                        nearest_prim=bp, nearest_t=bt)


def golden_paths(name, objects, lin, lights, camera, fdist, W, H, spp, D, rand_seed, seed_base,
                 semantics):
    sc = r_scene.Scene(camera=camera, lights=lights, width=W, height=H, max_depth=D, f_distance=fdist,
                       number_of_samples=spp)
    # Scene.__init__ is jit-compiled, so its np.random.rand (R/src/scene.py:68-69) draws from numba's
    # own unseedable-from-Python generator; replace the arrays by a reproducible NumPy stream.
    np.random.seed(rand_seed)
    rand_0 = np.random.rand(H, W, spp, D)
    rand_1 = np.random.rand(H, W, spp, D)
    sc.rand_0 = rand_0.copy()
    sc.rand_1 = rand_1.copy()
    out = np.zeros((H, W, spp, 3))
    t0 = time.time()
    (drive_reference if semantics == "reference" else drive_fix1)(sc, objects, lin, seed_base, out)
    dt = time.time() - t0
    nL = len(lights)
    lidx = np.zeros((H * W * spp, K_LIGHT), np.int16)
    tmp = np.zeros(K_LIGHT, np.int64)
    for p in range(H * W * spp):
        draw_light_idx(seed_base + p, nL, tmp)
        lidx[p] = tmp
    # the tests regenerate rand_0/rand_1 from rand_seed (legacy MT19937 stream, stable across numpy)
    print(f"{name}: {H*W*spp} paths in {dt:.1f}s ({dt/(H*W*spp)*1e6:.0f} us/path incl. JIT), "
          f"mean radiance {out.mean():.4f}, max {out.max():.2f}, min {out.min():.3g}")
    np.savez_compressed(os.path.join(OUT, name), radiance=out, light_idx=lidx.reshape(H, W, spp, K_LIGHT),
                        W=W, H=H, spp=spp, D=D, rand_seed=rand_seed, seed_base=seed_base,
                        camera=camera[:3], f_distance=fdist, semantics=semantics,
                        image=np.clip(out.mean(axis=2), 0, 1))


@numba.njit
def drive_more(scene, prims, bvh, seed_base, px_h, px_w, rand_0_2, rand_1_2, samples_2, rlp_2, minus_2, plus_2, s_2):
    # restates the "more samples" block of render_scene_samples, R/src/path_tracing.py:310-364, for given pixels and random
    # numbers (the block itself draws them from numba's generator inside its loop, :328-329, and is what stops the
    # reference's top-level render_scene from compiling under numba 0.65)
    H, W = scene.height, scene.width
    top_bottom = np.linspace(scene.top, scene.bottom, H)
    left_right = np.linspace(scene.left, scene.right, W)
    n = rand_0_2.shape[1]
    for _i in range(len(px_h)):
        i = px_h[_i]
        j = px_w[_i]
        y = top_bottom[i]
        for _sample in range(n):
            rand_0_set = rand_0_2[_i, _sample, :]
            rand_1_set = rand_1_2[_i, _sample, :]
            x = left_right[j]
            end = np.array([x, y, scene.f_distance, 1], dtype=np.float64)
            end[0] += rand_0_set[0] / W
            end[1] += rand_0_set[0] / H
            origin = np.array([scene.camera[0], scene.camera[1], scene.camera[2], 1], dtype=np.float64)
            direction = normalize(end - origin)
            ray = Ray(origin, direction)
            np.random.seed(seed_base + _i * n + _sample)
            tp_result, _rlp, _ll = ref_trace(scene, prims, bvh, ray, 0, rand_0_set, rand_1_set, (i, j, _sample), False)
            m, p, ss = ref_grads(scene, prims, bvh, origin, direction, rand_0_set, rand_1_set, (i, j, _sample), False)
            samples_2[_i, _sample, :] = tp_result
            rlp_2[_i, _sample, :] = _rlp
            minus_2[_i, _sample, :, :] = m
            plus_2[_i, _sample, :, :] = p
            s_2[_i, _sample, :] = ss


def golden_more(name, objects, lin, lights, camera, fdist, W, H, n, D, pixels, rand_seed, seed_base):
    sc = r_scene.Scene(camera=camera, lights=lights, width=W, height=H, max_depth=D, f_distance=fdist, number_of_samples=n)
    P_ = len(pixels)
    np.random.seed(rand_seed)
    rand_0_2 = np.random.rand(P_, n, D)
    rand_1_2 = np.random.rand(P_, n, D)
    px_h = np.array([p[0] for p in pixels], np.int64)
    px_w = np.array([p[1] for p in pixels], np.int64)
    samples_2 = np.zeros((P_, n, 3)); rlp_2 = np.zeros((P_, n, D))
    minus_2 = np.zeros((P_, n, 2 * D, D)); plus_2 = np.zeros((P_, n, 2 * D, D)); s_2 = np.zeros((P_, n, 2 * D))
    t0 = time.time()
    drive_more(sc, objects, lin, seed_base, px_h, px_w, rand_0_2, rand_1_2, samples_2, rlp_2, minus_2, plus_2, s_2)
    nL = len(lights)
    lidx = np.zeros((P_ * n, K_LIGHT), np.int16)
    tmp = np.zeros(K_LIGHT, np.int64)
    for p in range(P_ * n):
        draw_light_idx(seed_base + p, nL, tmp)
        lidx[p] = tmp
    print(f"{name}: {P_} pixels x {n} more samples (+ {4*D} ungated replays each) in {time.time()-t0:.1f}s")
    np.savez_compressed(os.path.join(OUT, name), samples_2=samples_2, record_log_pdf_2=rlp_2, rlp_minus_2=minus_2, rlp_plus_2=plus_2,
                        record_s_2=s_2, rand_0_2=rand_0_2, rand_1_2=rand_1_2, pixels=np.array(pixels, np.int32),
                        light_idx=lidx.reshape(P_, n, K_LIGHT), W=W, H=H, n=n, D=D, camera=camera[:3], f_distance=fdist)


def golden_hg():
    """hg_table.npz: the reference's own henyey_greenstein (R/src/medium_samples.py:14-16) on a (g, cos theta) grid"""
    r_ms = importlib.import_module(P + "medium_samples")
    g = np.array([-0.9, -0.5, -0.1, 0.0, 0.1, 0.3, 0.8, 0.95])
    cos_theta = np.linspace(-1, 1, 41)
    value = np.array([[r_ms.henyey_greenstein(c, gg) for c in cos_theta] for gg in g])
    np.savez_compressed(os.path.join(OUT, "hg_table.npz"), g=g, cos_theta=cos_theta, value=value)
    print("hg_table.npz:", value.shape)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only == "hg":
        golden_hg()
        return
    if only == "more":
        objects, lights, camera, depth = scenes.lts_scene(ref_ns, new_list=numba.typed.List)
        lin = ref_bvh(objects)
        golden_more("more_samples_lts_d4.npz", objects, lin, lights, camera, depth, 24, 20, 6, 4, [(5, 7), (10, 3), (19, 23), (12, 12)], 9, 300000)
        return
    if only == "records":
        objects, lights, camera, depth = scenes.lts_scene(ref_ns, new_list=numba.typed.List)
        lin = ref_bvh(objects)
        golden_records("records_lts_d4.npz", objects, lin, lights, camera, depth, 24, 24, 2, 4, 5, 700000)
        golden_records("records_lts_d8.npz", objects, lin, lights, camera, depth, 16, 12, 2, 8, 6, 800000)
        return
    # ---- verify the light-index replay recipe (Q12) ----
    a = check_choice(123, 2000, 16)
    b = np.zeros(16, np.int64)
    draw_light_idx(123, 2000, b)
    assert np.array_equal(a, b), "np.random.choice(n,1)[0] != randint(0,n) stream"

    # ---- LTS scene (C1/C2/C5) ----
    objects, lights, camera, depth = scenes.lts_scene(ref_ns, new_list=numba.typed.List)
    lin = ref_bvh(objects)
    packed = _pack.pack_scene(objects, lights)
    nodes = compat_bvh.from_linear_bvh(lin)
    print("LTS scene:", len(objects), "prims,", len(lin), "nodes,", len(lights), "lights")
    d = packed.to_npz_dict()
    d.update(node_bmin=nodes.bmin, node_bmax=nodes.bmax, node_prim_off=nodes.prim_off,
             node_second=nodes.second, node_n_prims=nodes.n_prims, node_axis=nodes.axis,
             camera=camera[:3], f_distance=depth)
    np.savez_compressed(os.path.join(OUT, "lts_scene.npz"), **d)
    golden_rays("rays_lts.npz", objects, lin, packed, camera, depth, seed=11)
    golden_paths("paths_lts_ref_d8.npz", objects, lin, lights, camera, depth, 48, 48, 4, 8, 1, 1000, "reference")
    golden_paths("paths_lts_ref_d16.npz", objects, lin, lights, camera, depth, 32, 24, 4, 16, 2, 500000, "reference")
    golden_paths("paths_lts_iter_d8.npz", objects, lin, lights, camera, depth, 48, 48, 4, 8, 3, 900000, "iterative")
    if only in (None, "records"):
        golden_records("records_lts_d4.npz", objects, lin, lights, camera, depth, 24, 24, 2, 4, 5, 700000)
        golden_records("records_lts_d8.npz", objects, lin, lights, camera, depth, 16, 12, 2, 8, 6, 800000)

    # ---- mirror box (C4a) ----
    mobj, mlights, mcam, mf = scenes.mirror_box(ref_ns, new_list=numba.typed.List)
    mlin = ref_bvh(mobj)
    mpacked = _pack.pack_scene(mobj, mlights)
    mnodes = compat_bvh.from_linear_bvh(mlin)
    print("mirror box:", len(mobj), "prims,", len(mlin), "nodes; compat defined:", compat_bvh.compat_defined(mnodes))
    d = mpacked.to_npz_dict()
    d.update(node_bmin=mnodes.bmin, node_bmax=mnodes.bmax, node_prim_off=mnodes.prim_off,
             node_second=mnodes.second, node_n_prims=mnodes.n_prims, node_axis=mnodes.axis,
             camera=mcam[:3], f_distance=mf)
    np.savez_compressed(os.path.join(OUT, "mirror_scene.npz"), **d)
    golden_rays("rays_mirror.npz", mobj, mlin, mpacked, mcam, mf, seed=12)
    golden_paths("paths_mirror_ref_d256.npz", mobj, mlin, mlights, mcam, mf, 32, 32, 2, 256, 4, 1, "reference")
    golden_hg()
    golden_more("more_samples_lts_d4.npz", objects, lin, lights, camera, depth, 24, 20, 6, 4, [(5, 7), (10, 3), (19, 23), (12, 12)], 9, 300000)


if __name__ == "__main__":
    main()
