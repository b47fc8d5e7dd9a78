// lt_intersect.cuh — ray/primitive tests and the three closest-hit queries.
//   compat_query   : R/src/bvh_new.py:413-482 intersect_bvh, results identical (defects included)
//   nearest_query  : all-primitive scan, the intended semantics of R/src/utils.py:34-46
//   lbvh_query     : the same semantics through the GPU-built LBVH, stack in registers/local
// A query takes (o, d, eps, tlimit, anyhit): hits must satisfy eps < t < tlimit; with anyhit the
// search stops at the first such hit (shadow rays: "visible iff nearest t >= r - eps",
// R/src/light_samples.py:53, is exactly "no hit in (eps, r - eps)").  `skip` names a primitive whose
// triangle test is treated as a miss: the float32 walk passes the primitive a ray starts on, because
// the reference's 1e-6 origin offset (constants.py:12) is below float32 resolution at |x| ~ 7.5 and a
// ray cannot re-hit the planar triangle it leaves — in exact arithmetic skipping it changes nothing
// (COMPAT's visited marks are still applied).  The float64 walk passes -1.
//
// Everything here is written to keep a warp CONVERGED: the primitive tests are branch-free (one
// predicate at the end instead of the reference's early returns — same truth table, NaNs included),
// and the loops are shaped so that every lane runs the same code on its own data (ncu round 1: the
// early-return version executed with 7.7 of 32 lanes active).
#pragma once
#include "lt_scene.cuh"

namespace lt {

struct Stats {
    uint32_t box_tests = 0, tri_tests = 0;
};

// R/src/intersects.py:45-104 triangle_intersect — Moeller-Trumbore, two-sided.  The reference
// returns None at five points; `hit` is the conjunction of the five negated conditions, each written
// in the reference's form so a NaN falls through exactly where it does there.
template <typename R>
LT_DEVINL bool tri_intersect(const TriRec<R>& T, V3<R> o, V3<R> d, R& t_out) {
    const R eps = (R)0.0000001;
    V3<R> a = mk<R>(T.v0[0], T.v0[1], T.v0[2]);
    V3<R> ab = mk<R>(T.ab[0], T.ab[1], T.ab[2]);
    V3<R> ac = mk<R>(T.ac[0], T.ac[1], T.ac[2]);
    V3<R> n = mk<R>(T.n[0], T.n[1], T.n[2]);
    R ray_dot_plane = dot(d, n);
    V3<R> pvec = cross(d, ac);
    R det = dot(ab, pvec);
    R inv_det = (R)1 / det;
    V3<R> tvec = o - a;
    R u = dot(tvec, pvec) * inv_det;
    V3<R> qvec = cross(tvec, ab);
    R v = dot(d, qvec) * inv_det;
    R t = dot(ac, qvec) * inv_det;
    bool miss = (rabs(ray_dot_plane) <= eps) | ((-eps < det) & (det < eps)) | (u < 0) | (u > 1) | (v < 0) | (u + v > 1);
    t_out = t;
    return !miss & (t > eps);
}

// float32 walk on the shared-memory copy (TriPad, see stage_tris_padded): the same test from the triangle's plane and its
// dual edge vectors.  With det = d.n, t' = (a - o).n and P' = det (o - a) + t' d (= det x the hit point relative to a):
// u' = P'.n1 = det u, v' = P'.n2 = det v, so "u >= 0, v >= 0, u + v <= 1" are sign agreements with det (u <= 1 follows) and
// t = t' / det.  34 instructions instead of 55: no cross products per ray.  A NaN (degenerate triangle) fails every comparison.
#if LT_TRI_PLANES
LT_DEVINL bool tri_intersect(const TriPad& T, V3<float> o, V3<float> d, float& t_out) {
    const float eps = 0.0000001f;
    const V3<float> tv = mk<float>(o.x - T.v0[0], o.y - T.v0[1], o.z - T.v0[2]);
    const V3<float> n = mk<float>(T.gn[0], T.gn[1], T.gn[2]);
    const float det = dot(d, n);
    const float tp = -dot(tv, n);
    const V3<float> P = mk<float>(fmaf(det, tv.x, tp * d.x), fmaf(det, tv.y, tp * d.y), fmaf(det, tv.z, tp * d.z));
    const float u = dot(P, mk<float>(T.n1[0], T.n1[1], T.n1[2])), v = dot(P, mk<float>(T.n2[0], T.n2[1], T.n2[2]));
    const float w = det - u - v;
    const float t = tp / det;
    t_out = t;
    return (u * det >= 0.0f) & (v * det >= 0.0f) & (w * det >= 0.0f) & (fabsf(det) > T.thr) & (t > eps);
}
#endif

// R/src/intersects.py:178-196 intersect_bounds on [0, inf); gamma(3) from float32 epsilon (Q10),
// evaluated in the working precision like the reference evaluates it in float64.  The reference
// returns False as soon as t0 > t1; t0 only grows and t1 only shrinks (a NaN candidate never
// replaces them), so one comparison after the third axis gives the same answer.
template <typename R>
LT_DEVINL bool slab(const R* bmin, const R* bmax, V3<R> o, V3<R> inv, V3<R> = V3<R>{(R)0, (R)0, (R)0}) {
    const R eps32 = (R)5.9604644775390625e-08;
    const R widen = (R)1 + (R)2 * (((R)3 * eps32) / ((R)1 - (R)3 * eps32));
    R t0 = 0, t1 = Lim<R>::inf();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        R oi = comp(o, i), ii = comp(inv, i);
        R t_near = (bmin[i] - oi) * ii;
        R t_far = (bmax[i] - oi) * ii;
        bool sw = t_near > t_far;
        R tn = sw ? t_far : t_near, tf = sw ? t_near : t_far;
        tf *= widen;
        t0 = tn > t0 ? tn : t0;
        t1 = tf < t1 ? tf : t1;
    }
    return !(t0 > t1);
}

// float32 overload: min/max instructions instead of compare+select (6 FMNMX fewer per axis).  Differs
// from the form above only when a slab distance is NaN (origin exactly on a box plane AND a direction
// component exactly zero): fminf/fmaxf drop the NaN operand, the reference keeps its swap order.
//
// `sh` (float32 walk): the reference offsets every secondary ray's origin by EPSILON = 1e-6 (along the new direction, or
// along the surface normal: path_tracing.py:69,82,104,134,139), which float32 cannot represent at |x| ~ 7.5 — the origin
// stays ON the surface.  For the triangle tests that is handled by skipping the origin primitive; for BOUNDS it matters
// too: a wall's leaf box is flat, a ray that starts on it sees slab distances 0 <= 0 and "hits" the box, while the
// reference's ray, 1e-6 off the wall, misses it — and in COMPAT a bounds hit feeds the visited marks (kill masks), i.e.
// it changes which OTHER leaves are skipped (measured on C2: floor pixels 4 % too dark).  An origin offset delta shifts
// the slab distances of axis k by delta_k / d_k, so the walk carries sh_k = delta_k / d_k (= EPSILON on every axis for an
// offset along the ray itself) and the slab distances are formed as (b - o) * inv - sh, one FMA instead of the multiply.
LT_DEVINL bool slab(const float* bmin, const float* bmax, V3<float> o, V3<float> inv, V3<float> sh = V3<float>{0.0f, 0.0f, 0.0f}) {
    const float widen = 1.0f + 2.0f * ((3.0f * 5.9604644775390625e-08f) / (1.0f - 3.0f * 5.9604644775390625e-08f));
    float ax = fmaf(bmin[0] - o.x, inv.x, -sh.x), bx = fmaf(bmax[0] - o.x, inv.x, -sh.x);
    float ay = fmaf(bmin[1] - o.y, inv.y, -sh.y), by = fmaf(bmax[1] - o.y, inv.y, -sh.y);
    float az = fmaf(bmin[2] - o.z, inv.z, -sh.z), bz = fmaf(bmax[2] - o.z, inv.z, -sh.z);
    float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * widen;
    return !(t0 > t1);
}

// ---- float32 leaf bounds, two leaves per step ---------------------------------------------------------
// The hit mask over the COMPAT leaves is a third of the whole walk's instructions, so the float32 path
// evaluates it from a table laid out for it (built per CTA in shared memory by build_slab_pairs):
//   * per axis the NEAR plane is bmin when the direction component is positive and bmax when it is
//     negative, so the per-axis min/max of the two slab distances becomes an ADDRESS choice made once
//     per ray: each (pair, axis) holds {near_a, near_b, far_a, far_b} for both signs, 16 bytes apart.
//     Rounding is monotone, so the selected value is bit-identical to fminf/fmaxf of the two products
//     (they differ only when a distance is NaN: origin exactly on a box plane and a zero direction).
//   * the subtract and the multiply run on two leaves at a time (add.f32x2 / mul.f32x2, sm_100),
//     each half rounded exactly like the scalar instruction.
// 13.5 instructions per leaf instead of 27.5.
#define LT_SLAB_PAIR_FLOATS 24
#ifndef LT_SLAB_UNROLL
#define LT_SLAB_UNROLL 4      // pairs per trip of the leaf-bounds loop; measured on C2: 1: 20.1, 2: 20.0, 4: 20.6, 5: 19.9 Grays/s
#endif
#define LT_PRAGMA_(x) _Pragma(#x)
#define LT_PRAGMA_UNROLL(n) LT_PRAGMA_(unroll n)
LT_DEVINL int slab_pairs_bytes(int n_leaves) { return ((n_leaves + 1) >> 1) * LT_SLAB_PAIR_FLOATS * 4 + 16; }

// Built by one WARP-ALIGNED group of threads of the CTA (all of them call it).  The float trailing the pairs holds the largest
// |coordinate| of any leaf box: the float64 filter's error bound needs it (slab_pairs_filter).
template <typename R>
LT_DEVINL void build_slab_pairs(float* tab, const CompatLeaf<R>* leaves, int n_leaves) {
    const int n_pairs = (n_leaves + 1) >> 1;
    for (int i = threadIdx.x; i < n_pairs * 3; i += blockDim.x) {
        const int p = i / 3, k = i - 3 * p;
        const CompatLeaf<R>& a = leaves[2 * p];
        const CompatLeaf<R>& b = leaves[min(2 * p + 1, n_leaves - 1)];     // odd count: the last leaf twice, bit masked off
        float* t = tab + LT_SLAB_PAIR_FLOATS * p + 8 * k;
        const float an = (float)a.bmin[k], bn = (float)b.bmin[k], ax = (float)a.bmax[k], bx = (float)b.bmax[k];
        t[0] = an; t[1] = bn; t[2] = ax; t[3] = bx;     // direction component >= 0
        t[4] = ax; t[5] = bx; t[6] = an; t[7] = bn;     // direction component < 0
    }
    if (threadIdx.x == 0) {
        float m = 0.0f;
        for (int li = 0; li < n_leaves; ++li)
            for (int k = 0; k < 3; ++k) m = fmaxf(m, fmaxf(fabsf((float)leaves[li].bmin[k]), fabsf((float)leaves[li].bmax[k])));
        tab[LT_SLAB_PAIR_FLOATS * n_pairs] = m;
    }
}

// flattened leaf-order primitive sequence (position -> primitive index), 32 entries, for compat_query's small-scene path
template <typename R>
LT_DEVINL void build_flat_prims(uint8_t* tab, const CompatLeaf<R>* leaves, int n_leaves) {
    for (int li = threadIdx.x; li < n_leaves; li += blockDim.x)
        for (int i = 0; i < leaves[li].n_prims; ++i)
            if (leaves[li].flat_base + i < 32) tab[leaves[li].flat_base + i] = (uint8_t)(leaves[li].prim_off + i);
}
// the small-scene path needs positions and primitive indices that fit 32-bit masks / bytes
LT_DEVINL bool compat_small(int n_leaves, int n_cflat, int n_tris) { return n_leaves <= 32 && n_cflat <= 32 && n_tris <= 256; }

LT_DEVINL uint64_t slab_pairs_mask(const float* tab, int n_leaves, V3<float> o, V3<float> inv, V3<float> sh) {
    const float widen = 1.0f + 2.0f * ((3.0f * 5.9604644775390625e-08f) / (1.0f - 3.0f * 5.9604644775390625e-08f));
    const float4* tx = reinterpret_cast<const float4*>(tab) + (__float_as_uint(inv.x) >> 31);
    const float4* ty = reinterpret_cast<const float4*>(tab) + 2 + (__float_as_uint(inv.y) >> 31);
    const float4* tz = reinterpret_cast<const float4*>(tab) + 4 + (__float_as_uint(inv.z) >> 31);
    const float2 nox = make_float2(-o.x, -o.x), noy = make_float2(-o.y, -o.y), noz = make_float2(-o.z, -o.z);
    const float2 ivx = make_float2(inv.x, inv.x), ivy = make_float2(inv.y, inv.y), ivz = make_float2(inv.z, inv.z);
    const float2 w2 = make_float2(widen, widen);
    const float2 nsx = make_float2(-sh.x, -sh.x), nsy = make_float2(-sh.y, -sh.y), nsz = make_float2(-sh.z, -sh.z);
    const int n_pairs = (n_leaves + 1) >> 1;
    uint64_t hm = 0;
    for (int p0 = 0; p0 < n_pairs; p0 += 16) {          // 32 leaves per 32-bit word
        const int pe = min(n_pairs - p0, 16);
        uint32_t m = 0, bit = 1u;
LT_PRAGMA_UNROLL(LT_SLAB_UNROLL)
        for (int q = 0; q < pe; ++q) {
            const int p = p0 + q;
            const float4 X = tx[6 * p], Y = ty[6 * p], Z = tz[6 * p];
            const float2 nx = __ffma2_rn(__fadd2_rn(make_float2(X.x, X.y), nox), ivx, nsx);
            const float2 fx = __ffma2_rn(__fadd2_rn(make_float2(X.z, X.w), nox), ivx, nsx);
            const float2 ny = __ffma2_rn(__fadd2_rn(make_float2(Y.x, Y.y), noy), ivy, nsy);
            const float2 fy = __ffma2_rn(__fadd2_rn(make_float2(Y.z, Y.w), noy), ivy, nsy);
            const float2 nz = __ffma2_rn(__fadd2_rn(make_float2(Z.x, Z.y), noz), ivz, nsz);
            const float2 fz = __ffma2_rn(__fadd2_rn(make_float2(Z.z, Z.w), noz), ivz, nsz);
            const float t0a = fmaxf(fmaxf(nx.x, ny.x), fmaxf(nz.x, 0.0f));
            const float t0b = fmaxf(fmaxf(nx.y, ny.y), fmaxf(nz.y, 0.0f));
            const float2 t1 = __fmul2_rn(make_float2(fminf(fminf(fx.x, fy.x), fz.x), fminf(fminf(fx.y, fy.y), fz.y)), w2);
            if (!(t0a > t1.x)) m |= bit;
            if (!(t0b > t1.y)) m |= bit << 1;
            bit <<= 2;
        }
        hm |= (uint64_t)m << (2 * p0);
    }
    return n_leaves >= 64 ? hm : hm & ((1ull << n_leaves) - 1ull);
}

// ---- float64: the same table as a FILTER --------------------------------------------------------------
// The float64 walk must take exactly the reference's decisions, but most leaf boxes are hit or missed by a wide
// margin.  So the float64 query first evaluates all leaves in float32 from the table above (13.5 instructions
// per leaf instead of ~45 float64 ones) and keeps the float32 verdict only where it is PROVABLY the float64
// one: with b, o, 1/d rounded to float32 and two float32 operations per slab distance, each distance is within
//     4 * 2^-24 * (|b| + |o_k|) * |1/d_k|
// of the float64 value (three input roundings + subtract + multiply, |distance| <= (|b|+|o|)|1/d|), max/min are
// 1-Lipschitz, and the widen factor adds 2^-24 relative.  `margin` = 2^-19 * max_k (B + |o_k|) |1/d_k| with B the
// largest box coordinate is > 3x the sum of both sides' bounds; |t1 - t0| above it is a sure verdict, anything else
// (incl. every NaN / infinite case: a zero direction component makes the margin infinite) is re-tested in float64.
// Flat boxes (walls) that a ray really crosses are always ambiguous — t0 == t1 up to the widen factor — which is
// 1-2 float64 tests per ray instead of one per leaf.
LT_DEVINL float filter_margin(const float* tab, int n_leaves, V3<float> o, V3<float> inv) {
    const float B = tab[LT_SLAB_PAIR_FLOATS * ((n_leaves + 1) >> 1)];
    return 1.9073486328125e-06f * fmaxf(fmaxf((B + fabsf(o.x)) * fabsf(inv.x), (B + fabsf(o.y)) * fabsf(inv.y)), (B + fabsf(o.z)) * fabsf(inv.z));
}
LT_DEVINL void slab_pairs_filter(const float* tab, int n_leaves, V3<float> o, V3<float> inv, float margin, uint64_t& sure_hit, uint64_t& unsure) {
    const float widen = 1.0f + 2.0f * ((3.0f * 5.9604644775390625e-08f) / (1.0f - 3.0f * 5.9604644775390625e-08f));
    const int n_pairs = (n_leaves + 1) >> 1;
    const float4* tx = reinterpret_cast<const float4*>(tab) + (__float_as_uint(inv.x) >> 31);
    const float4* ty = reinterpret_cast<const float4*>(tab) + 2 + (__float_as_uint(inv.y) >> 31);
    const float4* tz = reinterpret_cast<const float4*>(tab) + 4 + (__float_as_uint(inv.z) >> 31);
    const float2 nox = make_float2(-o.x, -o.x), noy = make_float2(-o.y, -o.y), noz = make_float2(-o.z, -o.z);
    const float2 ivx = make_float2(inv.x, inv.x), ivy = make_float2(inv.y, inv.y), ivz = make_float2(inv.z, inv.z);
    const float2 w2 = make_float2(widen, widen);
    uint64_t hs = 0, us = 0;
    for (int p0 = 0; p0 < n_pairs; p0 += 16) {
        const int pe = min(n_pairs - p0, 16);
        uint32_t mh = 0, mu = 0, bit = 1u;
LT_PRAGMA_UNROLL(LT_SLAB_UNROLL)
        for (int q = 0; q < pe; ++q) {
            const int p = p0 + q;
            const float4 X = tx[6 * p], Y = ty[6 * p], Z = tz[6 * p];
            const float2 nx = __fmul2_rn(__fadd2_rn(make_float2(X.x, X.y), nox), ivx);
            const float2 fx = __fmul2_rn(__fadd2_rn(make_float2(X.z, X.w), nox), ivx);
            const float2 ny = __fmul2_rn(__fadd2_rn(make_float2(Y.x, Y.y), noy), ivy);
            const float2 fy = __fmul2_rn(__fadd2_rn(make_float2(Y.z, Y.w), noy), ivy);
            const float2 nz = __fmul2_rn(__fadd2_rn(make_float2(Z.x, Z.y), noz), ivz);
            const float2 fz = __fmul2_rn(__fadd2_rn(make_float2(Z.z, Z.w), noz), ivz);
            const float t0a = fmaxf(fmaxf(nx.x, ny.x), fmaxf(nz.x, 0.0f));
            const float t0b = fmaxf(fmaxf(nx.y, ny.y), fmaxf(nz.y, 0.0f));
            const float2 t1 = __fmul2_rn(make_float2(fminf(fminf(fx.x, fy.x), fz.x), fminf(fminf(fx.y, fy.y), fz.y)), w2);
            const float da = t1.x - t0a, db = t1.y - t0b;
            if (da > margin) mh |= bit;
            if (!(fabsf(da) > margin)) mu |= bit;
            if (db > margin) mh |= bit << 1;
            if (!(fabsf(db) > margin)) mu |= bit << 1;
            bit <<= 2;
        }
        hs |= (uint64_t)mh << (2 * p0);
        us |= (uint64_t)mu << (2 * p0);
    }
    const uint64_t all = n_leaves >= 64 ? ~0ull : ((1ull << n_leaves) - 1ull);
    sure_hit = hs & all;
    unsure = us & all;
}

// ---- float32 initial descent without a dependent chain ---------------------------------------------------
// intersect_bvh's initial descent follows the NEAR child while interior bounds are hit, and "near" is decided by
// the sign of the direction component on the node's split axis — so the chain of nodes a ray can visit is a
// function of its direction OCTANT alone.  build_chains() walks the eight chains once per CTA and lays their
// interior boxes out like the leaf pairs (near/far already selected for the octant); a ray then evaluates its
// whole chain branch-free, two nodes per step, and the descent's outcome is bit arithmetic on the hit mask:
// root missed -> nothing is visited; all interior boxes hit -> the chain's leaf is the first leaf processed (its
// own bounds test is the leaf's bit of the hit mask); otherwise the descent stopped at the first missed node.
// Replaces ~6 dependent {load node, slab test, branch} levels per ray.  Chains longer than 8 interior nodes (or a
// root that is a leaf) keep the generic loop.
#define LT_CHAIN_PAIRS 4
#define LT_NEAREST_BOXES 32   // scenes up to this many triangles get per-triangle bounds for the NEAREST scan
#define LT_CHAIN_FLOATS (8 * LT_CHAIN_PAIRS * 12)
#define LT_CHAIN_BYTES (LT_CHAIN_FLOATS * 4 + 32)
struct ChainTab {
    const float* tab = nullptr;       // [octant][pair][axis]{near_a, near_b, far_a, far_b}
    const uint16_t* meta = nullptr;   // [octant] interior nodes on the chain | leaf list position << 8; 0xffff = generic descent
};

template <typename R>
LT_DEVINL void build_chains(float* tab, uint16_t* meta, const CompatNode<R>* nodes) {
    if (threadIdx.x < 8) {
        const int oct = threadIdx.x;
        float* t = tab + oct * LT_CHAIN_PAIRS * 12;
        int cur = 0, j = 0;
        bool ok = true;
        while (nodes[cur].n_prims == 0) {
            if (j >= 2 * LT_CHAIN_PAIRS) { ok = false; break; }
            for (int s2 = j; s2 < 2 * LT_CHAIN_PAIRS; ++s2)          // this node, and padding for the slots behind it
                for (int k = 0; k < 3; ++k) {
                    const bool neg = (oct >> k) & 1;
                    float* b = t + (s2 >> 1) * 12 + k * 4 + (s2 & 1);
                    b[0] = (float)(neg ? nodes[cur].bmax[k] : nodes[cur].bmin[k]);
                    b[2] = (float)(neg ? nodes[cur].bmin[k] : nodes[cur].bmax[k]);
                }
            cur = ((oct >> nodes[cur].axis) & 1) ? nodes[cur].second : cur + 1;
            ++j;
        }
        meta[oct] = (ok && j >= 1) ? (uint16_t)(j | (nodes[cur].axis << 8)) : (uint16_t)0xffffu;
        // meta[8]: pairs any chain needs (the evaluation loop is warp-uniform)
        const unsigned need = (ok && j >= 1) ? (unsigned)((j + 1) >> 1) : 0u;
        unsigned mx = need;
        for (int d = 4; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffu, mx, d));
        if (oct == 0) meta[8] = (uint16_t)mx;
    }
}

LT_DEVINL uint32_t chain_mask(const float* tab_oct, int n_pairs, V3<float> o, V3<float> inv, V3<float> sh) {
    const float widen = 1.0f + 2.0f * ((3.0f * 5.9604644775390625e-08f) / (1.0f - 3.0f * 5.9604644775390625e-08f));
    const float4* t4 = reinterpret_cast<const float4*>(tab_oct);
    const float2 nox = make_float2(-o.x, -o.x), noy = make_float2(-o.y, -o.y), noz = make_float2(-o.z, -o.z);
    const float2 ivx = make_float2(inv.x, inv.x), ivy = make_float2(inv.y, inv.y), ivz = make_float2(inv.z, inv.z);
    const float2 w2 = make_float2(widen, widen);
    const float2 nsx = make_float2(-sh.x, -sh.x), nsy = make_float2(-sh.y, -sh.y), nsz = make_float2(-sh.z, -sh.z);
    uint32_t m = 0;
    for (int p = 0; p < n_pairs; ++p) {
        const float4 X = t4[3 * p], Y = t4[3 * p + 1], Z = t4[3 * p + 2];
        const float2 nx = __ffma2_rn(__fadd2_rn(make_float2(X.x, X.y), nox), ivx, nsx);
        const float2 fx = __ffma2_rn(__fadd2_rn(make_float2(X.z, X.w), nox), ivx, nsx);
        const float2 ny = __ffma2_rn(__fadd2_rn(make_float2(Y.x, Y.y), noy), ivy, nsy);
        const float2 fy = __ffma2_rn(__fadd2_rn(make_float2(Y.z, Y.w), noy), ivy, nsy);
        const float2 nz = __ffma2_rn(__fadd2_rn(make_float2(Z.x, Z.y), noz), ivz, nsz);
        const float2 fz = __ffma2_rn(__fadd2_rn(make_float2(Z.z, Z.w), noz), ivz, nsz);
        const float t0a = fmaxf(fmaxf(nx.x, ny.x), fmaxf(nz.x, 0.0f));
        const float t0b = fmaxf(fmaxf(nx.y, ny.y), fmaxf(nz.y, 0.0f));
        const float2 t1 = __fmul2_rn(make_float2(fminf(fminf(fx.x, fy.x), fz.x), fminf(fminf(fx.y, fy.y), fz.y)), w2);
        if (!(t0a > t1.x)) m |= 1u << (2 * p);
        if (!(t0b > t1.y)) m |= 2u << (2 * p);
    }
    return m;
}

// per-warp shared-memory scratch of compat_query's pooled triangle tests: LT_POOL_CAP items per lane and pass
#define LT_POOL_CAP 8
#define LT_POOL_SLOTS (32 * LT_POOL_CAP)
template <typename R>
struct WarpScratch {
    uint16_t* items = nullptr;    // [LT_POOL_SLOTS]  owner lane << 8 | primitive index
    R* res = nullptr;             // [LT_POOL_SLOTS]  t of a valid hit, else +inf
};
LT_DEVINL float shfl_real(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
LT_DEVINL double shfl_real(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ---- COMPAT --------------------------------------------------------------------------------------
// intersect_bvh keeps a `visited` list over NODES, starts at the root and (a) follows near children
// while interior bounds are hit, then (b) repeatedly jumps to the LAST unvisited index (its scan has
// no break, :445-449).  In phase (b) an interior node can only jump to an index above itself, which
// is already visited, so interior bounds tests have no effect there; what remains is: walk the
// LEAVES in descending node order, skipping visited ones, and for a leaf whose bounds are hit mark
// visited[primitive index] (:436 indexes the NODE list with a PRIMITIVE index) and test its
// primitives with EPSILON < t < best, first found wins ties.
//
// Only the visited state of LEAVES can change the result, so it is kept as one 64-bit mask over
// positions in `leaves` (the descending list; at most 64 leaves for the <=128 nodes COMPAT accepts):
//   * a node record of a leaf carries its list position in `axis`;
//   * a leaf record carries `kill` = the list positions of the leaves whose NODE index equals one of
//     its PRIMITIVE indices — processing the leaf ORs it into the visited mask (that is :436).
// Bounds tests are pure functions of (leaf, ray): they are all evaluated first, converged, into a hit
// mask; then every lane repeatedly takes its lowest hit-and-unvisited position.  Loops are
// warp-uniform (`amask` = the lanes running this query together) so the lanes stay in lock step.
template <typename R, typename TR = TriRec<R>>
LT_DEVINL void compat_query(unsigned amask, const CompatNode<R>* nodes, const CompatLeaf<R>* leaves, int n_leaves, const float* slabtab,
                            const uint8_t* flatprim, const TR* tris, V3<R> o, V3<R> d, R eps, R tlimit, bool anyhit, int skip, int& prim, R& best,
                            Stats& st, bool live = true, WarpScratch<R> ws = WarpScratch<R>(), ChainTab chain = ChainTab(), bool descent = true,
                            V3<R> sh = V3<R>{(R)0, (R)0, (R)0}) {
    V3<R> inv = mk<R>((R)1 / d.x, (R)1 / d.y, (R)1 / d.z);
    prim = -1;
    best = tlimit;
    // phase (a): the initial descent.  Ends at a missed node, or at a leaf (`first` = its list position
    // if its bounds are hit: it is processed before every other leaf).
    int cur = 0, first = -1;
    uint64_t lvis = 0;
    bool done = !live, descending = live && descent;     // a lane without a query of its own only helps with the pooled triangle tests;
                                                         // descent == false: NEAREST over per-triangle bounds (no tree, no kills)
    int chain_leaf = -1;
    const V3<float> of = mk<float>((float)o.x, (float)o.y, (float)o.z), invf = mk<float>((float)inv.x, (float)inv.y, (float)inv.z);
    if (sizeof(R) == 4 && chain.tab && live) {            // float32: the descent as one branch-free chain evaluation
        const unsigned oct = (__float_as_uint(invf.x) >> 31) | ((__float_as_uint(invf.y) >> 31) << 1) | ((__float_as_uint(invf.z) >> 31) << 2);
        const unsigned meta = chain.meta[oct];
        if (meta != 0xffffu) {
            const int n_int = (int)(meta & 255u);
            const uint32_t allm = (1u << n_int) - 1u;
            const uint32_t cm = chain_mask(chain.tab + oct * LT_CHAIN_PAIRS * 12, (int)chain.meta[8], of, invf, mk<float>((float)sh.x, (float)sh.y, (float)sh.z)) & allm;
            const bool all = cm == allm;
            st.box_tests += all ? n_int + 1 : __ffs((int)~cm);
            done = !(cm & 1u);
            if (all) { chain_leaf = (int)(meta >> 8); lvis |= 1ull << chain_leaf; }
            descending = false;
        }
    }
    while (__any_sync(amask, descending)) {
        if (descending) {
            const CompatNode<R>& nd = nodes[cur];
            st.box_tests++;
            bool hit = slab(nd.bmin, nd.bmax, o, inv, sh);
            bool leaf = nd.n_prims > 0;
            if (leaf) lvis |= 1ull << nd.axis;
            if (!hit || leaf) {
                done = (cur == 0);
                first = (hit && leaf) ? nd.axis : -1;
                descending = false;
            } else {
                cur = (comp(inv, nd.axis) < 0) ? nd.second : cur + 1;   // strictly increasing: never revisits
            }
        }
    }
    // phase (b) part 1: bounds of all leaves (a root that ends the walk needs none)
    uint64_t hm = 0;
    if (!done) {
        if (sizeof(R) == 4 && slabtab) {            // float32: paired, sign-selected table (see slab_pairs_mask)
            hm = slab_pairs_mask(slabtab, n_leaves, of, invf, mk<float>((float)sh.x, (float)sh.y, (float)sh.z));
        } else if (sizeof(R) == 8 && slabtab) {     // float64: float32 filter, float64 only for the leaves it cannot decide
            uint64_t unsure;
            slab_pairs_filter(slabtab, n_leaves, of, invf, filter_margin(slabtab, n_leaves, of, invf), hm, unsure);
            while (unsure) {
                const int li = __ffsll((long long)unsure) - 1;
                unsure &= unsure - 1;
                const CompatLeaf<R>& lf = leaves[li];
                hm |= (uint64_t)slab(lf.bmin, lf.bmax, o, inv, sh) << li;
            }
        } else {                                    // no table (the record kernel)
#pragma unroll 1
            for (int li = 0; li < n_leaves; ++li) {
                const CompatLeaf<R>& lf = leaves[li];
                hm |= (uint64_t)slab(lf.bmin, lf.bmax, o, inv, sh) << li;
            }
        }
        st.box_tests += n_leaves;
    }
    if (chain_leaf >= 0) first = ((hm >> chain_leaf) & 1ull) ? chain_leaf : -1;
    // phase (b) part 2, small scenes (<= 32 leaves and <= 32 referenced primitives, `flatprim` given): which
    // leaves a lane processes depends only on the hit mask and the kill masks, not on the triangle tests, so
    // the leaf walk is pure bit arithmetic (a short per-lane loop, no warp votes) that yields the set of
    // primitives to test as a mask over the flattened leaf-order primitive sequence.  The tests then run two
    // per round — the descent leaf's primitives first, then ascending positions: exactly the reference's order,
    // so "first found wins ties" is kept — with half the rounds and twice the independent arithmetic per round.
    if (flatprim) {
        uint32_t lv = (uint32_t)lvis, cur = 0, rest = 0;
        const uint32_t h = (uint32_t)hm;
        if (first >= 0) {
            const CompatLeaf<R>& lf = leaves[first];
            lv |= (1u << first) | (uint32_t)lf.kill;
            cur = (0xffffffffu >> (32 - lf.n_prims)) << lf.flat_base;
        }
        for (uint32_t cand = h & ~lv; cand; cand = h & ~lv) {
            const int li = __ffs((int)cand) - 1;
            const CompatLeaf<R>& lf = leaves[li];
            lv |= (1u << li) | (uint32_t)lf.kill;
            rest |= (0xffffffffu >> (32 - lf.n_prims)) << lf.flat_base;
        }
        if (ws.items && amask == 0xffffffffu) {
            // Pooled tests: per-lane candidate counts have a long tail (a ray through the cone's overlapping leaf
            // boxes holds 5-8 candidates while the average is 2), and a warp iterates as long as its worst lane.
            // So the whole warp — including the lanes that sit this phase out — shares the tests: every lane
            // appends its (lane, primitive) items to a per-warp list in shared memory at the offset a warp scan
            // gives it, lane j tests item base+j on the owner's ray (fetched with shuffles), and every owner
            // folds its own results in list order (strict <: first found still wins ties).
            const unsigned full = 0xffffffffu;
            const int lane = threadIdx.x & 31;
            for (;;) {
                if (anyhit && prim >= 0) { cur = 0; rest = 0; }
                const int c = min(LT_POOL_CAP, __popc(cur) + __popc(rest));
                int incl = c;
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const int v = __shfl_up_sync(full, incl, dlt);
                    if (lane >= dlt) incl += v;
                }
                const int total = __shfl_sync(full, incl, 31);
                if (total == 0) break;
                const int off = incl - c;
                for (int i = 0; i < c; ++i) {
                    if (!cur) { cur = rest; rest = 0; }
                    const int p = __ffs((int)cur) - 1;
                    cur &= cur - 1;
                    ws.items[off + i] = (uint16_t)((lane << 8) | flatprim[p]);
                }
                __syncwarp();
                for (int base = 0; base < total; base += 32) {
                    const int slot = base + lane;
                    const bool has = slot < total;
                    const int item = has ? (int)ws.items[slot] : (lane << 8);
                    const int owner = item >> 8, k = item & 255;
                    const V3<R> oo = mk<R>(shfl_real(o.x, owner), shfl_real(o.y, owner), shfl_real(o.z, owner));
                    const V3<R> dd = mk<R>(shfl_real(d.x, owner), shfl_real(d.y, owner), shfl_real(d.z, owner));
                    const int sk = __shfl_sync(full, skip, owner);
                    R t;
                    const bool hh = tri_intersect(tris[k], oo, dd, t);
                    if (has) ws.res[slot] = (hh & (k != sk) & (eps < t)) ? t : Lim<R>::inf();
                }
                st.tri_tests += c;
                __syncwarp();
                for (int i = 0; i < c; ++i) {
                    const R t = ws.res[off + i];
                    if (t < best) { best = t; prim = (int)(ws.items[off + i] & 255); }
                }
                __syncwarp();
            }
            return;
        }
        for (;;) {
            if (anyhit && prim >= 0) { cur = 0; rest = 0; }
            if (!cur) { cur = rest; rest = 0; }
            int p1 = -1, p2 = -1;
            if (cur) { p1 = __ffs((int)cur) - 1; cur &= cur - 1; }
            if (!cur) { cur = rest; rest = 0; }
            if (cur) { p2 = __ffs((int)cur) - 1; cur &= cur - 1; }
            if (!__any_sync(amask, p1 >= 0)) break;
            const int k1 = p1 >= 0 ? (int)flatprim[p1] : 0, k2 = p2 >= 0 ? (int)flatprim[p2] : 0;
            R t1, t2;
            const bool h1 = tri_intersect(tris[k1], o, d, t1), h2 = tri_intersect(tris[k2], o, d, t2);
            st.tri_tests += (p1 >= 0) + (p2 >= 0);
            if ((p1 >= 0) & h1 & (k1 != skip) & (eps < t1) & (t1 < best)) { best = t1; prim = k1; }
            if ((p2 >= 0) & h2 & (k2 != skip) & (eps < t2) & (t2 < best)) { best = t2; prim = k2; }
        }
        return;
    }
    // phase (b) part 2, general: every lane walks its own hit leaves in list (= descending node) order
    for (;;) {
        int li = first;
        first = -1;
        if (li < 0 && !(anyhit && prim >= 0)) {
            uint64_t cand = hm & ~lvis;
            if (cand) li = __ffsll((long long)cand) - 1;
        }
        const bool have = li >= 0;
        if (!__any_sync(amask, have)) break;
        int np = 0, off = 0;
        if (have) {
            const CompatLeaf<R>& lf = leaves[li];
            lvis |= (1ull << li) | lf.kill;
            np = lf.n_prims;
            off = lf.prim_off;
        }
        const int npmax = __reduce_max_sync(amask, np);
        for (int i = 0; i < npmax; ++i) {
            if (i < np) {
                int leaf_idx = off + i;
                R t;
                st.tri_tests++;
                bool h = tri_intersect(tris[leaf_idx], o, d, t);
                if (h & (leaf_idx != skip) & (eps < t) & (t < best)) { best = t; prim = leaf_idx; }
            }
        }
    }
}

// ---- NEAREST (all primitives) ----------------------------------------------------------------------
template <typename R, typename TR = TriRec<R>>
LT_DEVINL void nearest_query(const TR* tris, int n_tris, V3<R> o, V3<R> d, R eps, R tlimit, bool anyhit, int skip,
                             int& prim, R& best, Stats& st) {
    prim = -1;
    best = tlimit;
    for (int k = 0; k < n_tris; ++k) {
        R t;
        bool h = tri_intersect(tris[k], o, d, t);
        if (h & (k != skip) & (eps < t) & (t < best)) { best = t; prim = k; }
    }
    (void)anyhit;   // a shadow query only needs "any hit below tlimit"; the full scan answers that too
    st.tri_tests += n_tris;
}

// ---- LBVH ----------------------------------------------------------------------------------------
// Child boxes are fp32 rounded outwards from the float64 triangle bounds; the slab test runs in the
// working precision R on those boxes with the reference's (1+2*gamma3) widening.  A box is culled only
// when its entry distance exceeds best*(1+2*gamma3): the entry distance of a flat box and the
// Moeller-Trumbore t of a coplanar neighbour agree only to a few ulps, and candidates that TIE with
// the current best (rays along shared edges, Q13 diagonal pixels) must still be tested so that the
// lowest primitive index wins exactly as in the all-primitive scan.
template <typename R>
LT_DEVINL bool slab_t(R lx, R ly, R lz, R hx, R hy, R hz, V3<R> o, V3<R> inv, R best, R& tnear) {
    const R eps32 = (R)5.9604644775390625e-08;
    const R widen = (R)1 + (R)2 * (((R)3 * eps32) / ((R)1 - (R)3 * eps32));
    R ax = (lx - o.x) * inv.x, bx = (hx - o.x) * inv.x;
    R ay = (ly - o.y) * inv.y, by = (hy - o.y) * inv.y;
    R az = (lz - o.z) * inv.z, bz = (hz - o.z) * inv.z;
    R n0 = fmin(ax, bx), f0 = fmax(ax, bx) * widen;
    R n1 = fmin(ay, by), f1 = fmax(ay, by) * widen;
    R n2 = fmin(az, bz), f2 = fmax(az, bz) * widen;
    R t0 = fmax(fmax(n0, n1), fmax(n2, (R)0));
    R t1 = fmin(fmin(f0, f1), fmin(f2, best * widen));
    tnear = t0;
    return t0 <= t1;
}

// Both child boxes of one LbvhNode.  float32: the per-axis words hold {lo0, lo1, hi0, hi1}, so the six subtractions
// and six products of the two slab tests are three add.f32x2 + ... each half rounded like the scalar instruction;
// the widening is applied once to the minimum of the far distances (rounding is monotone: min(w*f_i) == w*min(f_i),
// bit for bit).  float64: the same arithmetic as slab_t on the re-laid node.
LT_DEVINL void slab_t2(float4 X, float4 Y, float4 Z, V3<float> o, V3<float> inv, float best, float& t0n, float& t1n, bool& h0, bool& h1) {
    const float eps32 = 5.9604644775390625e-08f;
    const float widen = 1.0f + 2.0f * ((3.0f * eps32) / (1.0f - 3.0f * eps32));
    const float2 nox = make_float2(-o.x, -o.x), noy = make_float2(-o.y, -o.y), noz = make_float2(-o.z, -o.z);
    const float2 ivx = make_float2(inv.x, inv.x), ivy = make_float2(inv.y, inv.y), ivz = make_float2(inv.z, inv.z);
    const float2 ax = __fmul2_rn(__fadd2_rn(make_float2(X.x, X.y), nox), ivx), bx = __fmul2_rn(__fadd2_rn(make_float2(X.z, X.w), nox), ivx);
    const float2 ay = __fmul2_rn(__fadd2_rn(make_float2(Y.x, Y.y), noy), ivy), by = __fmul2_rn(__fadd2_rn(make_float2(Y.z, Y.w), noy), ivy);
    const float2 az = __fmul2_rn(__fadd2_rn(make_float2(Z.x, Z.y), noz), ivz), bz = __fmul2_rn(__fadd2_rn(make_float2(Z.z, Z.w), noz), ivz);
    const float bw = best * widen;
    t0n = fmaxf(fmaxf(fminf(ax.x, bx.x), fminf(ay.x, by.x)), fmaxf(fminf(az.x, bz.x), 0.0f));
    t1n = fmaxf(fmaxf(fminf(ax.y, bx.y), fminf(ay.y, by.y)), fmaxf(fminf(az.y, bz.y), 0.0f));
    const float f0 = fminf(fminf(fminf(fmaxf(ax.x, bx.x), fmaxf(ay.x, by.x)), fmaxf(az.x, bz.x)) * widen, bw);
    const float f1 = fminf(fminf(fminf(fmaxf(ax.y, bx.y), fmaxf(ay.y, by.y)), fmaxf(az.y, bz.y)) * widen, bw);
    h0 = t0n <= f0;
    h1 = t1n <= f1;
}
LT_DEVINL void slab_t2(float4 X, float4 Y, float4 Z, V3<double> o, V3<double> inv, double best, double& t0n, double& t1n, bool& h0, bool& h1) {
    h0 = slab_t<double>((double)X.x, (double)Y.x, (double)Z.x, (double)X.z, (double)Y.z, (double)Z.z, o, inv, best, t0n);
    h1 = slab_t<double>((double)X.y, (double)Y.y, (double)Z.y, (double)X.w, (double)Y.w, (double)Z.w, o, inv, best, t1n);
}

#define LT_LBVH_DONE ((int)0x80000000)

// Warp-scheduled traversal.  Every lane holds either an interior node (cur >= 0), a leaf (cur < 0) or
// nothing (LT_LBVH_DONE).  Each iteration the WARP picks one step type — interior step or leaf step —
// by majority among the lanes that have work, and all lanes of that kind execute it together; the
// others keep their state and get their turn when they become the majority (a leaf is never lost, a
// minority only grows).  The classic per-lane while-while loop ran with 5.5 of 32 lanes active on the
// 1 M-triangle scene (ncu, round 1) because lanes drifted between the two phases.
// Near child first, far child on the stack (entry k of a lane at stack[k * stride]: shared memory [level][thread] in the walk,
// thread-local in the whole-query form; as deep as the tree, checked at build time).
// Traversal state of one lane; the stack lives in the caller's local memory so that a traversal can be
// suspended and resumed (the megakernel runs it in bounded chunks, see lt_megakernel.cuh).
template <typename R>
struct LbvhState {
    int cur, sp, prim, leaf; // interior node index >= 0, leaf ~position < 0, LT_LBVH_DONE; stack pointer; best primitive and its leaf-order position
    R best;
};

// ---- 8-wide compressed hierarchy (WNode / WTri, lt_scene.cuh) -------------------------------------------------------
// Same query semantics as the binary LBVH above (true closest hit, eps < t < tlimit, ties -> lowest primitive index), on
// the hierarchy lt_bvh_build collapses from the Karras tree.  Per-lane state is two 64-bit GROUPS:
//   G = (child_base, hits << 24 | imask)   interior children of one node still to visit
//   T = (tri_base, triangle hit bits)      triangles of one node's leaf slots still to test
// and a stack of postponed groups (8 bytes per wide level) that lives in SHARED memory in the walk and intersect kernels
// (entry k of a lane at stack[k * stride], so a warp's access is one conflict-free 256-byte row).
// A node step fetches ONE 80-byte node and decides all eight children; children are ordered front to back by the ray's
// direction octant without any sorting: slot s is visited in descending order of s ^ (7 ^ octant) because the builder
// placed each child in the slot whose diagonal it lies furthest along.
// Every iteration the warp votes between a node step and a triangle step (as the binary traversal does).
template <typename R>
struct WState {
    uint2 G, T;
    int sp;                   // stack entries in use; < 0: no query in flight
    int prim, leaf;           // best primitive so far and its position in the leaf-ordered triangle array
    R best;
};
#ifndef LT_WBVH_TRI_WEIGHT
#define LT_WBVH_TRI_WEIGHT 2  // a triangle step runs when (lanes waiting for one) * weight >= lanes waiting for a node step
#endif
#ifndef LT_WBVH_TRI_PER_STEP
#define LT_WBVH_TRI_PER_STEP 1  // triangles a lane tests per triangle step
#endif

template <typename R>
LT_DEVINL void wbvh_begin(WState<R>& s, int n_tris, R tlimit, bool live) {
    s.prim = -1; s.leaf = -1;
    s.best = tlimit;
    s.sp = live ? 0 : -1;
    s.G = make_uint2(0u, (live && n_tris > 1) ? 0x80000000u : 0u);     // the root: "slot 7 ^ octinv" of a virtual parent
    s.T = make_uint2(0u, (live && n_tris == 1) ? 1u : 0u);             // a single triangle has no node
}

LT_DEVINL uint32_t octinv_of(V3<float> inv) {
    return 7u ^ ((__float_as_uint(inv.x) >> 31) | ((__float_as_uint(inv.y) >> 31) << 1) | ((__float_as_uint(inv.z) >> 31) << 2));
}
LT_DEVINL uint32_t octinv_of(V3<double> inv) {
    return 7u ^ ((uint32_t)(__double2hiint(inv.x) >> 31 & 1) | ((uint32_t)(__double2hiint(inv.y) >> 31 & 1) << 1) |
                 ((uint32_t)(__double2hiint(inv.z) >> 31 & 1) << 2));
}

// hit word of a node from the eight per-child hit flags: interior children land on bit 24 + (slot ^ octinv), a leaf slot
// on `count` bits from its triangle offset (Ylitie et al. 2017, section 4.3)
LT_DEVINL uint32_t whit_bits(uint32_t meta4, uint32_t octinv4, uint32_t hit4 /* 0x01 per hit child byte */) {
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;          // bits 3 and 4 both set: 24..31
    const uint32_t inner_mask4 = (is_inner4 >> 4) * 7u;                       // 0x07 in the bytes of interior children
    const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if ((hit4 >> (8 * c)) & 1u) m |= ((child_bits4 >> (8 * c)) & 0xffu) << ((bit_index4 >> (8 * c)) & 0xffu);
    return m;
}

// float32 node step.  A quantised plane q of axis k turns into a ray distance with one conversion and one fused multiply-add.
// LT_WDECODE 0 (default): the conversion is a byte-permute that drops the byte into bits 8..15 of the float pattern
// 0x47000000 (= 2^15 + q exactly), so
//     t = f * A + B,   A = 2^e / d,   B = (p - 2^15 * 2^e - o) / d
// with the pattern's 2^15 folded into B.  That shifts the node's grid origin by 128 node extents, so the rounding of B costs
// |error(t)| <= ~3e-5 node extents / |d_k| (about 0.008 of a quantum) on top of the usual 2e-7 * |t|: the relative part is
// covered by the reference's (1 + 2 gamma_3) widening of the far distance, the absolute part by the 5 % of a quantum the
// builder adds to every quantised plane.  LT_WDECODE 1: I2F.U8 with a byte selector, exact, B = (p - o) / d.
// A zero direction component gives NaN distances, which fminf / fmaxf drop: the axis does not constrain the ray (conservative).
#ifndef LT_WDECODE
#define LT_WDECODE 0
#endif
LT_DEVINL float wplane(uint32_t word, int c, uint32_t magic) {
#if LT_WDECODE == 0
    // `magic` = 0x47000000 arrives as a kernel parameter: PRMT has ONE immediate slot, and with the pattern a compile-time
    // constant too ptxas spent it on the pattern and re-materialised the selector in a register before every PRMT
    return __uint_as_float(__byte_perm(word, magic, 0x7404u | ((uint32_t)c << 4)));
#else
    return (float)((word >> (8 * c)) & 0xffu);
#endif
}
LT_DEVINL void wnode_step(const WNode* __restrict__ nodes, uint32_t idx, V3<float> o, V3<float> inv, float best, uint32_t octinv, uint32_t magic,
                          uint2& G, uint2& T) {
    const uint4* p = reinterpret_cast<const uint4*>(nodes + idx);
    const uint4 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3), w4 = __ldg(p + 4);
    const float widen = 1.0f + 2.0f * ((3.0f * 5.9604644775390625e-08f) / (1.0f - 3.0f * 5.9604644775390625e-08f));
    const float sx = __uint_as_float((w0.w & 0xffu) << 23), sy = __uint_as_float(((w0.w >> 8) & 0xffu) << 23),
                sz = __uint_as_float(((w0.w >> 16) & 0xffu) << 23);
    const float Ax = sx * inv.x, Ay = sy * inv.y, Az = sz * inv.z;
    const float bias = LT_WDECODE == 0 ? 32768.0f : 0.0f;
    const float Bx = (__uint_as_float(w0.x) - sx * bias - o.x) * inv.x, By = (__uint_as_float(w0.y) - sy * bias - o.y) * inv.y,
                Bz = (__uint_as_float(w0.z) - sz * bias - o.z) * inv.z;
    const bool ngx = inv.x < 0.0f, ngy = inv.y < 0.0f, ngz = inv.z < 0.0f;
    const float bw = best * widen;
    uint32_t hitmask = 0;
    const uint32_t octinv4 = octinv * 0x01010101u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {                       // children 0..3, 4..7
        const uint32_t nx = h ? (ngx ? w2.w : w2.y) : (ngx ? w2.z : w2.x), fx = h ? (ngx ? w2.y : w2.w) : (ngx ? w2.x : w2.z);
        const uint32_t ny = h ? (ngy ? w3.w : w3.y) : (ngy ? w3.z : w3.x), fy = h ? (ngy ? w3.y : w3.w) : (ngy ? w3.x : w3.z);
        const uint32_t nz = h ? (ngz ? w4.w : w4.y) : (ngz ? w4.z : w4.x), fz = h ? (ngz ? w4.y : w4.w) : (ngz ? w4.x : w4.z);
        uint32_t hit4 = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float tnx = fmaf(wplane(nx, c, magic), Ax, Bx), tfx = fmaf(wplane(fx, c, magic), Ax, Bx);
            const float tny = fmaf(wplane(ny, c, magic), Ay, By), tfy = fmaf(wplane(fy, c, magic), Ay, By);
            const float tnz = fmaf(wplane(nz, c, magic), Az, Bz), tfz = fmaf(wplane(fz, c, magic), Az, Bz);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
            const float tf = fminf(fminf(fminf(tfx, tfy), tfz) * widen, bw);
            if (tn <= tf) hit4 |= 1u << (8 * c);
        }
        hitmask |= whit_bits(h ? w1.w : w1.z, octinv4, hit4);
    }
    G = make_uint2(w1.x, (hitmask & 0xff000000u) | (w0.w >> 24));
    T = make_uint2(w1.y, hitmask & 0x00ffffffu);
}

// float64 node step (parity path): planes decoded exactly (p + q * 2^e is exact in double), slab test as in the binary tree
LT_DEVINL void wnode_step(const WNode* __restrict__ nodes, uint32_t idx, V3<double> o, V3<double> inv, double best, uint32_t octinv, uint32_t,
                          uint2& G, uint2& T) {
    const uint4* p = reinterpret_cast<const uint4*>(nodes + idx);
    const uint4 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3), w4 = __ldg(p + 4);
    const double px = (double)__uint_as_float(w0.x), py = (double)__uint_as_float(w0.y), pz = (double)__uint_as_float(w0.z);
    const double sx = (double)__uint_as_float((w0.w & 0xffu) << 23), sy = (double)__uint_as_float(((w0.w >> 8) & 0xffu) << 23),
                 sz = (double)__uint_as_float(((w0.w >> 16) & 0xffu) << 23);
    const uint32_t qw[3][4] = {{w2.x, w2.y, w2.z, w2.w}, {w3.x, w3.y, w3.z, w3.w}, {w4.x, w4.y, w4.z, w4.w}};
    uint32_t hitmask = 0;
    const uint32_t octinv4 = octinv * 0x01010101u;
    for (int h = 0; h < 2; ++h) {
        uint32_t hit4 = 0;
        const uint32_t meta4 = h ? w1.w : w1.z;
        for (int c = 0; c < 4; ++c) {
            if (((meta4 >> (8 * c)) & 0xffu) == 0u) continue;
            const int sh = 8 * c;
            const double lx = px + (double)((qw[0][h] >> sh) & 0xffu) * sx, hx = px + (double)((qw[0][2 + h] >> sh) & 0xffu) * sx;
            const double ly = py + (double)((qw[1][h] >> sh) & 0xffu) * sy, hy = py + (double)((qw[1][2 + h] >> sh) & 0xffu) * sy;
            const double lz = pz + (double)((qw[2][h] >> sh) & 0xffu) * sz, hz = pz + (double)((qw[2][2 + h] >> sh) & 0xffu) * sz;
            double tn;
            if (slab_t<double>(lx, ly, lz, hx, hy, hz, o, inv, best, tn)) hit4 |= 1u << sh;
        }
        hitmask |= whit_bits(meta4, octinv4, hit4);
    }
    G = make_uint2(w1.x, (hitmask & 0xff000000u) | (w0.w >> 24));
    T = make_uint2(w1.y, hitmask & 0x00ffffffu);
}

// triangle of the leaf-ordered array: float32 tests the compact record, float64 the full record of the primitive
LT_DEVINL bool wtri_test(const WTri* __restrict__ wtris, const int32_t* __restrict__, const TriRec<float>* __restrict__, uint32_t ti, V3<float> o,
                         V3<float> d, float& t_out, int& prim) {
    const float4* p = reinterpret_cast<const float4*>(wtris + ti);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    const float eps = 0.0000001f;
    const V3<float> v0 = mk<float>(a.x, a.y, a.z), ab = mk<float>(b.x, b.y, b.z), ac = mk<float>(c.x, c.y, c.z);
    prim = __float_as_int(a.w);
    const V3<float> pvec = cross(d, ac);
    const float det = dot(ab, pvec);
    const float inv_det = 1.0f / det;
    const V3<float> tvec = o - v0;
    const float u = dot(tvec, pvec) * inv_det;
    const V3<float> qvec = cross(tvec, ab);
    const float v = dot(d, qvec) * inv_det;
    const float t = dot(ac, qvec) * inv_det;
    const bool miss = (fabsf(det) <= eps * b.w) | ((-eps < det) & (det < eps)) | (u < 0) | (u > 1) | (v < 0) | (u + v > 1);
    t_out = t;
    return !miss & (t > eps);
}
LT_DEVINL bool wtri_test(const WTri* __restrict__, const int32_t* __restrict__ leaf_prims, const TriRec<double>* __restrict__ tris, uint32_t ti,
                         V3<double> o, V3<double> d, double& t_out, int& prim) {
    prim = __ldg(leaf_prims + ti);
    return tri_intersect(tris[prim], o, d, t_out);
}

// Runs at most `max_steps` warp-voted steps (every lane of `amask` calls it; lanes without a query have sp < 0).
template <typename R>
LT_DEVINL void wbvh_run(unsigned amask, WState<R>& s, uint2* stack, int stride, const WNode* __restrict__ nodes, const WTri* __restrict__ wtris,
                        const int32_t* __restrict__ leaf_prims, const TriRec<R>* __restrict__ tris, V3<R> o, V3<R> d, V3<R> inv, R eps, bool anyhit,
                        int skip, Stats& st, int max_steps, uint32_t magic) {
    const uint32_t octinv = octinv_of(inv);
    for (int step = 0; step < max_steps; ++step) {
        const bool live = s.sp >= 0;
        const bool want_tri = live & (s.T.y != 0u);
        const bool want_node = live & !want_tri & ((s.G.y & 0xff000000u) != 0u);
        const unsigned nm = __ballot_sync(amask, want_node), tm = __ballot_sync(amask, want_tri);
        if (!(nm | tm)) break;
        if (__popc(tm) * LT_WBVH_TRI_WEIGHT >= __popc(nm)) {
            if (want_tri) {
#pragma unroll
                for (int rep = 0; rep < LT_WBVH_TRI_PER_STEP; ++rep) {
                    if (rep > 0 && s.T.y == 0u) break;
                    const uint32_t j = (uint32_t)__ffs((int)s.T.y) - 1u;
                    s.T.y &= s.T.y - 1u;
                    const uint32_t ti = s.T.x + j;
                    R t;
                    int k;
                    st.tri_tests++;
                    const bool h = wtri_test(wtris, leaf_prims, tris, ti, o, d, t, k);
                    if (h & (k != skip) & (eps < t) & ((t < s.best) | ((t == s.best) & (k < s.prim)))) {
                        s.best = t;
                        s.prim = k;
                        s.leaf = (int)ti;
                    }
                }
                if (anyhit && s.prim >= 0) { s.sp = -1; s.G.y = 0u; s.T.y = 0u; }
            }
        } else if (want_node) {
            const uint32_t hits = s.G.y;
            const uint32_t bit = 31u - (uint32_t)__clz((int)hits);
            const uint32_t rest = hits & ~(1u << bit);
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t idx = s.G.x + (uint32_t)__popc(hits & 0xffu & ~(0xffffffffu << slot));
            if (rest & 0xff000000u) { stack[s.sp * stride] = make_uint2(s.G.x, rest); s.sp++; }
            st.box_tests += 8;
            wnode_step(nodes, idx, o, inv, s.best, octinv, magic, s.G, s.T);
        }
        // a lane that has nothing left in hand takes the next postponed group, or is done
        if (s.sp >= 0 && s.T.y == 0u && !(s.G.y & 0xff000000u)) {
            if (s.sp > 0) {
                s.sp--;
                const uint2 e = stack[s.sp * stride];
                if (e.y & 0xff000000u) s.G = e; else s.T = e;
            } else {
                s.sp = -1;
            }
        }
    }
}

template <typename R>
LT_DEVINL void wbvh_query(unsigned amask, uint2* stack, int stride, const WNode* __restrict__ nodes, const WTri* __restrict__ wtris,
                          const int32_t* __restrict__ leaf_prims, const TriRec<R>* __restrict__ tris, int n_tris, V3<R> o, V3<R> d, R eps, R tlimit,
                          bool anyhit, int skip, int& prim, R& best, int& leaf, Stats& st, uint32_t magic) {
    WState<R> s;
    wbvh_begin<R>(s, n_tris, tlimit, true);
    V3<R> inv = mk<R>((R)1 / d.x, (R)1 / d.y, (R)1 / d.z);
    wbvh_run<R>(amask, s, stack, stride, nodes, wtris, leaf_prims, tris, o, d, inv, eps, anyhit, skip, st, 0x7fffffff, magic);
    prim = s.prim;
    best = s.best;
    leaf = s.leaf;
}

template <typename R>
LT_DEVINL void lbvh2_begin(LbvhState<R>& s, int n_tris, R tlimit, bool live) {
    s.prim = -1;
    s.leaf = -1;
    s.best = tlimit;
    s.sp = 0;
    s.cur = live ? (n_tris > 1 ? 0 : ~0) : LT_LBVH_DONE;
}

// Runs at most `max_steps` warp-voted steps (every lane of `amask` must call it, finished or idle lanes
// with cur == LT_LBVH_DONE).  Returns when all lanes are done or the budget is used up.
#ifndef LT_LDG256
#define LT_LDG256 1
#endif
// one 256-bit read-only load (32-byte aligned address)
LT_DEVINL void ldg256(const void* p, float4& a, float4& b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

template <typename R>
LT_DEVINL void lbvh2_run(unsigned amask, LbvhState<R>& s, int* stack, int stride, const LbvhNode* __restrict__ nodes,
                         const WTri* __restrict__ wtris, const int32_t* __restrict__ leaf_prims, const TriRec<R>* __restrict__ tris, V3<R> o, V3<R> d,
                         V3<R> inv, R eps, bool anyhit, int skip, Stats& st, int max_steps) {
    // (dropping stale stack entries at pop time — entry distance kept beside the node — was measured:
    //  fewer node visits but 6-9 % slower overall on both the 28- and the 1 M-triangle scene.
    //  prefetch.global.L1 of the children while the slab tests run: both children 3x slower, the chosen
    //  child 1.4x slower, the pushed far child -1.5 %: the loads are latency-bound but L1 requests are not free.
    //  L1 policies on the loads, C3 probe, base 3235 Mrays/s: triangle records L1::no_allocate 3267 (noise); nodes with few
    //  triangles below them (64 ... 16 k, a bit in the child reference) L1::no_allocate 1240 — LDG.E.NA.ENL2.256 is slow in
    //  itself —, L1::evict_first 2820-2920, L1::evict_last 2770-2830, against 2970 for the same predicated pair without a hint.
    //  The stack: ncu shows 47 % of the local-memory pops missing the L1 and pushes/pops as a third of the L1 sector traffic, yet
    //  its first 4/8/12/16 entries in shared memory ([level][thread]) give 2950/3010/2990/2985 against 3135 all-local.)
    for (int step = 0; step < max_steps; ++step) {
        const bool at_node = s.cur >= 0, at_leaf = (s.cur < 0) & (s.cur != LT_LBVH_DONE);
        const unsigned nm = __ballot_sync(amask, at_node), lm = __ballot_sync(amask, at_leaf);
        if (!(nm | lm)) break;
        if (__popc(nm) >= __popc(lm)) {
            if (at_node) {
                const float4* p = reinterpret_cast<const float4*>(nodes + s.cur);
#if LT_LDG256
                float4 X, Y, Z, rw;                      // the 64-byte node in two 256-bit loads (sm_100: LDG.E.ENL2.256) instead of four
                ldg256(p, X, Y);
                ldg256(p + 2, Z, rw);
                const int2 r = make_int2(__float_as_int(rw.x), __float_as_int(rw.y));
#else
                const float4 X = __ldg(p), Y = __ldg(p + 1), Z = __ldg(p + 2);
                const int4 r = __ldg(reinterpret_cast<const int4*>(p + 3));
#endif
                R t0n, t1n;
                bool h0, h1;
                st.box_tests += 2;
                slab_t2(X, Y, Z, o, inv, s.best, t0n, t1n, h0, h1);
                const int r0 = r.x, r1 = r.y;
                bool swap = t1n < t0n;
                int nearc = swap ? r1 : r0, farc = swap ? r0 : r1;
                if (h0 & h1) { stack[s.sp * stride] = farc; s.sp++; s.cur = nearc; }
                else if (h0 | h1) s.cur = h0 ? r0 : r1;
                else if (s.sp > 0) { s.sp--; s.cur = stack[s.sp * stride]; }
                else s.cur = LT_LBVH_DONE;
            }
        } else if (at_leaf) {
            const uint32_t lref = (uint32_t)~s.cur;        // position in the leaf-ordered triangle arrays | (triangles after this one) << 28
            const uint32_t ti = lref & 0x0fffffffu, more = lref >> 28;
            R t;
            int k;
            st.tri_tests++;
            const bool h = wtri_test(wtris, leaf_prims, tris, ti, o, d, t, k);
            if (h & (k != skip) & (eps < t) & ((t < s.best) | ((t == s.best) & (k < s.prim)))) {
                s.best = t;
                s.prim = k;
                s.leaf = (int)ti;
            }
            if (anyhit && s.prim >= 0) s.cur = LT_LBVH_DONE;
            else if (more) s.cur = ~(int)((ti + 1u) | ((more - 1u) << 28));
            else if (s.sp > 0) { s.sp--; s.cur = stack[s.sp * stride]; }
            else s.cur = LT_LBVH_DONE;
        }
    }
}

template <typename R>
LT_DEVINL void lbvh2_query(unsigned amask, const SceneView<R>& S, const TriRec<R>* __restrict__ tris, V3<R> o, V3<R> d, R eps, R tlimit, bool anyhit,
                           int skip, int& prim, R& best, Stats& st) {
    int stack[LT_LBVH_STACK];
    LbvhState<R> s;
    lbvh2_begin<R>(s, S.n_tris, tlimit, true);
    V3<R> inv = mk<R>((R)1 / d.x, (R)1 / d.y, (R)1 / d.z);
    lbvh2_run<R>(amask, s, stack, 1, S.lbvh, S.wtris, S.wleaf_prims, tris, o, d, inv, eps, anyhit, skip, st, 0x7fffffff);
    prim = s.prim;
    best = s.best;
}

// whole-query form with a thread-local stack (record / wavefront / explicit-ray kernels); `binary` picks the Karras tree
// (kept by lt_bvh_build(..., LT_BVH_KEEP_BINARY, ...)) instead of the 8-wide hierarchy
template <typename R>
LT_DEVINL void lbvh_query(unsigned amask, const SceneView<R>& S, const TriRec<R>* __restrict__ tris, V3<R> o, V3<R> d, R eps, R tlimit, bool anyhit,
                          int skip, int& prim, R& best, Stats& st, bool binary = false) {
    if (binary) {
        lbvh2_query<R>(amask, S, tris, o, d, eps, tlimit, anyhit, skip, prim, best, st);
    } else {
        uint2 stack[LT_WBVH_MAX_DEPTH];
        int leaf;
        wbvh_query<R>(amask, stack, 1, S.wnodes, S.wtris, S.wleaf_prims, tris, S.n_tris, o, d, eps, tlimit, anyhit, skip, prim, best, leaf, st, S.wmagic);
    }
}

}  // namespace lt
