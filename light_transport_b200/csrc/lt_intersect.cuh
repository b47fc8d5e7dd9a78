// lt_intersect.cuh — ray/primitive tests and the three closest-hit queries.
//   compat_query   : R/src/bvh_new.py:413-482 intersect_bvh, results identical (defects included)
//   nearest_query  : all-primitive scan, the intended semantics of R/src/utils.py:34-46
//   lbvh_query     : the same semantics through the GPU-built LBVH, stack in registers/local
// A query takes (o, d, eps, tlimit, anyhit): hits must satisfy eps < t < tlimit; with anyhit the
// search stops at the first such hit (shadow rays: "visible iff nearest t >= r - eps",
// R/src/light_samples.py:53, is exactly "no hit in (eps, r - eps)").
#pragma once
#include "lt_scene.cuh"

namespace lt {

struct Stats {
    uint32_t box_tests = 0, tri_tests = 0;
};

// R/src/intersects.py:45-104 triangle_intersect — Moeller-Trumbore, two-sided, comparisons in the
// reference's form (NaN compares false exactly where it does there).
template <typename R>
LT_DEVINL bool tri_intersect(const TriRec<R>& T, V3<R> o, V3<R> d, R& t_out) {
    const R eps = (R)0.0000001;
    V3<R> a = mk<R>(T.v0[0], T.v0[1], T.v0[2]);
    V3<R> ab = mk<R>(T.ab[0], T.ab[1], T.ab[2]);
    V3<R> ac = mk<R>(T.ac[0], T.ac[1], T.ac[2]);
    V3<R> n = mk<R>(T.n[0], T.n[1], T.n[2]);
    R ray_dot_plane = dot(d, n);
    if (rabs(ray_dot_plane) <= eps) return false;
    V3<R> pvec = cross(d, ac);
    R det = dot(ab, pvec);
    if (-eps < det && det < eps) return false;
    R inv_det = (R)1 / det;
    V3<R> tvec = o - a;
    R u = dot(tvec, pvec) * inv_det;
    if (u < 0 || u > 1) return false;
    V3<R> qvec = cross(tvec, ab);
    R v = dot(d, qvec) * inv_det;
    if (v < 0 || u + v > 1) return false;
    R t = dot(ac, qvec) * inv_det;
    if (t > eps) { t_out = t; return true; }
    return false;
}

// R/src/intersects.py:178-196 intersect_bounds on [0, inf); gamma(3) from float32 epsilon (Q10),
// evaluated in the working precision like the reference evaluates it in float64.
template <typename R>
LT_DEVINL bool slab(const R* bmin, const R* bmax, V3<R> o, V3<R> inv) {
    const R eps32 = (R)5.9604644775390625e-08;
    const R widen = (R)1 + (R)2 * (((R)3 * eps32) / ((R)1 - (R)3 * eps32));
    R t0 = 0, t1 = Lim<R>::inf();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        R oi = comp(o, i), ii = comp(inv, i);
        R t_near = (bmin[i] - oi) * ii;
        R t_far = (bmax[i] - oi) * ii;
        if (t_near > t_far) { R s = t_near; t_near = t_far; t_far = s; }
        t_far *= widen;
        t0 = t_near > t0 ? t_near : t0;
        t1 = t_far < t1 ? t_far : t1;
        if (t0 > t1) return false;
    }
    return true;
}

template <int NW> struct Bits {
    uint64_t w[NW];
    LT_DEVINL void clear() {
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = 0;
    }
    LT_DEVINL bool test(int i) const {
        if (NW == 1) return (w[0] >> i) & 1ull;
        uint64_t v = w[0];
#pragma unroll
        for (int k = 1; k < NW; ++k) v = (i >> 6) == k ? w[k] : v;
        return (v >> (i & 63)) & 1ull;
    }
    LT_DEVINL void set(int i) {
        if (NW == 1) { w[0] |= 1ull << i; return; }
#pragma unroll
        for (int k = 0; k < NW; ++k) w[k] |= (i >> 6) == k ? (1ull << (i & 63)) : 0ull;
    }
};

// ---- COMPAT --------------------------------------------------------------------------------------
// intersect_bvh keeps a `visited` list over NODES, starts at the root and (a) follows near children
// while interior bounds are hit, then (b) repeatedly jumps to the LAST unvisited index (its scan has
// no break, :445-449).  In phase (b) an interior node can only jump to an index above itself, which
// is already visited, so interior bounds tests have no effect there; what remains is: walk the
// LEAVES in descending node order, skipping those whose node bit is set, and for a leaf whose bounds
// are hit mark visited[primitive index] (:436 indexes the node list with a primitive index) and test
// its primitives with EPSILON < t < best, first found wins ties.  `leaves` is that descending list
// (field `second` = the leaf's node index).  Same hits, same ties, ~half the slab tests.
template <typename R, int NW>
LT_DEVINL void compat_leaf(const CompatNode<R>& nd, const TriRec<R>* tris, Bits<NW>& vis, V3<R> o, V3<R> d, R eps, int& prim,
                           R& best, Stats& st) {
    for (int i = 0; i < nd.n_prims; ++i) {
        int leaf_idx = nd.prim_off + i;
        vis.set(leaf_idx);
        R t;
        st.tri_tests++;
        if (tri_intersect(tris[leaf_idx], o, d, t) && eps < t && t < best) { best = t; prim = leaf_idx; }
    }
}

template <typename R, int NW>
LT_DEVINL void compat_query(const CompatNode<R>* nodes, const CompatNode<R>* leaves, int n_leaves, const TriRec<R>* tris,
                            V3<R> o, V3<R> d, R eps, R tlimit, bool anyhit, int& prim, R& best, Stats& st) {
    Bits<NW> vis;
    vis.clear();
    V3<R> inv = mk<R>((R)1 / d.x, (R)1 / d.y, (R)1 / d.z);
    prim = -1;
    best = tlimit;
    int cur = 0;
    bool done = false;
    // phase (a): the initial descent
    while (!vis.test(cur)) {
        vis.set(cur);
        const CompatNode<R>& nd = nodes[cur];
        st.box_tests++;
        if (slab(nd.bmin, nd.bmax, o, inv)) {
            if (nd.n_prims > 0) {
                compat_leaf<R, NW>(nd, tris, vis, o, d, eps, prim, best, st);
                done = (cur == 0);
                break;
            }
            bool neg = comp(inv, nd.axis) < 0;
            cur = neg ? nd.second : cur + 1;
        } else {
            done = (cur == 0);
            break;
        }
    }
    if (done) return;
    // phase (b): remaining leaves, descending node index
    for (int li = 0; li < n_leaves; ++li) {
        if (anyhit && prim >= 0) return;
        const CompatNode<R>& nd = leaves[li];
        if (vis.test(nd.second)) continue;
        vis.set(nd.second);
        st.box_tests++;
        if (slab(nd.bmin, nd.bmax, o, inv)) compat_leaf<R, NW>(nd, tris, vis, o, d, eps, prim, best, st);
    }
}

// ---- NEAREST (all primitives) ----------------------------------------------------------------------
template <typename R>
LT_DEVINL void nearest_query(const TriRec<R>* tris, int n_tris, V3<R> o, V3<R> d, R eps, R tlimit, bool anyhit, int& prim,
                             R& best, Stats& st) {
    prim = -1;
    best = tlimit;
    for (int k = 0; k < n_tris; ++k) {
        R t;
        st.tri_tests++;
        if (tri_intersect(tris[k], o, d, t) && eps < t && t < best) {
            best = t;
            prim = k;
            if (anyhit) return;
        }
    }
}

// ---- LBVH ----------------------------------------------------------------------------------------
// Child boxes are fp32 rounded outwards from the float64 triangle bounds; the slab test runs in the
// working precision R on those boxes with the reference's (1+2*gamma3) widening, entry distances are
// compared with `>` so that equal-t candidates with a lower primitive index are never culled.
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
    R t1 = fmin(fmin(f0, f1), fmin(f2, best));
    tnear = t0;
    return t0 <= t1;
}

template <typename R>
LT_DEVINL void lbvh_query(const LbvhNode* __restrict__ nodes, const TriRec<R>* __restrict__ tris, int n_tris, V3<R> o, V3<R> d,
                          R eps, R tlimit, bool anyhit, int& prim, R& best, Stats& st) {
    prim = -1;
    best = tlimit;
    V3<R> inv = mk<R>((R)1 / d.x, (R)1 / d.y, (R)1 / d.z);
    if (n_tris == 1) {
        R t;
        st.tri_tests++;
        if (tri_intersect(tris[0], o, d, t) && eps < t && t < best) { best = t; prim = 0; }
        return;
    }
    int stack[LT_LBVH_STACK];
    int sp = 0;
    int cur = 0;  // interior node index, or ~prim for a leaf
    for (;;) {
        if (cur >= 0) {
            const float4* p = reinterpret_cast<const float4*>(nodes + cur);
            float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
            int4 r = __ldg(reinterpret_cast<const int4*>(p + 3));
            R t0n, t1n;
            st.box_tests += 2;
            bool h0 = slab_t<R>((R)a.x, (R)a.y, (R)a.z, (R)a.w, (R)b.x, (R)b.y, o, inv, best, t0n);
            bool h1 = slab_t<R>((R)b.z, (R)b.w, (R)c.x, (R)c.y, (R)c.z, (R)c.w, o, inv, best, t1n);
            if (h0 && h1) {
                bool swap = t1n < t0n;
                int nearc = swap ? r.y : r.x, farc = swap ? r.x : r.y;
                stack[sp++] = farc;
                cur = nearc;
                continue;
            } else if (h0) { cur = r.x; continue; }
            else if (h1) { cur = r.y; continue; }
        } else {
            int k = ~cur;
            R t;
            st.tri_tests++;
            if (tri_intersect(tris[k], o, d, t) && eps < t && (t < best || (t == best && k < prim))) {
                best = t;
                prim = k;
                if (anyhit) return;
            }
        }
        if (sp == 0) return;
        cur = stack[--sp];
    }
}

}  // namespace lt
