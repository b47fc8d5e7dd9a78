// lt_megakernel.cuh — persistent-warp path tracer.
//
// Scheduling: the job is the linear range of work indices q = (pixel - pixel0) * ns + (s - s0).
// Warps take granules of `paths_per_warp` consecutive indices round-robin (persistent grid, a few
// CTAs per SM).  Inside a granule the lanes REGENERATE: a lane whose path has ended takes the next
// index of the granule.  The next index is warp-uniform state, handed out with ballot + popc prefix
// ranks — a warp-aggregated fetch that needs no atomic at all because the counter never leaves the
// warp.  Consecutive indices are samples of the same pixel, so lanes stay coherent on the camera
// segment and a lane's film contribution accumulates in registers; it is flushed (fp64 atomicAdd,
// warp-aggregated across lanes holding the same pixel at granule end) only when the lane's pixel
// changes.  Small scenes (COMPAT / NEAREST) are staged in shared memory once per CTA.
#pragma once
#include <type_traits>

#include "lt_walk.cuh"

namespace lt {

#define LT_BLOCK 128

// float32 triangle records re-laid with an 80-byte stride (TriPad, lt_scene.cuh)
LT_DEVINL const TriPad* stage_tris_padded(char*& cursor, const TriRec<float>* src, int n) {
    TriPad* dst = reinterpret_cast<TriPad*>(cursor);
#if LT_TRI_PLANES
    // once per CTA and triangle, in float64 (layout and meaning: TriPad in lt_scene.cuh)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const TriRec<float> S = src[i];
        const double ab[3] = {S.ab[0], S.ab[1], S.ab[2]}, ac[3] = {S.ac[0], S.ac[1], S.ac[2]};
        const double N[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
        const double nn = N[0] * N[0] + N[1] * N[1] + N[2] * N[2], len = sqrt(nn);
        const double n1[3] = {ac[1] * N[2] - ac[2] * N[1], ac[2] * N[0] - ac[0] * N[2], ac[0] * N[1] - ac[1] * N[0]};
        const double n2[3] = {N[1] * ab[2] - N[2] * ab[1], N[2] * ab[0] - N[0] * ab[2], N[0] * ab[1] - N[1] * ab[0]};
        TriPad T;
        for (int c = 0; c < 3; ++c) {
            T.v0[c] = S.v0[c]; T.n[c] = S.n[c];
            T.gn[c] = (float)(N[c] / len); T.n1[c] = (float)(n1[c] / nn); T.n2[c] = (float)(n2[c] / nn);
        }
        // R/src/intersects.py:60-66 rejects |d.n| <= eps and |det| < eps with det = -(d.n) |N|: one bound on |d.n|
        T.thr = (float)fmax(1e-7, 1e-7 / len);
        T.mat = S.mat; T.is_light = S.is_light; T.pad0 = T.pad1 = 0.0f;
        dst[i] = T;
    }
#else
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int i = threadIdx.x; i < 4 * n; i += blockDim.x) d4[5 * (i >> 2) + (i & 3)] = __ldg(s4 + i);
#endif
    cursor += (size_t)n * sizeof(TriPad);
    return dst;
}

template <typename T>
LT_DEVINL const T* stage_array(char*& cursor, const T* src, int n) {
    T* dst = reinterpret_cast<T*>(cursor);
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    int n16 = (int)((sizeof(T) * (size_t)n) / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = __ldg(s4 + i);
    cursor += ((sizeof(T) * (size_t)n + 15) / 16) * 16;
    return dst;
}

LT_DEVINL void flush_film(double* film, int32_t pix, double acc[3]) {
    if (pix >= 0 && (acc[0] != 0.0 || acc[1] != 0.0 || acc[2] != 0.0)) {
        double* f = film + (size_t)pix * 3;
        atomicAdd(f + 0, acc[0]);
        atomicAdd(f + 1, acc[1]);
        atomicAdd(f + 2, acc[2]);
    }
    acc[0] = acc[1] = acc[2] = 0.0;
}

// granule-end flush: lanes holding the same pixel are summed by the lowest such lane (shuffles),
// which then issues the atomics — one triple per distinct pixel per warp.
LT_DEVINL void flush_film_warp(double* film, int32_t pix, double acc[3]) {
    const unsigned full = 0xffffffffu;
    int lane = threadIdx.x & 31;
    unsigned same = __match_any_sync(full, pix);
    bool leader = (__ffs(same) - 1) == lane;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int src = 0; src < 32; ++src) {
        double a = __shfl_sync(full, acc[0], src), b = __shfl_sync(full, acc[1], src), c = __shfl_sync(full, acc[2], src);
        if ((same >> src) & 1u) { s0 += a; s1 += b; s2 += c; }
    }
    if (leader) { acc[0] = s0; acc[1] = s1; acc[2] = s2; flush_film(film, pix, acc); }
    acc[0] = acc[1] = acc[2] = 0.0;
}

// float32: 5 CTAs x 4 warps per SM = 96 registers, no spills.  Measured on C2 (Grays/s): 4: 14.7, 5: 15.5, 6 (80
// registers, 72 B of spills): 15.0, 7: 14.4, 8: 13.9 — the walk is bound by dependent-instruction latency and
// divergence, not by the number of resident warps.
#ifndef LT_MIN_BLOCKS_F32
#define LT_MIN_BLOCKS_F32 5
#endif
#ifndef LT_MIN_BLOCKS_F32_LBVH
#define LT_MIN_BLOCKS_F32_LBVH 5          // 8-wide hierarchy: instruction-bound (96 registers); 4 or 5 resident CTAs: 2.06 vs 2.11 Grays/s on C3
#endif
#ifndef LT_MIN_BLOCKS_F32_LBVH2
#define LT_MIN_BLOCKS_F32_LBVH2 8         // binary tree, memory-latency bound: C3 probe 5: 2.29, 6: 2.53, 7: 2.64, 8: 2.74, 10: 2.73, 12: 2.72 Grays/s
#endif
// (float64: the kernel is 110 KB of SASS and a third of its stall samples are instruction fetch (profiles/r02_walk_compat_f64_*);
//  moving normalize / slab / tri_intersect out of line (__noinline__, one copy each) made it slower, 6.6 -> 5.6 Grays/s: the
//  calls' argument traffic and generic loads cost more than the fetch misses they remove.)
#ifndef LT_MIN_BLOCKS_F64
#define LT_MIN_BLOCKS_F64 4          // C2 float64, resident CTAs 2: 4.96, 3: 6.23, 4: 6.52, 5: 6.25, 6: 6.08 Grays/s (before the float32 leaf filter)
#endif
// COMPAT: triangle tests of a warp pooled across its lanes (lt_intersect.cuh).  Measured on C2: float64 5.36 -> 6.03
// Grays/s (the test is expensive at the fp64 rate, balancing pays), float32 18.0 -> 16.9 (list + scan + two warp
// syncs cost more than the idle lanes they fill): on for float64 only.  Round 2, float32 again (C2, base 20.0): the pool for
// every test 17.9; two tests per lane in place and the pool only for the tail 18.2; skipping a round's second test when no
// lane has one (a vote) 19.85.  The tail is 3.3 rounds deep but cheaper than any scheme that balances it.
#ifndef LT_POOLED
#define LT_POOLED (sizeof(R) == 8)
#endif
#ifndef LT_EARLY_ROOT
#define LT_EARLY_ROOT 1
#endif
#ifndef LT_PHASED
#define LT_PHASED 16          // lanes waiting for a shadow query that trigger a shadow phase; 0 = no phase alignment
#endif

template <typename R, int TRAV, int SEM, int PH = 1>
__global__ void __launch_bounds__(LT_BLOCK, sizeof(R) == 4 ? (TRAV == LT_TRAV_LBVH ? LT_MIN_BLOCKS_F32_LBVH : TRAV == LT_TRAV_LBVH2 ? LT_MIN_BLOCKS_F32_LBVH2 : LT_MIN_BLOCKS_F32)
                                                           : LT_MIN_BLOCKS_F64)
walk_megakernel(const WalkParams<R> P) {
    extern __shared__ __align__(16) char smem[];
    constexpr bool HIER = TRAV == LT_TRAV_LBVH || TRAV == LT_TRAV_LBVH2;      // a GPU-built hierarchy in global memory
    constexpr bool PADDED = sizeof(R) == 4 && !HIER;                          // shared-memory triangles with the conflict-free stride
    using TRS = typename std::conditional<PADDED, TriPad, TriRec<R>>::type;
    const TRS* tris = reinterpret_cast<const TRS*>(P.S.tris);                 // (re-pointed at the staged copy below unless HIER)
    const MatRec<R>* mats = P.S.mats;
    const CompatNode<R>* cnodes = P.S.cnodes;
    const CompatLeaf<R>* cleaves = P.S.cleaves;
    const float* slabtab = nullptr;
    const uint8_t* flatprim = nullptr;
    ChainTab chain;
    if (!HIER) {
        char* cur = smem;
        if constexpr (PADDED) tris = stage_tris_padded(cur, reinterpret_cast<const TriRec<float>*>(P.S.tris), P.S.n_tris);
        else tris = reinterpret_cast<const TRS*>(stage_array(cur, P.S.tris, P.S.n_tris));
        mats = stage_array(cur, P.S.mats, P.S.n_mats);
        if (TRAV == LT_TRAV_COMPAT) {
            cnodes = stage_array(cur, P.S.cnodes, P.S.n_cnodes);
            cleaves = stage_array(cur, P.S.cleaves, P.S.n_cleaves);
        }
        if (TRAV == LT_TRAV_NEAREST && P.S.n_nleaves > 0) cleaves = stage_array(cur, P.S.nleaves, P.S.n_nleaves);
        __syncthreads();
        if (TRAV == LT_TRAV_NEAREST && P.S.n_nleaves > 0) {
            // small scene: the all-primitive scan culls with per-triangle bounds through COMPAT's machinery (no descent, no kills)
            build_slab_pairs(reinterpret_cast<float*>(cur), cleaves, P.S.n_nleaves);      // float64 uses it as a filter
            slabtab = reinterpret_cast<const float*>(cur);
            cur += (slab_pairs_bytes(P.S.n_nleaves) + 15) / 16 * 16;
            build_flat_prims<R>(reinterpret_cast<uint8_t*>(cur), cleaves, P.S.n_nleaves);
            flatprim = reinterpret_cast<const uint8_t*>(cur);
            __syncthreads();
        }
        if (TRAV == LT_TRAV_COMPAT) {
            // leaf bounds re-laid for slab_pairs_mask (float32) / slab_pairs_filter (float64)
            build_slab_pairs(reinterpret_cast<float*>(cur), cleaves, P.S.n_cleaves);
            slabtab = reinterpret_cast<const float*>(cur);
            cur += (slab_pairs_bytes(P.S.n_cleaves) + 15) / 16 * 16;
            if (compat_small(P.S.n_cleaves, P.S.n_cflat, P.S.n_tris)) {
                build_flat_prims<R>(reinterpret_cast<uint8_t*>(cur), cleaves, P.S.n_cleaves);
                flatprim = reinterpret_cast<const uint8_t*>(cur);
            }
            cur += 32;
            if (sizeof(R) == 4) {                           // float32: per-octant descent chains (as a float64 filter: measured, no gain)
                build_chains(reinterpret_cast<float*>(cur), reinterpret_cast<uint16_t*>(cur + LT_CHAIN_FLOATS * 4), cnodes);
                chain.tab = reinterpret_cast<const float*>(cur);
                chain.meta = reinterpret_cast<const uint16_t*>(cur + LT_CHAIN_FLOATS * 4);
            }
            __syncthreads();
        }
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gwarp = gtid >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t ppw = P.paths_per_warp;
    const R eps = (R)P.eps;

    Frame<R> frames[SEM == LT_SEM_REFERENCE ? LT_LOCAL_FRAMES : 1];
    Lane<R> ln;
    ln.qtype = 0;
    // LBVH: the traversal of a lane is RESUMABLE and runs in chunks of P.lbvh_chunk voted steps per loop
    // iteration, so a lane whose ray is done shades and starts its next query while its neighbours keep
    // descending — nobody waits for the longest traversal of the warp (1 M-triangle scene: traversal lengths
    // spread 10..80 steps; whole-query iterations kept ~10 of 32 lanes busy).  Measured on that scene:
    // whole query 1.76, chunk 8: 2.21, 16: 2.48, 24: 2.51, 64: 2.18 Grays/s; letting finished lanes pile up
    // before shading (4..16 lanes) changed nothing (+-1 %).
    // 8-wide hierarchy: stack of postponed groups in dynamic shared memory, [wide level][thread], 8 bytes each
    uint2* const wstack = reinterpret_cast<uint2*>(smem) + threadIdx.x;
    WState<R> ws8;
    ws8.G = make_uint2(0u, 0u); ws8.T = make_uint2(0u, 0u); ws8.sp = -1; ws8.prim = -1; ws8.leaf = -1; ws8.best = (R)0;
    // binary tree: stack in local memory (measured on C3: 3.35 Grays/s; the same stack in shared memory, [level][thread]: 3.28)
    int tstack[TRAV == LT_TRAV_LBVH2 ? LT_LBVH_STACK : 1];
    LbvhState<R> ts;
    ts.cur = LT_LBVH_DONE; ts.sp = 0; ts.prim = -1; ts.leaf = -1; ts.best = (R)0;
    V3<R> tinv = mk<R>((R)0, (R)0, (R)0);
    bool tracing = false;
    Stats st;
    // Counters live in SHARED memory, one row per warp, bumped by lane 0 with result-less atomics (fire and
    // forget) from warp-uniform values (ballot/popc, warp reductions at the converged ends of an iteration).
    // As eleven per-lane registers they did not fit the 80-register budget, were spilled, and every `n++` became
    // a local-memory load + store whose latency was 12 % of all stall samples.
    enum { C_EXT = 0, C_SHADOW, C_STEPS, C_ITERS, C_MEMO, C_LIDX, C_FOVF, C_MAXF, C_PATHS, C_N };
    __shared__ unsigned int sh_cnt[LT_BLOCK / 32][C_N];
    unsigned int* const cnt = sh_cnt[threadIdx.x >> 5];
    if (lane < C_N) cnt[lane] = 0;
    __syncwarp();
    double acc[3] = {0, 0, 0};
    int32_t acc_pix = -1;
    WarpScratch<R> ws;
    constexpr bool pooled = TRAV == LT_TRAV_COMPAT && LT_POOLED;
    __shared__ uint16_t pool_items[pooled ? LT_BLOCK / 32 : 1][pooled ? LT_POOL_SLOTS : 1];
    __shared__ R pool_res[pooled ? LT_BLOCK / 32 : 1][pooled ? LT_POOL_SLOTS : 1];
    if (pooled && flatprim) { ws.items = pool_items[threadIdx.x >> 5]; ws.res = pool_res[threadIdx.x >> 5]; }

    // The warp's granules [g*ppw, g*ppw + g_n), g = gwarp, gwarp + nwarps, ...: (pixel, sample) of a granule's first index
    // from one 64-bit division, 32-bit ones per path.  Lanes move on to the next granule as soon as the current one
    // has no index left — the warp drains only once, at the end of its whole list (draining per granule left 43 % of
    // the lane slots of a small job like C1 idle: two paths per lane and granule).
    int64_t g = gwarp;
    int32_t g_n = 0, g_pix = 0, next = 0;
    uint32_t g_s = 0;
    auto load_granule = [&]() {
        const int64_t g_first = g * ppw;
        g_n = g_first < P.total_paths ? (int32_t)min((int64_t)ppw, P.total_paths - g_first) : 0;
        g_pix = (int32_t)(g_first / P.ns);
        g_s = (uint32_t)(g_first - (int64_t)g_pix * P.ns);
        next = 0;
        if (lane == 0 && g_n > 0) atomicAdd(cnt + C_PATHS, (unsigned)g_n);   // every started path is tallied exactly once
    };
    load_granule();
    {
        for (;;) {
            // ---- regeneration: idle lanes take the next indices of the granule (not ahead of a shadow phase, which
            // they could not join: they are refilled together before the next extension phase)
            const bool phased = PH && !HIER && LT_PHASED > 0;
            const bool shadow_phase = phased && __popc(__ballot_sync(full, ln.qtype == 2)) >= LT_PHASED;
            bool need = ln.qtype == 0;
            unsigned m = shadow_phase ? 0u : __ballot_sync(full, need);
            if (m) {
                if (next >= g_n && g_n > 0) { g += nwarps; load_granule(); }        // warp-uniform
                int32_t mine = next + __popc(m & ((1u << lane) - 1u));
                next += __popc(m);
                if (need && mine < g_n) {
                    uint32_t off = (uint32_t)mine + g_s;
                    uint32_t dp = off / (uint32_t)P.ns;
                    int32_t pix = P.pixel0 + g_pix + (int32_t)dp;
                    int s_rel = (int)(off - dp * (uint32_t)P.ns);
                    if (pix != acc_pix) { flush_film(P.film, acc_pix, acc); acc_pix = pix; }
                    start_path<R>(P, ln, pix, s_rel);
                    if (ln.qtype == 0) {   // max_depth == 0
                        if (P.samples) { double* o = P.samples + ((size_t)pix * P.ns + s_rel) * 3; o[0] = o[1] = o[2] = 0; }
                    }
                }
            }
            bool active = ln.qtype != 0;
            unsigned am = __ballot_sync(full, active);
            if (!am) {
                if (g_n > 0) continue;       // nobody holds a ray, but the list is not exhausted: the next pass refills
                break;
            }
            int want = ln.qtype;
            if (phased) {
                // Phase alignment: an iteration serves EITHER extension OR shadow queries, so the lanes that shade
                // afterwards run one branch family together (the traversal is the cheap, converged part of an
                // iteration; the shading branches are the expensive, divergent one).  Lanes holding the other kind of
                // query sit the iteration out.  A shadow phase runs once LT_PHASED lanes wait for one (or nobody has
                // an extension ray).  Measured on C2: off 15.5, strict alternation 16.2, threshold 8/12/16/20/24/28:
                // 16.2/16.3/16.4/16.4/16.2/16.2 Grays/s, with 83 % instead of 99 % of the traversal lanes busy.
                // It pays where most extension hits spawn a shadow query; in the mirror box of C4a (one diffuse wall
                // in six) the waiting costs more than the convergence gives: 10.9 aligned vs 12.5 Grays/s not aligned.
                // The host picks the variant (PH) from the scene's diffuse share of surface area.
                want = (shadow_phase || !__any_sync(full, ln.qtype == 1)) ? 2 : 1;
                active = ln.qtype == want;
                am = __ballot_sync(full, active);
            }
            uint32_t it_memo = 0, it_lidx = 0, it_fovf = 0, it_maxf = 0, it_early = 0;    // this iteration's per-lane events, reduced below
            if constexpr (TRAV == LT_TRAV_LBVH) {
                if (active && !tracing) {
                    wbvh_begin<R>(ws8, P.S.n_tris, ln.tlimit, true);
                    tinv = mk<R>((R)1 / ln.d.x, (R)1 / ln.d.y, (R)1 / ln.d.z);
                    tracing = true;
                }
                wbvh_run<R>(full, ws8, wstack, LT_BLOCK, P.S.wnodes, P.S.wtris, P.S.wleaf_prims, tris, ln.o, ln.d, tinv, eps, ln.qtype == 2, ln.skip, st,
                            P.lbvh_chunk, P.S.wmagic);
                active = tracing && ws8.sp < 0;                   // only lanes whose query just completed go on to shade
                if (active) { tracing = false; ts.prim = ws8.prim; ts.best = ws8.best; ts.leaf = ws8.leaf; }
            }
            if constexpr (TRAV == LT_TRAV_LBVH2) {
                if (active && !tracing) {
                    lbvh2_begin<R>(ts, P.S.n_tris, ln.tlimit, true);
                    tinv = mk<R>((R)1 / ln.d.x, (R)1 / ln.d.y, (R)1 / ln.d.z);
                    tracing = true;
                }
                lbvh2_run<R>(full, ts, tstack, 1, P.S.lbvh, P.S.wtris, P.S.wleaf_prims, tris, ln.o, ln.d, tinv, eps, ln.qtype == 2, ln.skip, st,
                             P.lbvh_chunk);
                active = tracing && ts.cur == LT_LBVH_DONE;       // only lanes whose query just completed go on to shade
                if (active) tracing = false;
            }
            {
                const unsigned em = __ballot_sync(full, active && ln.qtype == 1), sm = __ballot_sync(full, active && ln.qtype == 2);
                if (lane == 0) {
                    atomicAdd(cnt + C_ITERS, 1u);
                    atomicAdd(cnt + C_STEPS, (unsigned)__popc(am));
                    if (em) atomicAdd(cnt + C_EXT, (unsigned)__popc(em));
                    if (sm) atomicAdd(cnt + C_SHADOW, (unsigned)__popc(sm));
                }
            }
            // ---- the one ray query of this iteration
            int prim = ts.prim;
            R t = ts.best;
            if (TRAV == LT_TRAV_COMPAT) {
                if (LT_POOLED) {
                    // the whole warp enters: lanes without a query of their own help with the pooled triangle tests
                    compat_query<R, TRS>(full, cnodes, cleaves, P.S.n_cleaves, slabtab, flatprim, tris, ln.o, ln.d, eps, ln.tlimit, ln.qtype == 2, ln.skip,
                                    prim, t, st, active, ws, chain, true, ln.qshift);
                } else if (active) {
                    compat_query<R, TRS>(am, cnodes, cleaves, P.S.n_cleaves, slabtab, flatprim, tris, ln.o, ln.d, eps, ln.tlimit, ln.qtype == 2, ln.skip, prim, t, st,
                                    true, ws, chain, true, ln.qshift);
                }
            }
            if (active) {
                const bool anyhit = ln.qtype == 2;
                if (TRAV == LT_TRAV_NEAREST) {
                    if (flatprim)
                        compat_query<R, TRS>(am, nullptr, cleaves, P.S.n_nleaves, slabtab, flatprim, tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st,
                                        true, WarpScratch<R>(), ChainTab(), false);
                    else
                        nearest_query<R, TRS>(tris, P.S.n_tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st);
                }
                (void)anyhit;
                int stt;
                R adv_rand0 = (R)0;
                if (want == 1) {            // warp-uniform when phases are aligned
                    HitRec<R> hit_rec;
                    if constexpr (HIER && sizeof(R) == 4) {
                        // the float32 hierarchy walks shade from the compact leaf-order record they have just tested: the
                        // normal is rebuilt from ab x ac (R/src/primitives.py:108-111), material and is_light ride in c.w
                        if (prim >= 0) {
                            const float4* wp = reinterpret_cast<const float4*>(P.S.wtris + ts.leaf);
                            const float4 wb = __ldg(wp + 1), wc = __ldg(wp + 2);
                            hit_rec.n = normalize_fast(cross(mk<float>(wb.x, wb.y, wb.z), mk<float>(wc.x, wc.y, wc.z)));
                            const int ml = __float_as_int(wc.w);
                            hit_rec.mat = ml & 0x7fffffff; hit_rec.is_light = (ml >> 31) & 1; hit_rec.valid = true;
                        }
                    }
                    stt = shade_extension<R, SEM, TRS>(P, ln, prim, t, tris, mats, it_lidx, adv_rand0, hit_rec);
                } else {
                    stt = shade_shadow<R, SEM, LT_LOCAL_FRAMES>(P, ln, frames, gtid, prim >= 0, it_fovf, it_maxf, adv_rand0);
                }
                bool alive = true;
                if (stt != ST_QUERY) {      // roulette / bounce++ / frame pops: one shared copy for every branch
                    alive = advance<R, SEM, LT_LOCAL_FRAMES>(P, ln, frames, gtid, stt == ST_ADVANCE_ENDED, adv_rand0, it_memo);
                    ln.qtype = 1;
                    ln.tlimit = Lim<R>::inf();
                }
                // COMPAT: a new extension ray that misses the ROOT bounds is a miss (the descent ends at once and no leaf is
                // visited), and a miss ends the whole path (see advance()).  A quarter of the walk's rays on the
                // Cornell scene are like that (Q3 sends them through a wall).  Decided here, the lane is refilled
                // ahead of the next extension phase instead of idling through one; the ray is counted as traced.
                // NEAREST: the same with the bounds of all vertices — a ray that misses them (widened like every slab
                // test, so conservatively) cannot hit a triangle.
                if ((TRAV == LT_TRAV_COMPAT || TRAV == LT_TRAV_NEAREST) && LT_EARLY_ROOT && alive && ln.qtype == 1) {
                    const V3<R> inv = mk<R>((R)1 / ln.d.x, (R)1 / ln.d.y, (R)1 / ln.d.z);
                    R bmin[3], bmax[3];
                    for (int c = 0; c < 3; ++c) {
                        bmin[c] = TRAV == LT_TRAV_COMPAT ? cnodes[0].bmin[c] : (R)P.S.bbox_min[c];
                        bmax[c] = TRAV == LT_TRAV_COMPAT ? cnodes[0].bmax[c] : (R)P.S.bbox_max[c];
                    }
                    if (!slab(bmin, bmax, ln.o, inv, TRAV == LT_TRAV_COMPAT ? ln.qshift : mk<R>((R)0, (R)0, (R)0))) {
                        if (TRAV == LT_TRAV_COMPAT) st.box_tests++;
                        it_early++;
                        ln.memo |= 1;
                        alive = advance<R, SEM, LT_LOCAL_FRAMES>(P, ln, frames, gtid, true, adv_rand0, it_memo);
                    }
                }
                if (!alive) {
                    // ---- tally (R/src/path_tracing.py:297-298)
                    acc[0] += (double)ln.L.x; acc[1] += (double)ln.L.y; acc[2] += (double)ln.L.z;
                    if (P.samples) {
                        double* o = P.samples + ((size_t)ln.pix * P.ns + ln.s_rel) * 3;
                        o[0] = (double)ln.L.x; o[1] = (double)ln.L.y; o[2] = (double)ln.L.z;
                    }
                    ln.qtype = 0;
                }
            }
            // ---- converged again: fold this iteration's per-lane events into the warp's counters
            if (SEM == LT_SEM_REFERENCE) {
                const unsigned memo = __reduce_add_sync(full, it_memo), maxf = __reduce_max_sync(full, it_maxf);
                if (lane == 0) {
                    if (memo) atomicAdd(cnt + C_MEMO, memo);
                    if (maxf) atomicMax(cnt + C_MAXF, maxf);
                }
            }
            if ((TRAV == LT_TRAV_COMPAT || TRAV == LT_TRAV_NEAREST) && LT_EARLY_ROOT) {
                const unsigned early = __reduce_add_sync(full, it_early);
                if (lane == 0 && early) atomicAdd(cnt + C_EXT, early);
            }
            if (__any_sync(full, (it_lidx | it_fovf) != 0)) {
                const unsigned a = __reduce_add_sync(full, it_lidx), b = __reduce_add_sync(full, it_fovf);
                if (lane == 0) { atomicAdd(cnt + C_LIDX, a); atomicAdd(cnt + C_FOVF, b); }
            }
        }
        flush_film_warp(P.film, acc_pix, acc);
        acc_pix = -1;
    }
    // ---- counters: one global atomic per warp and field (box/triangle tests are the only per-lane ones left)
    if (P.counters) {
        unsigned long long bt = st.box_tests, tt = st.tri_tests;
        for (int o = 16; o > 0; o >>= 1) { bt += __shfl_xor_sync(full, bt, o); tt += __shfl_xor_sync(full, tt, o); }
        __syncwarp();
        if (lane == 0) {
            const unsigned long long memo = cnt[C_MEMO];
            const unsigned long long v[13] = {cnt[C_PATHS], cnt[C_EXT] + memo, cnt[C_SHADOW], bt, tt, 0, cnt[C_LIDX], cnt[C_FOVF],
                                              cnt[C_STEPS], (unsigned long long)cnt[C_ITERS] * 32ull, 0, 0, memo};
#pragma unroll
            for (int k = 0; k < 13; ++k)
                if (k != 5 && k != 10 && k != 11 && v[k]) atomicAdd(P.counters + k, v[k]);
            atomicMax(P.counters + 5, SEM == LT_SEM_REFERENCE ? (unsigned long long)max(cnt[C_MAXF], 1u) : 1ull);
        }
    }
}

// ---- explicit-ray closest-hit kernel behind lt_intersect -------------------------------------------
template <typename R>
struct IntersectParams {
    SceneView<R> S;
    const double* o;
    const double* d;
    int64_t n;
    int32_t* prim;
    double* t;
    unsigned long long* counters;
    double eps;
    int32_t traversal;
};

template <typename R, int TRAV>
__global__ void __launch_bounds__(LT_BLOCK) intersect_kernel(const IntersectParams<R> P) {
    Stats st;
    extern __shared__ __align__(16) char smem[];
    const float* slabtab = nullptr;
    const uint8_t* flatprim = nullptr;
    ChainTab chain;
    if (TRAV == LT_TRAV_COMPAT) {                       // dynamic shared memory = 32 + slab_pairs_bytes(n_cleaves)
        if (compat_small(P.S.n_cleaves, P.S.n_cflat, P.S.n_tris)) {
            build_flat_prims<R>(reinterpret_cast<uint8_t*>(smem), P.S.cleaves, P.S.n_cleaves);
            flatprim = reinterpret_cast<const uint8_t*>(smem);
        }
        build_slab_pairs(reinterpret_cast<float*>(smem + 32), P.S.cleaves, P.S.n_cleaves);
        slabtab = reinterpret_cast<const float*>(smem + 32);
        if (sizeof(R) == 4) {
            char* c2 = smem + 32 + (slab_pairs_bytes(P.S.n_cleaves) + 15) / 16 * 16;
            build_chains(reinterpret_cast<float*>(c2), reinterpret_cast<uint16_t*>(c2 + LT_CHAIN_FLOATS * 4), P.S.cnodes);
            chain.tab = reinterpret_cast<const float*>(c2);
            chain.meta = reinterpret_cast<const uint16_t*>(c2 + LT_CHAIN_FLOATS * 4);
        }
        __syncthreads();
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // warp-uniform loop bound so that the ballot below is executed by whole warps
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < P.n; base += stride) {
        const int64_t r = base + lane;
        const bool valid = r < P.n;
        const unsigned am = __ballot_sync(full, valid);
        if (valid) {
            V3<R> o = mk<R>((R)P.o[3 * r], (R)P.o[3 * r + 1], (R)P.o[3 * r + 2]);
            V3<R> d = mk<R>((R)P.d[3 * r], (R)P.d[3 * r + 1], (R)P.d[3 * r + 2]);
            int prim;
            R t;
            if (TRAV == LT_TRAV_COMPAT)
                compat_query<R>(am, P.S.cnodes, P.S.cleaves, P.S.n_cleaves, slabtab, flatprim, P.S.tris, o, d, (R)P.eps, Lim<R>::inf(), false, -1, prim, t, st, true,
                                WarpScratch<R>(), chain);
            else if (TRAV == LT_TRAV_NEAREST)
                nearest_query<R>(P.S.tris, P.S.n_tris, o, d, (R)P.eps, Lim<R>::inf(), false, -1, prim, t, st);
            else
                lbvh_query<R>(am, P.S, P.S.tris, o, d, (R)P.eps, Lim<R>::inf(), false, -1, prim, t, st, TRAV == LT_TRAV_LBVH2);
            P.prim[r] = prim;
            P.t[r] = prim >= 0 ? (double)t : (double)Lim<double>::inf();
        }
    }
    if (P.counters) {
        unsigned long long b = st.box_tests, tt = st.tri_tests;
        for (int o = 16; o > 0; o >>= 1) { b += __shfl_xor_sync(full, b, o); tt += __shfl_xor_sync(full, tt, o); }
        if (lane == 0) { atomicAdd(P.counters + 3, b); atomicAdd(P.counters + 4, tt); }
    }
}

// host-side dispatch, instantiated per R in lt_walk_f32.cu / lt_walk_f64.cu
template <typename R> int launch_walk(const WalkParams<R>& P, int grid, size_t smem, cudaStream_t stream);
template <typename R> int launch_intersect(const IntersectParams<R>& P, int grid, cudaStream_t stream);
template <typename R> int walk_kernel_attrs(int trav, int sem, int phased, int* regs, int* local_bytes, int* max_blocks_per_sm, size_t smem);

}  // namespace lt
