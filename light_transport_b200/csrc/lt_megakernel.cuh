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
#include "lt_walk.cuh"

namespace lt {

#define LT_BLOCK 128

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

LT_DEVINL void flush_film(double* film, int64_t pix, double acc[3]) {
    if (pix >= 0 && (acc[0] != 0.0 || acc[1] != 0.0 || acc[2] != 0.0)) {
        atomicAdd(film + pix * 3 + 0, acc[0]);
        atomicAdd(film + pix * 3 + 1, acc[1]);
        atomicAdd(film + pix * 3 + 2, acc[2]);
    }
    acc[0] = acc[1] = acc[2] = 0.0;
}

// granule-end flush: lanes holding the same pixel are summed by the lowest such lane (shuffles),
// which then issues the atomics — one triple per distinct pixel per warp.
LT_DEVINL void flush_film_warp(double* film, int64_t pix, double acc[3]) {
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

#ifndef LT_MIN_BLOCKS_F32
#define LT_MIN_BLOCKS_F32 6
#endif
#ifndef LT_MIN_BLOCKS_F64
#define LT_MIN_BLOCKS_F64 3
#endif

template <typename R, int TRAV, int SEM>
__global__ void __launch_bounds__(LT_BLOCK, sizeof(R) == 4 ? LT_MIN_BLOCKS_F32 : LT_MIN_BLOCKS_F64) walk_megakernel(const WalkParams<R> P) {
    extern __shared__ __align__(16) char smem[];
    const TriRec<R>* tris = P.S.tris;
    const MatRec<R>* mats = P.S.mats;
    const CompatNode<R>* cnodes = P.S.cnodes;
    const CompatLeaf<R>* cleaves = P.S.cleaves;
    if (TRAV != LT_TRAV_LBVH) {
        char* cur = smem;
        tris = stage_array(cur, P.S.tris, P.S.n_tris);
        mats = stage_array(cur, P.S.mats, P.S.n_mats);
        if (TRAV == LT_TRAV_COMPAT) {
            cnodes = stage_array(cur, P.S.cnodes, P.S.n_cnodes);
            cleaves = stage_array(cur, P.S.cleaves, P.S.n_cleaves);
        }
        __syncthreads();
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
    int tstack[TRAV == LT_TRAV_LBVH ? LT_LBVH_STACK : 1];
    LbvhState<R> ts;
    ts.cur = LT_LBVH_DONE; ts.sp = 0; ts.prim = -1; ts.best = (R)0;
    V3<R> tinv = mk<R>((R)0, (R)0, (R)0);
    bool tracing = false;
    Stats st;
    uint32_t n_ext = 0, n_shadow = 0, n_paths = 0, lidx_overflow = 0, frame_overflow = 0, max_frames = 0;
    uint32_t lane_steps = 0, warp_iters = 0, n_memo = 0;
    double acc[3] = {0, 0, 0};
    int64_t acc_pix = -1;

    for (int64_t g = gwarp; g * ppw < P.total_paths; g += nwarps) {
        int64_t next = g * ppw;
        const int64_t end = min(P.total_paths, next + ppw);
        // (pixel, sample) of the granule's first index: one 64-bit division per granule, 32-bit ones per path
        const int64_t g_pix = next / P.ns;
        const uint32_t g_s = (uint32_t)(next - g_pix * P.ns);
        const int64_t g_first = next;
        for (;;) {
            // ---- regeneration: idle lanes take the next indices of the granule
            bool need = ln.qtype == 0;
            unsigned m = __ballot_sync(full, need);
            if (m) {
                int rank = __popc(m & ((1u << lane) - 1u));
                int64_t mine = next + rank;
                next += __popc(m);
                if (need && mine < end) {
                    uint32_t off = (uint32_t)(mine - g_first) + g_s;
                    uint32_t dp = off / (uint32_t)P.ns;
                    int64_t pix = P.pixel0 + g_pix + dp;
                    int s_rel = (int)(off - dp * (uint32_t)P.ns);
                    if (pix != acc_pix) { flush_film(P.film, acc_pix, acc); acc_pix = pix; }
                    start_path<R>(P, ln, pix, s_rel);
                    if (ln.qtype == 0) {   // max_depth == 0
                        n_paths++;
                        if (P.samples) { double* o = P.samples + ((size_t)pix * P.ns + s_rel) * 3; o[0] = o[1] = o[2] = 0; }
                    }
                }
            }
            bool active = ln.qtype != 0;
            unsigned am = __ballot_sync(full, active);
            if (!am) break;
            warp_iters++;
            if (TRAV == LT_TRAV_LBVH) {
                if (active && !tracing) {
                    lbvh_begin<R>(ts, P.S.n_tris, ln.tlimit, true);
                    tinv = mk<R>((R)1 / ln.d.x, (R)1 / ln.d.y, (R)1 / ln.d.z);
                    tracing = true;
                }
                lbvh_run<R>(full, ts, tstack, P.S.lbvh, tris, ln.o, ln.d, tinv, eps, ln.qtype == 2, ln.skip, st, P.lbvh_chunk);
                if (active) lane_steps++;
                active = tracing && ts.cur == LT_LBVH_DONE;       // only lanes whose query just completed go on to shade
                if (active) tracing = false;
            }
            if (active) {
                if (TRAV != LT_TRAV_LBVH) lane_steps++;
                // ---- the one ray query of this iteration
                int prim = ts.prim;
                R t = ts.best;
                const bool anyhit = ln.qtype == 2;
                if (TRAV == LT_TRAV_COMPAT)
                    compat_query<R>(am, cnodes, cleaves, P.S.n_cleaves, tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st);
                else if (TRAV == LT_TRAV_NEAREST)
                    nearest_query<R>(tris, P.S.n_tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st);
                (void)anyhit;
                int stt;
                R adv_rand0 = (R)0;
                if (ln.qtype == 1) {
                    n_ext++;
                    stt = shade_extension<R, SEM>(P, ln, prim, t, tris, mats, lidx_overflow, adv_rand0);
                } else {
                    n_shadow++;
                    stt = shade_shadow<R, SEM, LT_LOCAL_FRAMES>(P, ln, frames, gtid, prim >= 0, frame_overflow, max_frames, adv_rand0);
                }
                bool alive = true;
                if (stt != ST_QUERY) {      // roulette / bounce++ / frame pops: one shared copy for every branch
                    alive = advance<R, SEM, LT_LOCAL_FRAMES>(P, ln, frames, gtid, stt == ST_ADVANCE_ENDED, adv_rand0, n_memo);
                    ln.qtype = 1;
                    ln.tlimit = Lim<R>::inf();
                }
                if (!alive) {
                    // ---- tally (R/src/path_tracing.py:297-298)
                    n_paths++;
                    acc[0] += (double)ln.L.x; acc[1] += (double)ln.L.y; acc[2] += (double)ln.L.z;
                    if (P.samples) {
                        double* o = P.samples + ((size_t)ln.pix * P.ns + ln.s_rel) * 3;
                        o[0] = (double)ln.L.x; o[1] = (double)ln.L.y; o[2] = (double)ln.L.z;
                    }
                    ln.qtype = 0;
                }
            }
        }
        flush_film_warp(P.film, acc_pix, acc);
        acc_pix = -1;
    }
    // ---- counters: warp-reduced, one atomic per warp and field
    if (P.counters) {
        unsigned long long v[13] = {n_paths, (unsigned long long)n_ext + n_memo, n_shadow, st.box_tests, st.tri_tests, 0, lidx_overflow,
                                    frame_overflow, lane_steps, (unsigned long long)warp_iters, 0, 0, n_memo};
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            if (k == 5 || k == 10 || k == 11) continue;
            unsigned long long x = v[k];
            if (k == 9) { x = lane == 0 ? x * 32ull : 0ull; }
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
            if (lane == 0 && x) atomicAdd(P.counters + k, x);
        }
        if (SEM == LT_SEM_REFERENCE) {
            uint32_t mf = max(max_frames, 1u);
            for (int o = 16; o > 0; o >>= 1) mf = max(mf, __shfl_xor_sync(full, mf, o));
            if (lane == 0) atomicMax(P.counters + 5, (unsigned long long)mf);
        } else if (lane == 0 && threadIdx.x == 0 && blockIdx.x == 0) {
            atomicMax(P.counters + 5, 1ull);
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
                compat_query<R>(am, P.S.cnodes, P.S.cleaves, P.S.n_cleaves, P.S.tris, o, d, (R)P.eps, Lim<R>::inf(), false, -1, prim, t, st);
            else if (TRAV == LT_TRAV_NEAREST)
                nearest_query<R>(P.S.tris, P.S.n_tris, o, d, (R)P.eps, Lim<R>::inf(), false, -1, prim, t, st);
            else
                lbvh_query<R>(am, P.S.lbvh, P.S.tris, P.S.n_tris, o, d, (R)P.eps, Lim<R>::inf(), false, -1, prim, t, st);
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
template <typename R> int walk_kernel_attrs(int trav, int sem, int* regs, int* local_bytes, int* max_blocks_per_sm, size_t smem);

}  // namespace lt
