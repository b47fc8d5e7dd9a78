// lt_wavefront.cuh — wavefront path scheduler, CTA-resident.
//
// The megakernel keeps a path in one lane for life, so the lanes of a warp sit in different shading
// branches (ncu round 1: shading executes with 7-8 of 32 lanes).  Here a CTA owns a POOL of path
// slots whose state lives in SHARED memory (structure-of-arrays, one 32-bit word plane per field, so
// consecutive slots hit consecutive banks) and every round runs two converged stages:
//
//   TRACE   thread t <-> slot t: regenerate if the slot is idle (generate queue), run the slot's ONE
//           pending query (extend or shadow — identical code), store the hit, and append the slot to
//           the queue of its class: {shadow-resolve, miss, diffuse, specular}.  Queue append is a warp
//           ballot + popc prefix rank with one shared atomicAdd per warp and class.
//   SHADE   thread t <-> task t of the concatenated class queues: the warps of a round therefore run
//           ONE shading branch each (at most three warps straddle a class boundary), then the shared
//           bounce tail advance(); paths that end are tallied (terminate) and their slot becomes idle.
//
// Work is pulled by persistent CTAs: a block-aggregated atomicAdd on a global granule counter, then
// warp-aggregated atomicAdds on the CTA's shared `next` index while lanes regenerate.  Film: ended
// paths of a warp are summed per pixel with match_any + shuffles and one lane issues the fp64 atomics.
// Suspended REFERENCE frames live in the global arena (L2), touched only at push/pop.
// The per-path arithmetic is the same code as the megakernel's (lt_walk.cuh), so both schedulers
// produce identical per-path results.
#pragma once
#include "lt_megakernel.cuh"

namespace lt {

#define LT_WF_CLASSES 4
enum { WF_RESOLVE = 0, WF_MISS = 1, WF_DIFFUSE = 2, WF_SPECULAR = 3 };

template <typename R> struct WfCfg;
template <> struct WfCfg<float> { static constexpr int NT = 256, MINB = 3; };
template <> struct WfCfg<double> { static constexpr int NT = 128, MINB = 2; };

template <typename R> struct LaneWords {
    static constexpr int N = (int)(sizeof(Lane<R>) / 4);
    static constexpr int QUERY = (int)(offsetof(Lane<R>, hit_prim) / 4);        // o, d, tlimit, skip, qtype
    static constexpr int HIT_END = (int)(offsetof(Lane<R>, pix) / 4);           // + hit_prim, hit_pad, hit_t
};

template <typename R, int W0, int W1>
LT_DEVINL void lane_load(const uint32_t* sm, int nt, int slot, Lane<R>& ln) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&ln);
#pragma unroll
    for (int k = W0; k < W1; ++k) w[k] = sm[k * nt + slot];
}
template <typename R, int W0, int W1>
LT_DEVINL void lane_store(uint32_t* sm, int nt, int slot, const Lane<R>& ln) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&ln);
#pragma unroll
    for (int k = W0; k < W1; ++k) sm[k * nt + slot] = w[k];
}

struct WfShared {
    int next, end;                         // current granule: next work index, end
    long long g_pix;                       // (pixel, sample) of the granule's first index
    unsigned g_s;
    long long g_first;
    int more;                              // global work may remain
    int count[2][LT_WF_CLASSES];           // class queue sizes, double-buffered by round parity
};

template <typename R, int TRAV, int SEM>
__global__ void __launch_bounds__(WfCfg<R>::NT, WfCfg<R>::MINB) walk_wavefront(const WalkParams<R> P, unsigned long long* granule_counter) {
    constexpr int NT = WfCfg<R>::NT;
    constexpr int NWORDS = LaneWords<R>::N;
    extern __shared__ __align__(16) char smem[];
    // ---- shared-memory carve-up: scene | lane planes | class queues | control
    char* cur = smem;
    const TriRec<R>* tris = P.S.tris;
    const MatRec<R>* mats = P.S.mats;
    const CompatNode<R>* cnodes = P.S.cnodes;
    const CompatLeaf<R>* cleaves = P.S.cleaves;
    if (TRAV != LT_TRAV_LBVH) {
        tris = stage_array(cur, P.S.tris, P.S.n_tris);
        mats = stage_array(cur, P.S.mats, P.S.n_mats);
        if (TRAV == LT_TRAV_COMPAT) {
            cnodes = stage_array(cur, P.S.cnodes, P.S.n_cnodes);
            cleaves = stage_array(cur, P.S.cleaves, P.S.n_cleaves);
        }
    }
    // the megakernel's traversal tables (lt_intersect.cuh), built from the global copy: the barrier below publishes them
    const float* slabtab = nullptr;
    const uint8_t* flatprim = nullptr;
    if (TRAV == LT_TRAV_COMPAT) {
        build_slab_pairs(reinterpret_cast<float*>(cur), P.S.cleaves, P.S.n_cleaves);
        slabtab = reinterpret_cast<const float*>(cur);
        cur += (slab_pairs_bytes(P.S.n_cleaves) + 15) / 16 * 16;
        if (compat_small(P.S.n_cleaves, P.S.n_cflat, P.S.n_tris)) {
            build_flat_prims<R>(reinterpret_cast<uint8_t*>(cur), P.S.cleaves, P.S.n_cleaves);
            flatprim = reinterpret_cast<const uint8_t*>(cur);
            cur += 32;
        }
    }
    uint32_t* lanes = reinterpret_cast<uint32_t*>(cur);
    cur += (size_t)NWORDS * NT * 4;
    uint16_t* queue = reinterpret_cast<uint16_t*>(cur);          // [class][NT]
    cur += (size_t)LT_WF_CLASSES * NT * 2;
    WfShared* sh = reinterpret_cast<WfShared*>(cur);

    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t gslot = (int64_t)blockIdx.x * NT + tid;        // arena column of this CTA's slot `tid`
    const R eps = (R)P.eps;
    const long long ppg = P.paths_per_warp;                      // granule size (paths per CTA granule)

    if (tid == 0) {
        sh->next = 0; sh->end = 0; sh->more = 1;
        for (int b = 0; b < 2; ++b) for (int c = 0; c < LT_WF_CLASSES; ++c) sh->count[b][c] = 0;
    }
    for (int k = 0; k < NWORDS; ++k) lanes[k * NT + tid] = 0u;   // every slot idle (qtype == 0)
    __syncthreads();

    Stats st;
    uint32_t n_ext = 0, n_shadow = 0, n_paths = 0, lidx_overflow = 0, frame_overflow = 0, max_frames = 0;
    uint32_t lane_steps = 0, warp_iters = 0, n_memo = 0;

    for (int round = 0;; ++round) {
        const int par = round & 1;
        // ---- granule supply (block-aggregated pull from the global counter)
        if (tid == 0) {
            if (sh->more && sh->next >= sh->end) {
                long long g = (long long)atomicAdd(granule_counter, 1ull);
                long long first = g * ppg;
                if (first >= P.total_paths) sh->more = 0;
                else {
                    long long last = min((long long)P.total_paths, first + ppg);
                    sh->g_first = first;
                    sh->g_pix = first / P.ns;
                    sh->g_s = (unsigned)(first - sh->g_pix * P.ns);
                    sh->next = 0;
                    sh->end = (int)(last - first);
                }
            }
        }
        __syncthreads();
        const int more_now = sh->more;       // stable for the whole round (only thread 0 writes it, above)
        // =================== TRACE: thread t <-> slot t ===================
        Lane<R> ln;
        {
            uint32_t* w = reinterpret_cast<uint32_t*>(&ln);
#pragma unroll
            for (int k = 0; k < NWORDS; ++k) w[k] = 0u;
        }
        lane_load<R, 0, LaneWords<R>::QUERY>(lanes, NT, tid, ln);
        // generate: idle slots take the next indices of the granule (warp-aggregated shared atomic)
        {
            bool need = ln.qtype == 0;
            unsigned m = __ballot_sync(full, need);
            if (m && sh->next < sh->end) {
                int base = 0;
                if (lane == (__ffs(m) - 1)) base = atomicAdd(&sh->next, __popc(m));
                base = __shfl_sync(full, base, __ffs(m) - 1);
                int mine = base + __popc(m & ((1u << lane) - 1u));
                if (need && mine < sh->end) {
                    unsigned off = (unsigned)mine + sh->g_s;
                    unsigned dp = off / (unsigned)P.ns;
                    long long pix = P.pixel0 + sh->g_pix + dp;
                    int s_rel = (int)(off - dp * (unsigned)P.ns);
                    start_path<R>(P, ln, pix, s_rel);
                    lane_store<R, 0, NWORDS>(lanes, NT, tid, ln);       // whole lane: fresh path state
                    if (ln.qtype == 0) {                                // max_depth == 0: ends at once
                        n_paths++;
                        if (P.samples) { double* o = P.samples + ((size_t)pix * P.ns + s_rel) * 3; o[0] = o[1] = o[2] = 0; }
                    }
                }
            }
        }
        const bool active = ln.qtype != 0;
        const unsigned am = __ballot_sync(full, active);
        int cls = -1;
        if (am) warp_iters++;
        if (active) {
            lane_steps++;
            int prim;
            R t;
            const bool anyhit = ln.qtype == 2;
            if (TRAV == LT_TRAV_COMPAT)
                compat_query<R>(am, cnodes, cleaves, P.S.n_cleaves, slabtab, flatprim, tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st, true,
                                WarpScratch<R>(), ChainTab(), true, ln.qshift);
            else if (TRAV == LT_TRAV_NEAREST)
                nearest_query<R>(tris, P.S.n_tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st);
            else
                lbvh_query<R>(am, P.S, tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st);
            ln.hit_prim = prim;
            ln.hit_pad = 0;
            ln.hit_t = t;
            lane_store<R, LaneWords<R>::QUERY, LaneWords<R>::HIT_END>(lanes, NT, tid, ln);
            if (ln.qtype == 2) { n_shadow++; cls = WF_RESOLVE; }
            else {
                n_ext++;
                if (prim < 0) cls = WF_MISS;
                else cls = (mats[tris[prim].mat].flags & 1) ? WF_DIFFUSE : WF_SPECULAR;
            }
        }
        // enqueue by class: ballot + popc rank, one shared atomic per warp and class
#pragma unroll
        for (int c = 0; c < LT_WF_CLASSES; ++c) {
            unsigned m = __ballot_sync(full, cls == c);
            if (m) {
                int base = 0;
                if (lane == (__ffs(m) - 1)) base = atomicAdd(&sh->count[par][c], __popc(m));
                base = __shfl_sync(full, base, __ffs(m) - 1);
                if (cls == c) queue[c * NT + base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)tid;
            }
        }
        const int any_active = __syncthreads_or(active ? 1 : 0);
        if (!any_active && !more_now) break;             // pool drained, granule exhausted, no granules left
        if (tid == 0)                                    // everyone is past round-1's SHADE: recycle its counters
            for (int c = 0; c < LT_WF_CLASSES; ++c) sh->count[par ^ 1][c] = 0;
        // =================== SHADE: thread t <-> task t of the concatenated class queues ===================
        const int c0 = sh->count[par][0], c1 = c0 + sh->count[par][1], c2 = c1 + sh->count[par][2], c3 = c2 + sh->count[par][3];
        bool ended_path = false;
        if (tid < c3) {
            const int tc = tid < c0 ? 0 : (tid < c1 ? 1 : (tid < c2 ? 2 : 3));
            const int slot = queue[tc * NT + (tid - (tc == 0 ? 0 : (tc == 1 ? c0 : (tc == 2 ? c1 : c2))))];
            const int64_t acol = (int64_t)blockIdx.x * NT + slot;
            lane_load<R, 0, NWORDS>(lanes, NT, slot, ln);
            int stt;
            R adv_rand0 = (R)0;
            if (tc == WF_RESOLVE)
                stt = shade_shadow<R, SEM, 0>(P, ln, nullptr, acol, ln.hit_prim >= 0, frame_overflow, max_frames, adv_rand0);
            else
                stt = shade_extension<R, SEM>(P, ln, ln.hit_prim, ln.hit_t, tris, mats, lidx_overflow, adv_rand0);
            bool alive = true;
            if (stt != ST_QUERY) {
                alive = advance<R, SEM, 0>(P, ln, nullptr, acol, stt == ST_ADVANCE_ENDED, adv_rand0, n_memo);
                ln.qtype = alive ? 1 : 0;
                ln.tlimit = Lim<R>::inf();
            }
            ended_path = !alive;
            lane_store<R, 0, NWORDS>(lanes, NT, slot, ln);
        }
        // ---- terminate: tally ended paths (R/src/path_tracing.py:297-298), per-pixel warp aggregation
        {
            unsigned em = __ballot_sync(full, ended_path);
            if (em) {
                long long pix = ended_path ? (long long)ln.pix : -1 - lane;
                double a0 = ended_path ? (double)ln.L.x : 0.0, a1 = ended_path ? (double)ln.L.y : 0.0, a2 = ended_path ? (double)ln.L.z : 0.0;
                unsigned same = __match_any_sync(full, pix);
                double s0 = 0, s1 = 0, s2 = 0;
                for (unsigned rest = em; rest; rest &= rest - 1) {
                    int src = __ffs(rest) - 1;
                    double b0 = __shfl_sync(full, a0, src), b1 = __shfl_sync(full, a1, src), b2 = __shfl_sync(full, a2, src);
                    if ((same >> src) & 1u) { s0 += b0; s1 += b1; s2 += b2; }
                }
                if (ended_path) {
                    n_paths++;
                    if ((__ffs(same) - 1) == lane) {
                        if (s0 != 0.0) atomicAdd(P.film + pix * 3 + 0, s0);
                        if (s1 != 0.0) atomicAdd(P.film + pix * 3 + 1, s1);
                        if (s2 != 0.0) atomicAdd(P.film + pix * 3 + 2, s2);
                    }
                    if (P.samples) {
                        double* o = P.samples + ((size_t)ln.pix * P.ns + ln.s_rel) * 3;
                        o[0] = (double)ln.L.x; o[1] = (double)ln.L.y; o[2] = (double)ln.L.z;
                    }
                }
            }
        }
        (void)gslot;
    }
    // ---- counters
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
        uint32_t mf = SEM == LT_SEM_REFERENCE ? max(max_frames, 1u) : 1u;
        for (int o = 16; o > 0; o >>= 1) mf = max(mf, __shfl_xor_sync(full, mf, o));
        if (lane == 0) atomicMax(P.counters + 5, (unsigned long long)mf);
    }
}

template <typename R> size_t wavefront_smem_bytes(size_t staged) {
    return staged + (size_t)LaneWords<R>::N * WfCfg<R>::NT * 4 + (size_t)LT_WF_CLASSES * WfCfg<R>::NT * 2 + sizeof(WfShared) + 16;
}
template <typename R> int launch_wavefront(const WalkParams<R>& P, unsigned long long* granule_counter, int grid, size_t smem, cudaStream_t stream);
template <typename R> int wavefront_kernel_attrs(int trav, int sem, int* regs, int* max_blocks_per_sm, size_t smem);

}  // namespace lt
