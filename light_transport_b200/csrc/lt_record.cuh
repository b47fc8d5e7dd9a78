// lt_record.cuh — the walk with every per-sample side output of the reference, and the
// finite-difference gradient replays.
//
// The reference's render_scene_samples (R/src/path_tracing.py:251-303) keeps, besides the radiance:
//   record_log_pdf[i,j,s,:]        the array trace_path RETURNS.  Each frame starts from zeros; after a
//                                  recursive call the CALLEE's array replaces the caller's (:93) and the
//                                  caller adds log(pdf) at its own bounce (:94) — so a frame's array is
//                                  "last recursion's array + one entry", earlier entries are lost.
//   scene.record_log_pdf[...]      written by every frame after its recursive call returns (:95-96):
//                                  the last writer of an index wins.
//   scene.bounce_record[...]       ones, zeroed from the bounce at which any frame terminates by a miss
//                                  (:50-52), a zero pdf (:77-79) or the roulette (:151-153).
//   direct/indirect_light_list     of the ROOT frame, per bounce (:99-100).
// calculate_gradients (:203-249) then replays the walk 4*D times per sample with one logit shifted by
// +-0.01 (rand = 1/(1+exp(logit)), Q16), gated by bounce_record (calculate_grad=True, :37-39), and
// keeps only the replays' returned record_log_pdf arrays.
//
// This kernel is the notebook's whole compute, not the throughput path: one lane = one work item
// (a sample, or one replay of a sample), REFERENCE semantics only, frames in local memory (D <= 32),
// the same query and shading code as the megakernel, one query per warp-uniform iteration.
#pragma once
#include "lt_megakernel.cuh"

namespace lt {

#define LT_REC_FRAMES 32

struct RecParams {
    double* rlp_a;            // base: record_log_pdf [items][D]; grad: rlp_minus [paths][2D][D]
    double* rlp_b;            // grad: rlp_plus
    double* scene_rlp;        // [paths][D]                 (base, out)
    signed char* bounce_record;  // [paths][D]              (base: out; grad: in)
    double* direct;           // [paths][3][D]              (base, out, optional)
    double* indirect;
    const double* logit0;     // [paths][D]                 (grad, in)
    const double* logit1;
    int grad;                 // 0 base walk with records, 1 gradient replays
    int gate;                 // gradient replays stop where bounce_record is 0 (calculate_grad=True, :37-39); 0: the "more samples" block
    int mult;                 // work items per sample: 1 (base) or 4*D (grad)
    long long row_base;       // pixel sub-range: array rows are path_id - row_base (= pixel0 * spp)
};

template <typename R>
struct RecLane {
    R lp[LT_REC_FRAMES];      // log(pdf) of the pending diffuse hit of each suspended frame
    unsigned long long has;   // bit k: the frame at stack depth k holds a non-zero record_log_pdf array
    int min_zero;             // bounce_record is zero from here on
    int variant;              // grad: which*2D + ri*2 + sign
    V3<R> Lsnap;              // root frame: L right after its direct term was added
    double* row;              // this item's returned-array row
};

template <typename R>
LT_DEVINL void rec_rand(const WalkParams<R>& P, const RecParams& RP, const Lane<R>& ln, const RecLane<R>& rl, int bounce, R& r0, R& r1) {
    if (!RP.grad) { get_rand<R>(P, ln.path_id, bounce, r0, r1); return; }      // (replay rows: P.row_base)
    const int D = P.max_depth;
    const int which = rl.variant / (2 * D), ri = (rl.variant % (2 * D)) / 2, sign = rl.variant & 1;
    double l0 = RP.logit0[(size_t)(ln.path_id - RP.row_base) * D + bounce], l1 = RP.logit1[(size_t)(ln.path_id - RP.row_base) * D + bounce];
    const double s = sign ? 0.01 : -0.01;                              // :210-214, :230-234
    if (bounce == ri) { if (which == 0) l0 += s; else l1 += s; }
    r0 = (R)(1.0 / (1.0 + exp(l0)));                                   // sigmoid of :19-21
    r1 = (R)(1.0 / (1.0 + exp(l1)));
}

template <typename R>
LT_DEVINL bool rec_gate(const WalkParams<R>& P, const RecParams& RP, const Lane<R>& ln) {   // :37-39, then :41
    if (RP.grad && RP.gate && ln.bounce < P.max_depth && RP.bounce_record[(size_t)(ln.path_id - RP.row_base) * P.max_depth + ln.bounce] == 0) return true;
    return ln.bounce >= P.max_depth;
}

// the bounce tail with records: roulette, bounce += 1, gate/depth test; on frame end pop the parent,
// hand the child's array over (:93-96) and resume the parent at ITS roulette.
template <typename R>
LT_DEVINL bool rec_advance(const WalkParams<R>& P, const RecParams& RP, Lane<R>& ln, RecLane<R>& rl, Frame<R>* frames, bool ended, R rand_0) {
    const int D = P.max_depth;
    for (;;) {
        if (!ended) {
            bool die = false;
            if (ln.bounce > P.rr_start) {
                R om = (R)1 - ln.T.y, fl = (R)P.rr_floor;
                R r_r = om > fl ? om : fl;
                if (rand_0 < r_r) { die = true; if (!RP.grad) rl.min_zero = min(rl.min_zero, ln.bounce + 1); }
                else { R k = (R)1 - r_r; ln.T = mk<R>(ln.T.x / k, ln.T.y / k, ln.T.z / k); }
            }
            if (!die) {
                ln.bounce += 1;
                if (!rec_gate<R>(P, RP, ln)) return true;
            }
        }
        if (ln.nframes == 0) return false;
        // ---- the child (stack depth nframes) returns to its parent
        const bool child_has = (rl.has >> ln.nframes) & 1ull;
        ln.nframes -= 1;
        const Frame<R>& f = frames[ln.nframes];
        ln.T = mk<R>(f.T[0], f.T[1], f.T[2]);
        ln.W = mk<R>(f.W[0], f.W[1], f.W[2]);
        ln.bounce = f.bounce;
        rand_0 = f.rand0;
        const double lp = (double)rl.lp[ln.nframes];
        if (!child_has) for (int b = 0; b < D; ++b) rl.row[b] = 0.0;       // :93 the callee's (all-zero) array replaces ours
        rl.row[ln.bounce] = lp;                                            // :94
        rl.has |= 1ull << ln.nframes;
        if (!RP.grad) {
            RP.scene_rlp[(size_t)(ln.path_id - RP.row_base) * D + ln.bounce] = lp;         // :95-96
            if (ln.nframes == 0 && RP.indirect) {                           // :98,100 indirect = throughput * tp
                double* il = RP.indirect + (size_t)(ln.path_id - RP.row_base) * 3 * D;
                il[0 * D + ln.bounce] = (double)(ln.L.x - rl.Lsnap.x);
                il[1 * D + ln.bounce] = (double)(ln.L.y - rl.Lsnap.y);
                il[2 * D + ln.bounce] = (double)(ln.L.z - rl.Lsnap.z);
            }
        }
        ended = false;
    }
}

template <typename R, int TRAV>
__global__ void __launch_bounds__(LT_BLOCK) record_kernel(const WalkParams<R> P, const RecParams RP) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int D = P.max_depth;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5, gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const R eps = (R)P.eps;
    const int64_t items = P.total_paths;            // samples * mult
    const int64_t per_warp = (items + nwarps - 1) / nwarps;
    int64_t next = gwarp * per_warp;
    const int64_t end = min(items, next + per_warp);
    Frame<R> frames[LT_REC_FRAMES];
    Lane<R> ln;
    RecLane<R> rl;
    ln.qtype = 0;
    Stats st;
    uint32_t n_ext = 0, n_shadow = 0, n_paths = 0, lidx_overflow = 0, frame_overflow = 0, max_frames = 0;

    for (;;) {
        bool need = ln.qtype == 0;
        unsigned m = __ballot_sync(full, need);
        if (m) {
            int64_t mine = next + __popc(m & ((1u << lane) - 1u));
            next += __popc(m);
            if (need && mine < end) {
                const int64_t sample = mine / RP.mult;                      // (pixel - pixel0) * ns + s_rel
                const int64_t pix = P.pixel0 + sample / P.ns;
                const int s_rel = (int)(sample % P.ns);
                start_path<R>(P, ln, pix, s_rel);                           // camera ray from the BASE rand_0 (:280-286)
                rl.variant = (int)(mine % RP.mult);
                rl.has = 0; rl.min_zero = D;
                rl.Lsnap = mk<R>((R)0, (R)0, (R)0);
                if (!RP.grad) {
                    rl.row = RP.rlp_a + (size_t)(ln.path_id - RP.row_base) * D;
                    for (int b = 0; b < D; ++b) RP.scene_rlp[(size_t)(ln.path_id - RP.row_base) * D + b] = 0.0;
                    if (RP.direct) for (int k = 0; k < 3 * D; ++k) { RP.direct[(size_t)(ln.path_id - RP.row_base) * 3 * D + k] = 0.0; RP.indirect[(size_t)(ln.path_id - RP.row_base) * 3 * D + k] = 0.0; }
                } else {
                    const int which = rl.variant / (2 * D), ri = (rl.variant % (2 * D)) / 2, sign = rl.variant & 1;
                    rl.row = (sign ? RP.rlp_b : RP.rlp_a) + ((size_t)(ln.path_id - RP.row_base) * 2 * D + (size_t)(which * D + ri)) * D;
                }
                for (int b = 0; b < D; ++b) rl.row[b] = 0.0;
                if (ln.qtype != 0 && rec_gate<R>(P, RP, ln)) ln.qtype = 0;   // gated at bounce 0: nothing traced
                if (ln.qtype == 0) {                                          // ended before the first query
                    n_paths++;
                    if (!RP.grad) for (int b = 0; b < D; ++b) RP.bounce_record[(size_t)(ln.path_id - RP.row_base) * D + b] = 1;
                    if (!RP.grad && P.samples) { double* o = P.samples + ((size_t)(pix - P.pixel0) * P.ns + s_rel) * 3; o[0] = o[1] = o[2] = 0; }
                }
            }
        }
        const bool active = ln.qtype != 0;
        const unsigned am = __ballot_sync(full, active);
        if (!am) { if (next >= end) break; else continue; }
        if (active) {
            int prim;
            R t;
            const bool anyhit = ln.qtype == 2;
            if (TRAV == LT_TRAV_COMPAT)
                compat_query<R>(am, P.S.cnodes, P.S.cleaves, P.S.n_cleaves, nullptr, nullptr, P.S.tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st, true,
                                WarpScratch<R>(), ChainTab(), true, ln.qshift);
            else if (TRAV == LT_TRAV_NEAREST)
                nearest_query<R>(P.S.tris, P.S.n_tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st);
            else
                lbvh_query<R>(am, P.S, P.S.tris, ln.o, ln.d, eps, ln.tlimit, anyhit, ln.skip, prim, t, st);
            int stt;
            R adv_rand0 = (R)0;
            if (ln.qtype == 1) {
                n_ext++;
                // shade_extension draws rand(path, bounce) itself; the replays need their own numbers, so
                // the random row is swapped in through a scoped override of the lane's replay inputs
                R r0, r1;
                rec_rand<R>(P, RP, ln, rl, ln.bounce, r0, r1);
                stt = shade_extension_with<R, LT_SEM_REFERENCE>(P, ln, prim, t, P.S.tris, P.S.mats, lidx_overflow, adv_rand0, r0, r1);
                if (prim < 0 && !RP.grad) rl.min_zero = min(rl.min_zero, ln.bounce);                  // :50-52
                if (prim >= 0 && (ln.pdf_zero & 1) && !RP.grad) rl.min_zero = min(rl.min_zero, ln.bounce + 1);   // :77-79
            } else {
                n_shadow++;
                const bool occluded = prim >= 0 || (ln.pdf_zero & 2);
                if (ln.pdf_zero & 1) {
                    ln.o = ln.new_o; ln.d = ln.new_d; ln.skip = (ln.pdf_zero >> 2) - 1;               // the untouched incoming ray
                    if (!RP.grad) rl.min_zero = min(rl.min_zero, ln.bounce + 1);                      // :77-79
                    stt = ST_ADVANCE_ENDED;
                } else {
                    ln.T = ln.T * ln.tfac;                                                            // :92
                    if (!occluded) ln.L = ln.L + ln.W * ln.contrib;
                    if (ln.nframes == 0) {
                        rl.Lsnap = ln.L;
                        if (!RP.grad && RP.direct) {                                                  // :99
                            double* dl = RP.direct + (size_t)(ln.path_id - RP.row_base) * 3 * D;
                            dl[0 * D + ln.bounce] = occluded ? 0.0 : (double)ln.contrib.x;
                            dl[1 * D + ln.bounce] = occluded ? 0.0 : (double)ln.contrib.y;
                            dl[2 * D + ln.bounce] = occluded ? 0.0 : (double)ln.contrib.z;
                        }
                    }
                    if (ln.nframes >= LT_REC_FRAMES - 1) frame_overflow++;
                    else {
                        Frame<R>& f = frames[ln.nframes];
                        f.T[0] = ln.T.x; f.T[1] = ln.T.y; f.T[2] = ln.T.z;
                        f.W[0] = ln.W.x; f.W[1] = ln.W.y; f.W[2] = ln.W.z;
                        f.bounce = ln.bounce; f.rand0 = ln.rand0_pending;
                        rl.lp[ln.nframes] = (R)log((double)ln.pdf_pending);
                        ln.nframes += 1;
                        rl.has &= ~(1ull << ln.nframes);                                              // the child starts from zeros
                    }
                    if ((uint32_t)ln.nframes + 1 > max_frames) max_frames = (uint32_t)ln.nframes + 1;
                    ln.W = ln.W * ln.T;
                    ln.T = mk<R>((R)1, (R)1, (R)1);
                    ln.o = ln.new_o; ln.d = ln.new_d;
                    ln.bounce += 1;
                    if (!rec_gate<R>(P, RP, ln)) { ln.qtype = 1; ln.tlimit = Lim<R>::inf(); stt = ST_QUERY; }
                    else stt = ST_ADVANCE_ENDED;
                }
            }
            bool alive = true;
            if (stt != ST_QUERY) {
                alive = rec_advance<R>(P, RP, ln, rl, frames, stt == ST_ADVANCE_ENDED, adv_rand0);
                ln.qtype = 1;
                ln.tlimit = Lim<R>::inf();
            }
            if (!alive) {
                n_paths++;
                if (!(rl.has & 1ull)) for (int b = 0; b < D; ++b) rl.row[b] = 0.0;     // the root never recursed: zeros
                if (!RP.grad) {
                    for (int b = 0; b < D; ++b) RP.bounce_record[(size_t)(ln.path_id - RP.row_base) * D + b] = b < rl.min_zero ? 1 : 0;
                    if (P.samples) {
                        double* o = P.samples + ((size_t)(ln.pix - P.pixel0) * P.ns + ln.s_rel) * 3;
                        o[0] = (double)ln.L.x; o[1] = (double)ln.L.y; o[2] = (double)ln.L.z;
                    }
                    if (P.film) {
                        atomicAdd(P.film + (size_t)ln.pix * 3 + 0, (double)ln.L.x);
                        atomicAdd(P.film + (size_t)ln.pix * 3 + 1, (double)ln.L.y);
                        atomicAdd(P.film + (size_t)ln.pix * 3 + 2, (double)ln.L.z);
                    }
                }
                ln.qtype = 0;
            }
        }
    }
    if (P.counters) {
        unsigned long long v[8] = {n_paths, n_ext, n_shadow, st.box_tests, st.tri_tests, 0, lidx_overflow, frame_overflow};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k == 5) continue;
            unsigned long long x = v[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
            if (lane == 0 && x) atomicAdd(P.counters + k, x);
        }
        uint32_t mf = max(max_frames, 1u);
        for (int o = 16; o > 0; o >>= 1) mf = max(mf, __shfl_xor_sync(full, mf, o));
        if (lane == 0) atomicMax(P.counters + 5, (unsigned long long)mf);
    }
}

template <typename R> int launch_record(const WalkParams<R>& P, const RecParams& RP, int grid, cudaStream_t stream);

}  // namespace lt
