// lt_walk.cuh — the path walk as a per-lane state machine (shared by the megakernel and, stage by
// stage, by the wavefront scheduler).
//
// One lane = one path at a time.  Every loop iteration of a lane performs exactly ONE ray query —
// an extension ray or a shadow ray — followed by the short piece of shading that query feeds, so
// the expensive part (traversal) is convergent across the warp whatever the lanes' path states.
//
// REFERENCE semantics (R/src/path_tracing.py:23-159, Q1): trace_path recurses at every diffuse hit
// and then KEEPS ITERATING with the ray its callee left behind.  That is a depth-first walk over
// "frames" (bounce, T).  With W_child = W_parent * T_parent(at call time) the root radiance is
// sum_frames W_f * (terms added inside f), so a lane carries one accumulator L, the current frame
// (bounce, T, W) and a stack of suspended parent frames; no values are returned upwards.
// ITERATIVE semantics (R/src/path_tracing_fix1.py:17-136) is the same machine without the stack.
#pragma once
#include "lt_intersect.cuh"

namespace lt {

#define LT_LOCAL_FRAMES 12
#define LT_ARENA_VALUES 8    // T[3], W[3], bounce, rand0 per spilled frame

template <typename R>
struct WalkParams {
    SceneView<R> S;
    int32_t width, height, spp, s0, ns, max_depth, rr_start, n_light_idx, semantics, traversal;
    int32_t pixel0, paths_per_warp;
    int32_t lbvh_chunk;           // megakernel, LBVH: voted traversal steps per loop iteration
    int32_t phased;               // megakernel, COMPAT / NEAREST: phase-aligned iterations (see lt_megakernel.cuh)
    int64_t total_paths;          // (pixel1 - pixel0) * ns
    int64_t row_base;             // replayed rand_0 / rand_1 / light_idx rows are path_id - row_base (lt_render_records on a pixel sub-range; else 0)
    double rr_floor, eps, cam[3], f_distance, top, bottom, step_x, step_y;
    uint64_t seed;
    const double* rand0;          // [H][W][spp][D] or null
    const double* rand1;
    const int32_t* light_idx;     // [H][W][spp][n_light_idx] or null
    double* film;                 // [H][W][3]
    double* samples;              // [H][W][ns][3] or null
    unsigned long long* counters; // lt_counters as u64[12]
    R* arena;                     // spilled frames: [(frame - LOCAL) * LT_ARENA_VALUES + k][n_threads]
    int64_t arena_threads;
    int32_t arena_frames;
};

// what shading needs to know about the primitive that was hit
template <typename R>
struct HitRec {
    V3<R> n;                      // geometric unit normal (unflipped)
    int32_t mat, is_light;
    bool valid = false;
};
template <typename R> LT_DEVINL HitRec<R> hit_record(const TriRec<R>& t) {
    HitRec<R> h;
    h.n = mk<R>(t.n[0], t.n[1], t.n[2]); h.mat = t.mat; h.is_light = t.is_light; h.valid = true;
    return h;
}

#if LT_TRI_PLANES
LT_DEVINL HitRec<float> hit_record(const TriPad& t) {
    HitRec<float> h;
    h.n = mk<float>(t.n[0], t.n[1], t.n[2]); h.mat = t.mat; h.is_light = t.is_light; h.valid = true;
    return h;
}
#endif

template <typename R>
struct Frame {
    R T[3], W[3];
    R rand0;                      // rand_0 of the bounce at which the frame recursed (its roulette reuses it, Q6)
    int32_t bounce;
};

template <typename R>
struct Lane {
    // query (the wavefront scheduler loads only this head for its trace stage)
    V3<R> o, d;
    V3<R> qshift;                 // float32: slab-distance shift of the query ray = (reference's 1e-6 origin offset)_k / d_k, see slab()
    R tlimit;
    int32_t skip;                 // primitive the query ray starts on (float32 walk only), else -1
    int32_t qtype;                // 0 idle, 1 extension, 2 shadow
    // result of the query (wavefront scheduler only; the megakernel keeps it in registers)
    int32_t hit_prim, hit_pad;
    R hit_t;
    // path
    int32_t pix, s_rel;           // pixel (row-major) and s - s0
    uint64_t path_id;             // (pixel * spp + s): Philox counter / replay row
    int32_t bounce, nframes, nee;
    int32_t pdf_zero;             // bit 0: the pending hemisphere sample has pdf == 0; bit 1: float32 self-occluded shadow query;
                                  // bits 2..: with bit 0, 1 + the skip primitive of the incoming ray
    int32_t memo;                 // bit 0: the current extension ray is known to miss (it already did): Q1 re-traces need no
                                  // traversal.  bits 1..: how many suspended frames would re-trace it (see resume_retraces)
    V3<R> T, W, L;
    // pending across the shadow query
    V3<R> contrib, new_o, new_d, tfac;
    R rand0_pending;
    R pdf_pending;                // pdf of the pending hemisphere sample (record kernel: log(pdf), R/src/path_tracing.py:94)
};

// (pixel * spp + s): Philox counter / replay row.  Derived at each use (one 64-bit multiply-add) instead of being
// carried in two registers through the whole walk.
template <typename R>
LT_DEVINL uint64_t path_id_of(const WalkParams<R>& P, const Lane<R>& ln) {
    return (uint64_t)(uint32_t)ln.pix * (uint64_t)(uint32_t)P.spp + (uint64_t)(uint32_t)(P.s0 + ln.s_rel);
}

// random numbers of (path, bounce): replayed Scene.rand_0/rand_1 rows or Philox stream 0
template <typename R>
LT_DEVINL void get_rand(const WalkParams<R>& P, uint64_t path_id, int bounce, R& r0, R& r1) {
    if (P.rand0) {
        size_t idx = (size_t)(path_id - (uint64_t)P.row_base) * (size_t)P.max_depth + (size_t)bounce;
        r0 = (R)__ldg(P.rand0 + idx);
        r1 = (R)__ldg(P.rand1 + idx);
    } else {
        uint32_t w[4];
        philox4x32_10((uint32_t)path_id, (uint32_t)(path_id >> 32), (uint32_t)bounce, 0u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), w);
        u01(w[0], w[1], r0);
        u01(w[2], w[3], r1);
    }
}
template <typename R>
LT_DEVINL R get_rand0(const WalkParams<R>& P, uint64_t path_id, int bounce) {
    R a, b;
    get_rand(P, path_id, bounce, a, b);
    return a;
}

template <typename R>
LT_DEVINL int pick_light(const WalkParams<R>& P, Lane<R>& ln, uint32_t& overflow) {
    uint32_t k = (uint32_t)ln.nee++;
    const uint64_t path_id = path_id_of(P, ln);
    if (P.light_idx) {
        if ((int)k >= P.n_light_idx) { overflow++; return 0; }
        return __ldg(P.light_idx + (size_t)(path_id - (uint64_t)P.row_base) * (size_t)P.n_light_idx + k);
    }
    if (P.S.n_lights == 1) return 0;
    uint32_t w[4];
    philox4x32_10((uint32_t)path_id, (uint32_t)(path_id >> 32), k, 1u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), w);
    return (int)__umulhi(w[0], (uint32_t)P.S.n_lights);
}

// R/src/utils.py:131-161 (+ :114-128, :71-80).  Returns pdf; Q3: the flip tests WORLD-space z.
template <typename R>
LT_DEVINL R cosine_sample(V3<R> n, V3<R> d_in, R u0, R u1, V3<R>& out) {
    const R pi_over_2 = (R)(3.141592653589793 / 2), pi_over_4 = (R)(0.5 * (3.141592653589793 / 2));
    const R inv_pi = (R)(1 / 3.141592653589793);
    R inc_z = -d_in.z;
    R ox = (R)2.0 * u0 - (R)1, oy = (R)2.0 * u1 - (R)1, dx, dy;
    if (ox == 0 && oy == 0) { dx = 0; dy = 0; }
    else {
        R r, theta;
        if (rabs(ox) > rabs(oy)) { r = ox; theta = pi_over_4 * (oy / ox); }
        else { r = oy; theta = pi_over_2 - pi_over_4 * (ox / oy); }
        R s, c;
        if (sizeof(R) == 4) { s = (R)sinf((float)theta); c = (R)cosf((float)theta); }
        else { s = (R)sin((double)theta); c = (R)cos((double)theta); }
        dx = r * c; dy = r * s;
    }
    R zz = (R)1 - dx * dx - dy * dy;
    R z = real_sqrt<R>(zz > 0 ? zz : (R)0);
    V3<R> v2;
    if (rabs(n.x) > rabs(n.y)) {
        R s = real_sqrt<R>(n.x * n.x + n.z * n.z);
        v2 = mk<R>(-n.z / s, (R)0, n.x / s);
    } else {
        R s = real_sqrt<R>(n.y * n.y + n.z * n.z);
        v2 = mk<R>((R)0, n.z / s, -n.y / s);
    }
    V3<R> v3 = cross(n, v2);
    R oz = inc_z < 0 ? -z : z;
    R pdf = (inc_z * oz > 0) ? rabs(z) * inv_pi : (R)0;
    out = mk<R>(dx * v2.x + dy * v3.x + oz * n.x, dx * v2.y + dy * v3.y + oz * n.y, dx * v2.z + dy * v3.z + oz * n.z);
    return pdf;
}

// R/src/brdf.py:7-9
template <typename R> LT_DEVINL V3<R> reflect_dir(V3<R> v, V3<R> n) {
    R k = (R)2 * dot(v, n);
    return normalize(mk<R>(v.x - k * n.x, v.y - k * n.y, v.z - k * n.z));
}

// LF = frames kept in the caller's local array (megakernel: LT_LOCAL_FRAMES; wavefront: 0, all in the arena)
template <typename R, int LF> LT_DEVINL void frame_store(const WalkParams<R>& P, Frame<R>* local, int64_t gtid, int idx, const Lane<R>& ln, uint32_t& overflow) {
    if (idx < LF) {
        Frame<R>& f = local[idx];
        f.T[0] = ln.T.x; f.T[1] = ln.T.y; f.T[2] = ln.T.z;
        f.W[0] = ln.W.x; f.W[1] = ln.W.y; f.W[2] = ln.W.z;
        f.bounce = ln.bounce;
        f.rand0 = ln.rand0_pending;
    } else if (idx - LF < P.arena_frames) {
        R* a = P.arena + ((int64_t)(idx - LF) * LT_ARENA_VALUES) * P.arena_threads + gtid;
        int64_t s = P.arena_threads;
        a[0] = ln.T.x; a[s] = ln.T.y; a[2 * s] = ln.T.z;
        a[3 * s] = ln.W.x; a[4 * s] = ln.W.y; a[5 * s] = ln.W.z;
        a[6 * s] = (R)ln.bounce;
        a[7 * s] = ln.rand0_pending;
    } else {
        overflow++;
    }
}
template <typename R, int LF> LT_DEVINL void frame_load(const WalkParams<R>& P, const Frame<R>* local, int64_t gtid, int idx, Lane<R>& ln, R& rand0) {
    if (idx < LF) {
        const Frame<R>& f = local[idx];
        ln.T = mk<R>(f.T[0], f.T[1], f.T[2]);
        ln.W = mk<R>(f.W[0], f.W[1], f.W[2]);
        ln.bounce = f.bounce;
        rand0 = f.rand0;
    } else if (idx - LF < P.arena_frames) {
        const R* a = P.arena + ((int64_t)(idx - LF) * LT_ARENA_VALUES) * P.arena_threads + gtid;
        int64_t s = P.arena_threads;
        ln.T = mk<R>(a[0], a[s], a[2 * s]);
        ln.W = mk<R>(a[3 * s], a[4 * s], a[5 * s]);
        ln.bounce = (int)a[6 * s];
        rand0 = a[7 * s];
    }
}

// A suspended frame resumes at its Russian roulette (R/src/path_tracing.py:148-155, on the T, bounce and rand_0 it
// was suspended with) and, if it survives and bounce+1 < max_depth, traces the ray object again (Q1).  Both are
// known when the frame is pushed, so the answer is computed THERE and counted in ln.memo: when a ray misses, every
// suspended frame that resumes re-traces that same ray, misses too and ends — the whole stack unwinds, and the only
// thing left to know is how many of those re-traces the reference would have made (they count as rays).  That
// count is ln.memo >> 1; no frame is read back.  (Reading them cost a dependent local-memory load per frame:
// 5.7 % of all stall samples of the float32 walk.)
template <typename R>
LT_DEVINL int resume_retraces(const WalkParams<R>& P, V3<R> T, int bounce, R rand_0) {
    bool die = false;
    if (bounce > P.rr_start) {
        R om = (R)1 - T.y, fl = (R)P.rr_floor;
        R r_r = om > fl ? om : fl;
        die = rand_0 < r_r;
    }
    return (!die && bounce + 1 < P.max_depth) ? 1 : 0;
}

// Shared tail of every bounce: Russian roulette (R/src/path_tracing.py:148-155, Q6: reuses rand_0 of
// the frame's current bounce), bounce += 1, the max-depth test at the top of the next iteration
// (:41), and — whenever a frame ends — resumption of its parent at the roulette (Q1).
// `ended` = the current frame has already ended (miss / pdf==0 / malformed material).
// Returns true when the lane has a new extension ray to trace, false when the whole path is over.
// n_memo counts the reference's intersect_bvh calls that are answered without a traversal (see above).
template <typename R, int SEM, int LF>
LT_DEVINL bool advance(const WalkParams<R>& P, Lane<R>& ln, const Frame<R>* frames, int64_t gtid, bool ended, R rand_0, uint32_t& n_memo) {
    for (;;) {
        if (!ended) {
            bool die = false;
            if (ln.bounce > P.rr_start) {
                R om = (R)1 - ln.T.y, fl = (R)P.rr_floor;
                R r_r = om > fl ? om : fl;
                if (rand_0 < r_r) die = true;
                else { R k = (R)1 - r_r; ln.T = mk<R>(ln.T.x / k, ln.T.y / k, ln.T.z / k); }
            }
            if (!die) {
                ln.bounce += 1;
                if (ln.bounce < P.max_depth) return true;
            }
        }
        // this frame is over
        if (SEM == LT_SEM_ITERATIVE || ln.nframes == 0) return false;
        if (ln.memo & 1) {                       // the ray missed: the stack unwinds without another traversal
            n_memo += (uint32_t)(ln.memo >> 1);
            ln.nframes = 0;
            return false;
        }
        ln.nframes -= 1;
        frame_load<R, LF>(P, frames, gtid, ln.nframes, ln, rand_0);
        ln.memo -= 2 * resume_retraces<R>(P, ln.T, ln.bounce, rand_0);
        ended = false;
    }
}

// What a shading step leaves behind: a new query in the lane (ST_QUERY), or a request for the shared
// bounce tail advance() — kept in ONE place in the kernel so that lanes arriving from different
// branches (miss, specular, shadow resolve) run it together instead of through five inlined copies.
enum { ST_QUERY = 0, ST_ADVANCE = 1, ST_ADVANCE_ENDED = 2 };

// Shading after an EXTENSION query, given the bounce's two random numbers.
template <typename R, int SEM, typename TR = TriRec<R>>
LT_DEVINL int shade_extension_with(const WalkParams<R>& P, Lane<R>& ln, int prim, R t, const TR* tris, const MatRec<R>* mats,
                                   uint32_t& lidx_overflow, R& adv_rand0, R rand_0, R rand_1, HitRec<R> hit_rec = HitRec<R>()) {
    const R eps = (R)P.eps;
    const R inv_pi = (R)(1 / 3.141592653589793);
    const int32_t skip_in = ln.skip;
    if (sizeof(R) == 4 && prim >= 0) ln.skip = prim;   // every continuation ray below starts on `prim`;
                                                       // on a miss the (unchanged) ray keeps its own skip
    if (prim < 0) { ln.memo |= 1; return ST_ADVANCE_ENDED; }           // miss: frame ends (:48-53)
    ln.memo &= ~1;
    ln.pdf_zero = 0;
    if (!hit_rec.valid) hit_rec = hit_record(tris[prim]);          // else: the record already in hand (wide-BVH float32 walk)
    const HitRec<R>& tr = hit_rec;
    const MatRec<R>& m = mats[tr.mat];
    V3<R> p = mk<R>(ln.o.x + t * ln.d.x, ln.o.y + t * ln.d.y, ln.o.z + t * ln.d.z);    // utils.py:65
    V3<R> n = tr.n;
    if (tr.is_light && (SEM == LT_SEM_ITERATIVE || ln.bounce == 0)) {   // :59-60 (Q4) / fix1:45-46
        V3<R> w = SEM == LT_SEM_REFERENCE ? ln.W * ln.T : ln.T;
        ln.L = ln.L + w * m.emission;
    }
    bool inside = false;
    if (dot(n, ln.d) > 0) { n = -n; inside = true; }                    // :62-65
    if (m.flags & 1) {
        // next-event estimation, R/src/light_samples.py:35-61 — set up the shadow query
        V3<R> so = mk<R>(p.x + eps * n.x, p.y + eps * n.y, p.z + eps * n.z);
        int k = pick_light<R>(P, ln, lidx_overflow);
        const LightRec<R>& lt_ = P.S.lights[k];
        V3<R> dv = mk<R>(lt_.pos[0] - so.x, lt_.pos[1] - so.y, lt_.pos[2] - so.z);
        R mag = norm(dv);
        V3<R> sd = mk<R>(dv.x / mag, dv.y / mag, dv.z / mag);
        R cos_theta = dot(n, sd);
        R cos_phi = dot(mk<R>(lt_.n[0], lt_.n[1], lt_.n[2]), -sd);
        R g = rabs(cos_theta * cos_phi) / (mag * mag);
        ln.contrib = mk<R>(lt_.lekd[0] * (m.kd[0] * inv_pi) * g * lt_.area, lt_.lekd[1] * (m.kd[1] * inv_pi) * g * lt_.area,
                           lt_.lekd[2] * (m.kd[2] * inv_pi) * g * lt_.area);
        // indirect direction, R/src/utils.py:131-161, evaluated now so the shadow query is the only
        // thing left between here and the recursion point
        V3<R> dir;
        R pdf = cosine_sample<R>(n, ln.d, rand_0, rand_1, dir);
        // pdf == 0 (u0 == 0 exactly, or a direction with zero z): the reference has ALREADY cast the shadow ray
        // (cast_one_shadow_ray runs before the sampling, :70-75 / fix1:56-59) and then breaks without touching the ray
        // (:76-80, dropping the direct light; fix1:59-67 has already added it).  So the shadow query is still issued and
        // counted; shade_shadow ends the frame.
        ln.pdf_zero = pdf == 0 ? 1 : 0;
        ln.rand0_pending = rand_0;
        ln.pdf_pending = pdf;
        R ct = dot(dir, n);
        ln.tfac = mk<R>((m.kd[0] * inv_pi) * ct / pdf, (m.kd[1] * inv_pi) * ct / pdf, (m.kd[2] * inv_pi) * ct / pdf);
        ln.new_o = mk<R>(p.x + eps * dir.x, p.y + eps * dir.y, p.z + eps * dir.z);
        ln.new_d = dir;
        if (pdf == 0) {        // the frame ends on the INCOMING ray, which its caller re-traces (Q1): keep it across the shadow query
            ln.new_o = ln.o; ln.new_d = ln.d;
            ln.pdf_zero |= (skip_in + 1) << 2;
        }
        ln.o = so;
        ln.d = sd;
        if (sizeof(R) == 4) ln.qshift = mk<R>(eps * n.x / sd.x, eps * n.y / sd.y, eps * n.z / sd.z);      // origin x + EPSILON*n
        // light_samples.py:53: visible iff nearest t >= r - EPSILON.  The target point lies ON a lamp
        // triangle that is itself in the primitive list, so the nearest t is r up to rounding; float32
        // cannot resolve 1e-6 at r ~ 10 and needs a margin relative to r.
        ln.tlimit = sizeof(R) == 4 ? mag - fmax(eps, mag * (R)4e-6) : mag - eps;
        // float32 skips the primitive a ray starts on instead of offsetting its origin by 1e-6 (below float32 resolution).
        // That is equivalent EXCEPT for a shadow ray whose light lies behind the surface (cos_theta < 0 on the flipped
        // normal): the reference's origin sits 1e-6 in FRONT of the surface (light_samples.py:41 gets x + EPSILON*n), so the
        // ray crosses the very triangle it left at t = 1e-6/|cos_theta| > EPSILON and is occluded by it — the back side of a
        // surface is dark.  (|cos_theta| <= 1e-7 is the triangle test's own grazing rejection, intersects.py:72.)  The query is
        // still issued and counted, with an empty interval; shade_shadow takes it as occluded.
        if (sizeof(R) == 4 && cos_theta < (R)-1e-7) { ln.pdf_zero |= 2; ln.tlimit = (R)-1; }
        ln.qtype = 2;
        return ST_QUERY;
    }
    bool ok = true;
    if (m.flags & 2) {                                                  // mirror, :103-106
        V3<R> r = reflect_dir<R>(ln.d, n);
        ln.o = mk<R>(p.x + eps * n.x, p.y + eps * n.y, p.z + eps * n.z);
        ln.d = r;
        if (sizeof(R) == 4) ln.qshift = mk<R>(eps * n.x / r.x, eps * n.y / r.y, eps * n.z / r.z);
    } else if (m.transmission > (R)0) {                                 // :108-141 (Q5)
        R n1 = inside ? m.ior : (R)1, n2 = inside ? (R)1 : m.ior;
        R qq = (n1 - n2) / (n1 + n2), R0 = qq * qq;
        R theta = dot(ln.d, n);
        R om = (R)1 - (sizeof(R) == 4 ? (R)cosf((float)theta) : (R)cos((double)theta));
        R om2 = om * om;
        R reflection_prob = R0 + ((R)1 - R0) * (om2 * om2 * om);
        R Nr = m.ior;
        if (dot(ln.d, n) > 0) Nr = (R)1 / Nr;
        Nr = (R)1 / Nr;
        R cos_theta = -(dot(ln.d, n));
        R sq = (R)1 - (Nr * Nr) * ((R)1 - cos_theta * cos_theta);
        if (sq > 0 && rand_0 > reflection_prob) {
            R k = Nr * cos_theta - real_sqrt<R>(sq);
            V3<R> td = normalize(mk<R>(ln.d.x * Nr + n.x * k, ln.d.y * Nr + n.y * k, ln.d.z * Nr + n.z * k));
            ln.o = mk<R>(p.x + (-eps * n.x), p.y + (-eps * n.y), p.z + (-eps * n.z));
            ln.d = td;
            if (sizeof(R) == 4) ln.qshift = mk<R>(-eps * n.x / td.x, -eps * n.y / td.y, -eps * n.z / td.z);
        } else {
            V3<R> r = reflect_dir<R>(ln.d, n);
            ln.o = mk<R>(p.x + eps * n.x, p.y + eps * n.y, p.z + eps * n.z);
            ln.d = r;
            if (sizeof(R) == 4) ln.qshift = mk<R>(eps * n.x / r.x, eps * n.y / r.y, eps * n.z / r.z);
        }
    } else {
        ok = false;                                                     // "error" branch, :143-145
    }
    adv_rand0 = rand_0;
    return ok ? ST_ADVANCE : ST_ADVANCE_ENDED;
}

// Shading after an EXTENSION query; draws rand(path, bounce) only when something was hit.
template <typename R, int SEM, typename TR = TriRec<R>>
LT_DEVINL int shade_extension(const WalkParams<R>& P, Lane<R>& ln, int prim, R t, const TR* tris, const MatRec<R>* mats,
                              uint32_t& lidx_overflow, R& adv_rand0, HitRec<R> hit_rec = HitRec<R>()) {
    R rand_0 = (R)0, rand_1 = (R)0;
    if (prim >= 0) get_rand<R>(P, path_id_of(P, ln), ln.bounce, rand_0, rand_1);
    return shade_extension_with<R, SEM, TR>(P, ln, prim, t, tris, mats, lidx_overflow, adv_rand0, rand_0, rand_1, hit_rec);
}

// After the SHADOW query.  occluded = a hit was found in (eps, r - eps).
template <typename R, int SEM, int LF>
LT_DEVINL int shade_shadow(const WalkParams<R>& P, Lane<R>& ln, Frame<R>* frames, int64_t gtid, bool occluded,
                           uint32_t& frame_overflow, uint32_t& max_frames, R& adv_rand0) {
    if (ln.pdf_zero & 2) occluded = true;                               // float32: the surface shadows its own back side (see shade_extension_with)
    const R e_ = (R)P.eps;
    ln.qshift = mk<R>(e_, e_, e_);                                      // the pending ray starts at x + EPSILON*dir (:82): every slab distance shifts by EPSILON
    if (ln.pdf_zero & 1) { ln.o = ln.new_o; ln.d = ln.new_d; ln.skip = (ln.pdf_zero >> 2) - 1; }   // the untouched incoming ray (its own shift is not kept: ~1e-7 of the bounces)
    if (SEM == LT_SEM_ITERATIVE) {
        if (!occluded) ln.L = ln.L + ln.T * ln.contrib;                 // fix1:59
        if (ln.pdf_zero & 1) return ST_ADVANCE_ENDED;                   // fix1:64-67
        ln.T = ln.T * ln.tfac;
        ln.o = ln.new_o;
        ln.d = ln.new_d;
        adv_rand0 = ln.rand0_pending;
        return ST_ADVANCE;
    }
    if (ln.pdf_zero & 1) return ST_ADVANCE_ENDED;                       // :76-80 — direct light is dropped
    ln.T = ln.T * ln.tfac;                                              // :92
    if (!occluded) ln.L = ln.L + ln.W * ln.contrib;                     // direct enters un-weighted by T (Q2)
    // recursion (:93): suspend this frame, open the child at bounce+1 on the new ray
    frame_store<R, LF>(P, frames, gtid, ln.nframes, ln, frame_overflow);
    ln.memo += 2 * resume_retraces<R>(P, ln.T, ln.bounce, ln.rand0_pending);
    ln.nframes += 1;
    if ((uint32_t)ln.nframes + 1 > max_frames) max_frames = (uint32_t)ln.nframes + 1;
    ln.W = ln.W * ln.T;
    ln.T = mk<R>((R)1, (R)1, (R)1);
    ln.o = ln.new_o;
    ln.d = ln.new_d;
    ln.bounce += 1;
    if (ln.bounce < P.max_depth) { ln.qtype = 1; ln.tlimit = Lim<R>::inf(); return ST_QUERY; }
    return ST_ADVANCE_ENDED;                                            // child frame ends at the depth test (:41)
}

// R/src/path_tracing.py:263-286: camera ray of (pixel, sample); Q13: the same number jitters x and y.
template <typename R>
LT_DEVINL void start_path(const WalkParams<R>& P, Lane<R>& ln, int64_t pix, int s_rel) {
    int s = P.s0 + s_rel;
    int i = (int)((uint32_t)pix / (uint32_t)P.width), j = (int)((uint32_t)pix % (uint32_t)P.width);
    ln.pix = (int32_t)pix;
    ln.s_rel = s_rel;
    ln.path_id = (uint64_t)pix * (uint64_t)P.spp + (uint64_t)s;
    ln.bounce = 0; ln.nframes = 0; ln.nee = 0;
    ln.T = mk<R>((R)1, (R)1, (R)1);
    ln.W = ln.T;
    ln.L = mk<R>((R)0, (R)0, (R)0);
    R u0, u1;
    get_rand<R>(P, ln.path_id, 0, u0, u1);
    // numba's np.linspace: arr[i] = start + i*step, arr[-1] = stop
    R y = (P.height > 1 && i == P.height - 1) ? (R)P.bottom : (R)P.top + (R)i * (R)P.step_y;
    R x = (P.width > 1 && j == P.width - 1) ? (R)1 : (R)-1 + (R)j * (R)P.step_x;
    V3<R> cam = mk<R>((R)P.cam[0], (R)P.cam[1], (R)P.cam[2]);
    V3<R> end = mk<R>(x + u0 / (R)P.width, y + u0 / (R)P.height, (R)P.f_distance);
    ln.o = cam;
    ln.d = normalize(end - cam);
    ln.tlimit = Lim<R>::inf();
    ln.skip = -1;
    ln.qshift = mk<R>((R)0, (R)0, (R)0);
    ln.memo = 0;
    ln.qtype = P.max_depth > 0 ? 1 : 0;
}

}  // namespace lt
