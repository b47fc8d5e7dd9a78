// lt_medium.cu — C4b: photon random walk in a homogeneous scattering medium (BASELINE.json config 4:
// "highly scattering medium / many-bounce walk (albedo 0.99, max depth 256) stressing divergence and
// Russian roulette").  The reference has no volume walk — only the Henyey-Greenstein evaluation
// R/src/medium_samples.py:14-16 — so this part is pinned by the oracle's restatement, by analytic
// expectations (geometric collision count, energy balance, <cos> = g) and by the HG function's golden
// table; see DESIGN.md.
//
// sigma_t = 1.  A photon starts at the centre of a sphere of radius `radius` mean free paths (<= 0:
// infinite medium) in an isotropic direction and repeats: free flight s = -ln(1-u); leave the sphere ->
// escaped; else collide: analog absorption (mode 0) or implicit capture followed by the reference's
// roulette rule on the weight (mode 1: q = max(rr_floor, 1-w), R/src/path_tracing.py:148-155); stop after
// max_depth collisions (truncated); scatter by Henyey-Greenstein.
//
// Scheduling = the walk megakernel's: persistent warps, lanes regenerate from a warp-uniform cursor
// (ballot + popc ranks).  TALLIES are staged in shared memory per CTA — the collision-count histogram
// and the radial absorbed-weight bins — through warp-aggregated atomics (match_any groups the lanes
// that hit the same bin; one lane adds the group's sum), and flushed once per CTA with global atomics.
#include <algorithm>
#include <cstring>

#include "lt_common.cuh"
#include "lt_host.h"

namespace lt {

#define LM_BLOCK 128

struct MediumDev {
    double albedo, g, radius, rr_floor;
    int32_t max_depth, mode, rr_start, n_rbins;
    uint64_t seed, first_photon, n_photons;
    int64_t photons_per_warp;
    unsigned long long* hist;      // [max_depth+1]
    double* rbins;                 // [n_rbins]
    double* totals;                // escaped, absorbed, truncated, rr_killed, rr_gained, sum_mu
    unsigned long long* counts;    // n_collisions, n_segments, n_scatters
};

template <typename R>
LT_DEVINL void medium_rand4(const MediumDev& P, uint64_t photon, uint32_t k, R u[4]);
template <>
LT_DEVINL void medium_rand4<double>(const MediumDev& P, uint64_t photon, uint32_t k, double u[4]) {
    uint32_t o[4];
    philox4x32_10((uint32_t)photon, (uint32_t)(photon >> 32), k, 2u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), o);
    u01(o[0], o[1], u[0]); u01(o[2], o[3], u[1]);
    philox4x32_10((uint32_t)photon, (uint32_t)(photon >> 32), k, 3u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), o);
    u01(o[0], o[1], u[2]); u01(o[2], o[3], u[3]);
}
template <>
LT_DEVINL void medium_rand4<float>(const MediumDev& P, uint64_t photon, uint32_t k, float u[4]) {
    uint32_t o[4];   // one block is enough for four 24-bit uniforms (top bits of the same words the f64 walk uses first)
    philox4x32_10((uint32_t)photon, (uint32_t)(photon >> 32), k, 2u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), o);
    u01(o[0], 0u, u[0]); u01(o[2], 0u, u[1]);
    u01(o[1], 0u, u[2]); u01(o[3], 0u, u[3]);
}

// Tallies go to shared memory with native 32-bit atomics (see the kernel).  Round 1 summed the lanes of a match_any group first
// and issued one atomic per distinct bin — needed while the radial bins were doubles (CAS loops); with fixed-point bins that
// loop (15 % of the kernel's instructions) is gone.
#define LM_FIXED_ONE 4294967296.0

template <typename R>
__global__ void __launch_bounds__(LM_BLOCK) medium_kernel(const MediumDev P) {
    extern __shared__ __align__(16) char smem[];
    // tallies of the CTA in shared memory, 32-bit words: 64-bit shared-memory atomic adds are CAS loops on this part (measured: 9 %
    // of the instructions, 24 % of the stall samples), 32-bit ones are native.  Collision-count histogram: 32-bit counts (the host
    // keeps a CTA below 2^31 collisions per launch).  Radial bins: 32.32 fixed point — a deposit is a weight in [0, 1] — held as a
    // low and a high word: the low word is added with the returning atomic, whose old value tells this very add whether it wrapped,
    // and the wrap goes to the high word; integer sums are exact to 2^-33 per deposit and order-independent.
    unsigned* sh_hist = reinterpret_cast<unsigned*>(smem);                                   // [D+1]
    unsigned* sh_rlo = sh_hist + ((P.max_depth + 2) & ~1);                                   // [NB]
    unsigned* sh_rhi = sh_rlo + P.n_rbins;                                                   // [NB]
    for (int k = threadIdx.x; k <= P.max_depth; k += blockDim.x) sh_hist[k] = 0u;
    for (int k = threadIdx.x; k < P.n_rbins; k += blockDim.x) { sh_rlo[k] = 0u; sh_rhi[k] = 0u; }
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const R a = (R)P.albedo, g = (R)P.g, Rs = (R)P.radius;
    const int D = P.max_depth, NB = P.n_rbins;
    const R two_pi = (R)6.283185307179586;

    R x[3], d[3], w = 0;
    int k = 0;
    uint64_t id = 0;
    bool alive = false;
    double t_esc = 0, t_abs = 0, t_trunc = 0, t_rrk = 0, t_rrg = 0, t_mu = 0;
    unsigned long long n_coll = 0, n_seg = 0, n_scat = 0;

    for (int64_t gq = gwarp; gq * P.photons_per_warp < (int64_t)P.n_photons; gq += nwarps) {
        int64_t next = gq * P.photons_per_warp;
        const int64_t end = min((int64_t)P.n_photons, next + P.photons_per_warp);
        for (;;) {
            // ---- regeneration (warp-aggregated fetch from the warp-uniform cursor)
            unsigned m = __ballot_sync(full, !alive);
            if (m) {
                int64_t mine = next + __popc(m & ((1u << lane) - 1u));
                next += __popc(m);
                if (!alive && mine < end) {
                    id = P.first_photon + (uint64_t)mine;
                    R u[4];
                    medium_rand4<R>(P, id, 0xffffffffu, u);
                    R mu = (R)1 - (R)2 * u[0], s2 = (R)1 - mu * mu, sn = real_sqrt<R>(s2 > 0 ? s2 : (R)0), ph = two_pi * u[1];
                    R sp, cp;
                    if (sizeof(R) == 4) { float s_, c_; sincosf((float)ph, &s_, &c_); sp = s_; cp = c_; } else { double s_, c_; sincos((double)ph, &s_, &c_); sp = (R)s_; cp = (R)c_; }
                    d[0] = sn * cp; d[1] = sn * sp; d[2] = mu;
                    x[0] = x[1] = x[2] = 0;
                    w = 1; k = 0; alive = true;
                }
            }
            if (!__any_sync(full, alive)) break;
            // ---- one flight segment + collision for every live lane
            bool fin = false;
            bool deposit = false;
            double dep = 0.0;
            int shell = 0;
            if (alive) {
                R u[4];
                medium_rand4<R>(P, id, (uint32_t)k, u);
                R s = sizeof(R) == 4 ? (R)-logf((float)((R)1 - u[0])) : (R)-log((double)((R)1 - u[0]));
                n_seg++;
                x[0] += s * d[0]; x[1] += s * d[1]; x[2] += s * d[2];
                R r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
                if (Rs > 0 && r2 >= Rs * Rs) { t_esc += (double)w; fin = true; }
                else {
                    k++;
                    n_coll++;
                    shell = NB > 0 ? (Rs > 0 ? (int)(real_sqrt<R>(r2) / Rs * (R)NB) : 0) : 0;
                    if (shell >= NB) shell = NB - 1;
                    if (P.mode == 0) {
                        if (u[1] >= a) { t_abs += (double)w; dep = (double)w; deposit = NB > 0; fin = true; }
                    } else {
                        R dw = w * ((R)1 - a);
                        t_abs += (double)dw; dep = (double)dw; deposit = NB > 0;
                        w *= a;
                        if (k - 1 > P.rr_start) {
                            R om = (R)1 - w, fl = (R)P.rr_floor, q = om > fl ? om : fl;
                            if (u[1] < q) { t_rrk += (double)w; fin = true; }
                            else { t_rrg += (double)(w / ((R)1 - q) - w); w /= (R)1 - q; }
                        }
                    }
                    if (!fin && k >= D) { t_trunc += (double)w; fin = true; }
                    if (!fin) {
                        R mu;
                        if (rabs(g) < (R)1e-3) mu = (R)1 - (R)2 * u[2];
                        else { R sq = ((R)1 - g * g) / ((R)1 - g + (R)2 * g * u[2]); mu = ((R)1 + g * g - sq * sq) / ((R)2 * g); }
                        mu = mu > (R)1 ? (R)1 : (mu < (R)-1 ? (R)-1 : mu);
                        t_mu += (double)mu;
                        n_scat++;
                        R sn = real_sqrt<R>((R)1 - mu * mu), ph = two_pi * u[3], sp, cp;
                        if (sizeof(R) == 4) { float s_, c_; sincosf((float)ph, &s_, &c_); sp = s_; cp = c_; } else { double s_, c_; sincos((double)ph, &s_, &c_); sp = (R)s_; cp = (R)c_; }
                        R nd0, nd1, nd2;
                        if (rabs(d[2]) > (R)0.99999) { nd0 = sn * cp; nd1 = sn * sp; nd2 = d[2] > 0 ? mu : -mu; }
                        else {
                            R den = real_sqrt<R>((R)1 - d[2] * d[2]);
                            nd0 = sn * (d[0] * d[2] * cp - d[1] * sp) / den + d[0] * mu;
                            nd1 = sn * (d[1] * d[2] * cp + d[0] * sp) / den + d[1] * mu;
                            nd2 = -sn * cp * den + d[2] * mu;
                        }
                        R nn = real_sqrt<R>(nd0 * nd0 + nd1 * nd1 + nd2 * nd2);
                        d[0] = nd0 / nn; d[1] = nd1 / nn; d[2] = nd2 / nn;
                    }
                }
            }
            // ---- tallies: shared-memory staging, warp-aggregated atomics
            if (deposit) {
                const unsigned long long fx = __double2ull_rn(dep * LM_FIXED_ONE);
                const unsigned xl = (unsigned)fx, xh = (unsigned)(fx >> 32);
                const unsigned old = atomicAdd(sh_rlo + shell, xl);
                const unsigned up = xh + ((old + xl) < old ? 1u : 0u);
                if (up) atomicAdd(sh_rhi + shell, up);
            }
            if (fin) atomicAdd(sh_hist + min(k, D), 1u);
            if (fin) alive = false;
        }
    }
    __syncthreads();
    for (int kk = threadIdx.x; kk <= D; kk += blockDim.x) if (sh_hist[kk]) atomicAdd(P.hist + kk, (unsigned long long)sh_hist[kk]);
    for (int kk = threadIdx.x; kk < NB; kk += blockDim.x) {
        const unsigned long long v = ((unsigned long long)sh_rhi[kk] << 32) | sh_rlo[kk];
        if (v) atomicAdd(P.rbins + kk, (double)v * (1.0 / LM_FIXED_ONE));
    }
    double tv[6] = {t_esc, t_abs, t_trunc, t_rrk, t_rrg, t_mu};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = tv[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
        if (lane == 0 && v != 0.0) atomicAdd(P.totals + i, v);
    }
    unsigned long long cv[3] = {n_coll, n_seg, n_scat};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        unsigned long long v = cv[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
        if (lane == 0 && v) atomicAdd(P.counts + i, v);
    }
}

}  // namespace lt

using namespace lt;

extern "C" int lt_medium_walk(lt_ctx* ctx, const lt_medium_params* p, uint64_t n_photons, uint64_t* hist, double* rbins,
                              lt_medium_totals* totals, float* kernel_ms) {
    if (!ctx || !p || !hist || !totals) { set_error("lt_medium_walk: null argument"); return LT_E_INVALID; }
    if (p->max_depth <= 0 || p->max_depth > 4096 || p->n_rbins < 0 || p->n_rbins > 1024 || !(p->albedo >= 0 && p->albedo <= 1) ||
        !(p->g > -1 && p->g < 1) || (p->mode != 0 && p->mode != 1) || (p->real != LT_REAL_F32 && p->real != LT_REAL_F64)) {
        set_error("lt_medium_walk: bad parameters");
        return LT_E_INVALID;
    }
    if (p->n_rbins > 0 && !rbins) { set_error("lt_medium_walk: rbins is NULL"); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int D = p->max_depth, NB = p->n_rbins;
    DevBuf<unsigned long long> b_hist, b_cnt;        // released on every return path
    DevBuf<double> b_rbins, b_tot;
    LT_CUDA_OK(b_hist.alloc(D + 1));
    LT_CUDA_OK(b_rbins.alloc(std::max(NB, 1)));
    LT_CUDA_OK(b_tot.alloc(6));
    LT_CUDA_OK(b_cnt.alloc(3));
    unsigned long long *d_hist = b_hist.p, *d_cnt = b_cnt.p;
    double *d_rbins = b_rbins.p, *d_tot = b_tot.p;
    LT_CUDA_OK(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * (D + 1), st));
    LT_CUDA_OK(cudaMemsetAsync(d_rbins, 0, sizeof(double) * std::max(NB, 1), st));
    LT_CUDA_OK(cudaMemsetAsync(d_tot, 0, sizeof(double) * 6, st));
    LT_CUDA_OK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 3, st));
    MediumDev P;
    P.albedo = p->albedo; P.g = p->g; P.radius = p->radius; P.rr_floor = p->rr_floor;
    P.max_depth = D; P.mode = p->mode; P.rr_start = p->rr_start; P.n_rbins = NB;
    P.seed = p->seed; P.first_photon = p->first_photon; P.n_photons = n_photons;
    P.hist = d_hist; P.rbins = d_rbins; P.totals = d_tot; P.counts = d_cnt;
    const size_t smem = sizeof(unsigned) * ((D + 2) & ~1) + 2 * sizeof(unsigned) * std::max(NB, 1);
    int bps = 0;
    if (p->real == LT_REAL_F64) LT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, medium_kernel<double>, LM_BLOCK, smem));
    else LT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, medium_kernel<float>, LM_BLOCK, smem));
    int grid = ctx->prop.multiProcessorCount * std::max(bps, 1);
    int64_t nwarps = (int64_t)grid * LM_BLOCK / 32;
    int64_t ppw = std::max<int64_t>(64, std::min<int64_t>((int64_t)n_photons / (nwarps * 6), 8192));
    ppw = (ppw + 31) / 32 * 32;
    int64_t granules = ((int64_t)n_photons + ppw - 1) / ppw;
    if (granules < nwarps) grid = (int)std::max<int64_t>(1, (granules * 32 + LM_BLOCK - 1) / LM_BLOCK);
    P.photons_per_warp = ppw;
    LT_CUDA_OK(cudaEventRecord(ctx->ev0, st));
    // a CTA's fixed-point radial bins hold 2^32 unit deposits: at most 2^31 collisions per CTA and launch
    const uint64_t per_launch = std::max<uint64_t>(1, ((1ull << 31) / (uint64_t)D)) * (uint64_t)grid;
    for (uint64_t done = 0; done < n_photons; done += per_launch) {
        P.first_photon = p->first_photon + done;
        P.n_photons = std::min<uint64_t>(per_launch, n_photons - done);
        if (p->real == LT_REAL_F64) medium_kernel<double><<<grid, LM_BLOCK, smem, st>>>(P);
        else medium_kernel<float><<<grid, LM_BLOCK, smem, st>>>(P);
    }
    LT_CUDA_OK(cudaGetLastError());
    LT_CUDA_OK(cudaEventRecord(ctx->ev1, st));
    std::vector<unsigned long long> h_hist(D + 1);
    double h_tot[6];
    unsigned long long h_cnt[3];
    LT_CUDA_OK(cudaMemcpyAsync(h_hist.data(), d_hist, sizeof(unsigned long long) * (D + 1), cudaMemcpyDeviceToHost, st));
    if (NB > 0) LT_CUDA_OK(cudaMemcpyAsync(rbins, d_rbins, sizeof(double) * NB, cudaMemcpyDeviceToHost, st));
    LT_CUDA_OK(cudaMemcpyAsync(h_tot, d_tot, sizeof h_tot, cudaMemcpyDeviceToHost, st));
    LT_CUDA_OK(cudaMemcpyAsync(h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
    LT_CUDA_OK(cudaStreamSynchronize(st));
    for (int k = 0; k <= D; ++k) hist[k] = h_hist[k];
    totals->escaped = h_tot[0]; totals->absorbed = h_tot[1]; totals->truncated = h_tot[2]; totals->rr_killed = h_tot[3];
    totals->rr_gained = h_tot[4]; totals->sum_mu = h_tot[5];
    totals->n_collisions = h_cnt[0]; totals->n_segments = h_cnt[1]; totals->n_scatters = h_cnt[2];
    if (kernel_ms) LT_CUDA_OK(cudaEventElapsedTime(kernel_ms, ctx->ev0, ctx->ev1));
    return LT_OK;
}

// R/src/medium_samples.py:14-16
extern "C" double lt_henyey_greenstein(double cosTheta, double g) {
    const double inv_4_pi = 0.25 * (1.0 / 3.141592653589793);
    double denom = 1 + g * g + 2 * g * cosTheta;
    return inv_4_pi * (1 - g * g) / (denom * sqrt(denom));
}
