// lt_lbvh.cu — GPU LBVH builder (replaces the host-side build_bvh/flatten_bvh of
// R/src/bvh_new.py:148-300 for NEAREST-semantics traversal).
//
//   1. per-primitive fp32 boxes, rounded OUTWARDS from the float64 vertices (+ centroid bounds)
//   2. 63-bit Morton codes of the box centroids (21 bits per axis)
//   3. LSD radix sort of (code, primitive) pairs — own 8-bit-digit histogram/scan/scatter kernels;
//      the scatter ranks keys with warp match_any so it is stable without atomics
//   4. Karras 2012 hierarchy (one thread per internal node; ties broken by position)
//   5. bottom-up refit with per-node arrival counters
//   6. emit 64-byte nodes that carry both child boxes (four 128-bit loads per visited node)
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <vector>

#include "lt_common.cuh"
#include "lt_host.h"

namespace lt {

#define LB_THREADS 256

struct Box {
    float lo[3], hi[3];
};

// ---- 1. primitive boxes ---------------------------------------------------------------------------
__global__ void prim_boxes_kernel(const double* __restrict__ v0, const double* __restrict__ v1, const double* __restrict__ v2, int n,
                                  Box* __restrict__ boxes, float* __restrict__ block_cb /*[blocks][6]*/) {
    __shared__ float red[6][LB_THREADS / 32];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float c_lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, c_hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        Box b;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double a = v0[3 * i + k], bb = v1[3 * i + k], c = v2[3 * i + k];
            b.lo[k] = __double2float_rd(fmin(a, fmin(bb, c)));
            b.hi[k] = __double2float_ru(fmax(a, fmax(bb, c)));
            float cen = 0.5f * (b.lo[k] + b.hi[k]);
            c_lo[k] = cen;
            c_hi[k] = cen;
        }
        boxes[i] = b;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            c_lo[k] = fminf(c_lo[k], __shfl_xor_sync(0xffffffffu, c_lo[k], o));
            c_hi[k] = fmaxf(c_hi[k], __shfl_xor_sync(0xffffffffu, c_hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) { red[k][threadIdx.x >> 5] = c_lo[k]; red[3 + k][threadIdx.x >> 5] = c_hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[threadIdx.x][0];
        for (int w = 1; w < LB_THREADS / 32; ++w) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
        block_cb[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void reduce_cb_kernel(const float* __restrict__ block_cb, int nblocks, float* __restrict__ cb /*[6]*/) {
    __shared__ float red[6][LB_THREADS];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], block_cb[b * 6 + k]); hi[k] = fmaxf(hi[k], block_cb[b * 6 + 3 + k]); }
    for (int k = 0; k < 3; ++k) { red[k][threadIdx.x] = lo[k]; red[3 + k][threadIdx.x] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[threadIdx.x][0];
        for (int t = 1; t < LB_THREADS; ++t) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][t]) : fmaxf(v, red[threadIdx.x][t]);
        cb[threadIdx.x] = v;
    }
}

// ---- 2. Morton codes ---------------------------------------------------------------------------------
LT_DEVINL uint64_t spread21(uint64_t x) {   // insert two zero bits after each of the low 21 bits
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void morton_kernel(const Box* __restrict__ boxes, int n, const float* __restrict__ cb, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t code = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float cen = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
        float ext = cb[3 + k] - cb[k];
        float u = ext > 0.f ? (cen - cb[k]) / ext : 0.f;
        uint64_t q = (uint64_t)fminf(fmaxf(u * 2097152.0f, 0.f), 2097151.0f);
        code |= spread21(q) << (2 - k);
    }
    keys[i] = code;
    vals[i] = (uint32_t)i;
}

// ---- 3. radix sort (8-bit digits, LSD) ---------------------------------------------------------------
#define RS_ITEMS 16                       // consecutive keys per thread-round set: tile = 256*16 = 4096 keys
#define RS_TILE (LB_THREADS * RS_ITEMS)
#define RS_WARPS (LB_THREADS / 32)

__global__ void rs_hist_kernel(const uint64_t* __restrict__ keys, int n, int shift, uint32_t* __restrict__ hist /*[256][nblocks]*/) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * RS_TILE;
    for (int k = threadIdx.x; k < RS_TILE; k += LB_THREADS) {
        int i = base + k;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan over hist viewed as one array of 256*nblocks counters (digit-major): single CTA
__global__ void rs_scan_kernel(uint32_t* __restrict__ hist, int total) {
    __shared__ uint32_t warp_sums[LB_THREADS / 32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += LB_THREADS) {
        int i = base + threadIdx.x;
        uint32_t v = i < total ? hist[i] : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_sums[w];
        uint32_t c = carry;
        if (i < total) hist[i] = c + woff + x - v;
        __syncthreads();
        if (threadIdx.x == LB_THREADS - 1) carry = c + woff + x;
        __syncthreads();
    }
}

// Stable scatter.  Warp w owns the contiguous slice [w*512, (w+1)*512) of the tile and walks it in
// rounds of 32 consecutive keys; within a round the rank of a key among equal digits comes from
// match_any, across rounds from a per-warp running count, across warps from a prefix over the
// per-warp digit totals.  Two sweeps over the slice: count, then place.
__global__ void rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int n, int shift,
                                  const uint32_t* __restrict__ hist, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t wcount[RS_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RS_WARPS * 256; k += LB_THREADS) (&wcount[0][0])[k] = 0;
    __syncthreads();
    const int slice = RS_TILE / RS_WARPS;
    const int base = blockIdx.x * RS_TILE + w * slice;
    // sweep 1: per-warp digit totals
    for (int r = 0; r < slice; r += 32) {
        int i = base + r + lane;
        bool ok = i < n;
        uint32_t dg = ok ? (uint32_t)((keys_in[i] >> shift) & 255u) : 256u + lane;   // unique pseudo-digit for padding
        unsigned peers = __match_any_sync(0xffffffffu, dg);
        if (ok && (__ffs(peers) - 1) == lane) wcount[w][dg] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps per digit, plus the global base of (digit, block)
    {
        int dg = threadIdx.x;   // LB_THREADS == 256 digits
        uint32_t run = hist[dg * gridDim.x + blockIdx.x];
        for (int ww = 0; ww < RS_WARPS; ++ww) { uint32_t c = wcount[ww][dg]; wcount[ww][dg] = run; run += c; }
    }
    __syncthreads();
    // sweep 2: place
    for (int r = 0; r < slice; r += 32) {
        int i = base + r + lane;
        bool ok = i < n;
        uint64_t key = ok ? keys_in[i] : 0ull;
        uint32_t dg = ok ? (uint32_t)((key >> shift) & 255u) : 256u + lane;
        unsigned peers = __match_any_sync(0xffffffffu, dg);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t off = 0;
        if (ok) off = wcount[w][dg] + rank;
        __syncwarp();
        if (ok && (__ffs(peers) - 1) == lane) wcount[w][dg] += __popc(peers);
        __syncwarp();
        if (ok) { keys_out[off] = key; vals_out[off] = vals_in[i]; }
    }
}

// ---- 4. Karras hierarchy -----------------------------------------------------------------------------
LT_DEVINL int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// children encoding in build arrays: >= 0 internal node, < 0 leaf at sorted position ~c
__global__ void karras_kernel(const uint64_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                              int* __restrict__ parent_internal, int* __restrict__ parent_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? ~gamma : gamma;
    int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    if (lc >= 0) parent_internal[lc] = i; else parent_leaf[~lc] = i;
    if (rc >= 0) parent_internal[rc] = i; else parent_leaf[~rc] = i;
    if (i == 0) parent_internal[0] = -1;
}

// ---- 5. refit ------------------------------------------------------------------------------------------
LT_DEVINL Box load_box_cg(const Box* p) {
    Box b;
    const float* f = reinterpret_cast<const float*>(p);
#pragma unroll
    for (int k = 0; k < 3; ++k) { b.lo[k] = __ldcg(f + k); b.hi[k] = __ldcg(f + 3 + k); }
    return b;
}
__global__ void refit_kernel(const Box* __restrict__ prim_boxes, const uint32_t* __restrict__ sorted_prim, int n, const int* __restrict__ left,
                             const int* __restrict__ right, const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf,
                             Box* __restrict__ node_boxes, int* __restrict__ arrive, int* __restrict__ max_depth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent_leaf[i];
    int depth = 1;
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&arrive[node], 1) == 0) return;   // first child to arrive stops; second proceeds
        Box b;
        int c0 = left[node], c1 = right[node];
        // sibling results were published by another SM: read them through L2 (L1 is not coherent)
        Box b0 = c0 >= 0 ? load_box_cg(node_boxes + c0) : prim_boxes[sorted_prim[~c0]];
        Box b1 = c1 >= 0 ? load_box_cg(node_boxes + c1) : prim_boxes[sorted_prim[~c1]];
#pragma unroll
        for (int k = 0; k < 3; ++k) { b.lo[k] = fminf(b0.lo[k], b1.lo[k]); b.hi[k] = fmaxf(b0.hi[k], b1.hi[k]); }
        node_boxes[node] = b;
        __threadfence();
        node = parent_internal[node];
        depth++;
    }
    (void)max_depth;
}

__global__ void depth_kernel(int n, const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf, int* __restrict__ max_depth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = 0;
    for (int node = parent_leaf[i]; node >= 0; node = parent_internal[node]) d++;
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(__activemask(), d, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, d);
}

// ---- 6. emit traversal nodes ---------------------------------------------------------------------------
__global__ void emit_kernel(const Box* __restrict__ prim_boxes, const uint32_t* __restrict__ sorted_prim, int n, const int* __restrict__ left,
                            const int* __restrict__ right, const Box* __restrict__ node_boxes, LbvhNode* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int c0 = left[i], c1 = right[i];
    Box b0 = c0 >= 0 ? node_boxes[c0] : prim_boxes[sorted_prim[~c0]];
    Box b1 = c1 >= 0 ? node_boxes[c1] : prim_boxes[sorted_prim[~c1]];
    LbvhNode nd;
    nd.x = make_float4(b0.lo[0], b1.lo[0], b0.hi[0], b1.hi[0]);
    nd.y = make_float4(b0.lo[1], b1.lo[1], b0.hi[1], b1.hi[1]);
    nd.z = make_float4(b0.lo[2], b1.lo[2], b0.hi[2], b1.hi[2]);
    nd.d = make_int4(c0 >= 0 ? c0 : ~(int)sorted_prim[~c0], c1 >= 0 ? c1 : ~(int)sorted_prim[~c1], 0, 0);
    out[i] = nd;
}

template <typename T>
struct Tmp {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return lt_dev_malloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    ~Tmp() { if (p) lt_dev_free(p); }
};

}  // namespace lt

using namespace lt;

extern "C" int lt_bvh_build(lt_scene* s, int flags, lt_bvh_stats* out) {
    (void)flags;
    if (!s) { set_error("lt_bvh_build: null scene"); return LT_E_INVALID; }
    lt_ctx* ctx = s->ctx;
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = s->n_tris;
    lt_bvh_stats stats;
    memset(&stats, 0, sizeof stats);
    stats.n_prims = n;
    if (s->lbvh_nodes) { lt_dev_free(s->lbvh_nodes); s->lbvh_nodes = nullptr; s->n_lbvh_nodes = 0; }
    if (n < 2) {   // a single primitive needs no hierarchy; the traversal tests it directly
        if (out) *out = stats;
        return LT_OK;
    }
    const int blocks = (n + LB_THREADS - 1) / LB_THREADS;
    const int tiles = (n + RS_TILE - 1) / RS_TILE;
    Tmp<Box> prim_boxes, node_boxes;
    Tmp<float> block_cb, cb;
    Tmp<uint64_t> keys_a, keys_b;
    Tmp<uint32_t> vals_a, vals_b, hist;
    Tmp<int> left, right, par_i, par_l, arrive, maxd;
    LT_CUDA_OK(prim_boxes.alloc(n)); LT_CUDA_OK(node_boxes.alloc(n));
    LT_CUDA_OK(block_cb.alloc(6 * (size_t)blocks)); LT_CUDA_OK(cb.alloc(6));
    LT_CUDA_OK(keys_a.alloc(n)); LT_CUDA_OK(keys_b.alloc(n)); LT_CUDA_OK(vals_a.alloc(n)); LT_CUDA_OK(vals_b.alloc(n));
    LT_CUDA_OK(hist.alloc(256 * (size_t)tiles));
    LT_CUDA_OK(left.alloc(n)); LT_CUDA_OK(right.alloc(n)); LT_CUDA_OK(par_i.alloc(n)); LT_CUDA_OK(par_l.alloc(n));
    LT_CUDA_OK(arrive.alloc(n)); LT_CUDA_OK(maxd.alloc(1));
    LT_CUDA_OK(cudaMemsetAsync(arrive.p, 0, (size_t)n * sizeof(int), st));
    LT_CUDA_OK(cudaMemsetAsync(maxd.p, 0, sizeof(int), st));
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    cudaEventRecord(e0, st);
    prim_boxes_kernel<<<blocks, LB_THREADS, 0, st>>>(s->d_v0, s->d_v1, s->d_v2, n, prim_boxes.p, block_cb.p);
    reduce_cb_kernel<<<1, LB_THREADS, 0, st>>>(block_cb.p, blocks, cb.p);
    morton_kernel<<<blocks, LB_THREADS, 0, st>>>(prim_boxes.p, n, cb.p, keys_a.p, vals_a.p);
    cudaEventRecord(e1, st);
    uint64_t *ki = keys_a.p, *ko = keys_b.p;
    uint32_t *vi = vals_a.p, *vo = vals_b.p;
    for (int pass = 0; pass < 8; ++pass) {           // 63-bit codes: 8 digits of 8 bits
        int shift = pass * 8;
        rs_hist_kernel<<<tiles, LB_THREADS, 0, st>>>(ki, n, shift, hist.p);
        rs_scan_kernel<<<1, LB_THREADS, 0, st>>>(hist.p, 256 * tiles);
        rs_scatter_kernel<<<tiles, LB_THREADS, 0, st>>>(ki, vi, n, shift, hist.p, ko, vo);
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    cudaEventRecord(e2, st);
    const int iblocks = (n - 1 + LB_THREADS - 1) / LB_THREADS;
    karras_kernel<<<iblocks, LB_THREADS, 0, st>>>(ki, n, left.p, right.p, par_i.p, par_l.p);
    refit_kernel<<<blocks, LB_THREADS, 0, st>>>(prim_boxes.p, vi, n, left.p, right.p, par_i.p, par_l.p, node_boxes.p, arrive.p, maxd.p);
    depth_kernel<<<blocks, LB_THREADS, 0, st>>>(n, par_i.p, par_l.p, maxd.p);
    LT_CUDA_OK(lt_dev_malloc(&s->lbvh_nodes, (size_t)(n - 1) * sizeof(LbvhNode)));
    emit_kernel<<<iblocks, LB_THREADS, 0, st>>>(prim_boxes.p, vi, n, left.p, right.p, node_boxes.p, s->lbvh_nodes);
    cudaEventRecord(e3, st);
    LT_CUDA_OK(cudaGetLastError());
    LT_CUDA_OK(cudaStreamSynchronize(st));
    float ms_all = 0, ms_sort = 0;
    cudaEventElapsedTime(&ms_all, e0, e3);
    cudaEventElapsedTime(&ms_sort, e1, e2);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    int depth = 0;
    Box root;
    LT_CUDA_OK(cudaMemcpy(&depth, maxd.p, sizeof(int), cudaMemcpyDeviceToHost));
    LT_CUDA_OK(cudaMemcpy(&root, node_boxes.p, sizeof(Box), cudaMemcpyDeviceToHost));
    s->n_lbvh_nodes = n - 1;
    stats.n_nodes = n - 1;
    stats.max_depth = depth;
    stats.build_ms = ms_all;
    stats.sort_ms = ms_sort;
    for (int k = 0; k < 3; ++k) { stats.root_bmin[k] = root.lo[k]; stats.root_bmax[k] = root.hi[k]; }
    if (out) *out = stats;
    if (depth > LT_LBVH_STACK) {
        lt_dev_free(s->lbvh_nodes); s->lbvh_nodes = nullptr; s->n_lbvh_nodes = 0;
        set_error("LBVH depth %d exceeds the traversal stack (%d)", depth, LT_LBVH_STACK);
        return LT_E_OVERFLOW;
    }
    return LT_OK;
}
