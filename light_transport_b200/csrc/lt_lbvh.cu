// lt_lbvh.cu — GPU LBVH builder (replaces the host-side build_bvh/flatten_bvh of
// R/src/bvh_new.py:148-300 for NEAREST-semantics traversal).
//
//   1. per-primitive fp32 boxes, rounded OUTWARDS from the float64 vertices (+ centroid bounds)
//   2. 63-bit Morton codes of the box centroids (21 bits per axis)
//   3. LSD radix sort of (code, primitive) pairs — own 8-bit-digit histogram/scan/scatter kernels;
//      the scatter ranks keys with warp match_any so it is stable without atomics
//   4. Karras 2012 hierarchy (one thread per internal node; ties broken by position)
//   5. bottom-up refit with per-node arrival counters
//   6. collapse the binary tree top-down, level by level, into an 8-wide hierarchy: a wide node starts from the two
//      children of its binary node and keeps opening the child with the largest surface area until it holds eight
//      (subtrees of <= 3 triangles become leaf slots); children are assigned to slots by direction octant so that a
//      traversal visits them front to back with bit arithmetic only; child boxes are quantised to 8 bits per plane on the
//      node's own grid -> 80-byte nodes (WNode) in breadth-first order, 48-byte float32 triangle records (WTri) in LEAF
//      order, and a leaf-order -> primitive index for the float64 traversal.
// Primitives that are large against the scene (extent > 1/16 of the scene's) get the top key bit: the Karras root then
// splits the large ones from the rest, and a handful of wall-sized triangles no longer inflate the boxes of every
// ancestor they would otherwise share with a fine mesh.
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

#define LB_BIG_FRACTION (1.0f / 16.0f)
__global__ void morton_kernel(const Box* __restrict__ boxes, int n, const float* __restrict__ cb, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t code = 0;
    const float scene_ext = fmaxf(fmaxf(cb[3] - cb[0], cb[4] - cb[1]), cb[5] - cb[2]);
    float prim_ext = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float cen = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
        float ext = cb[3 + k] - cb[k];
        float u = ext > 0.f ? (cen - cb[k]) / ext : 0.f;
        uint64_t q = (uint64_t)fminf(fmaxf(u * 2097152.0f, 0.f), 2097151.0f);
        code |= spread21(q) << (2 - k);
        prim_ext = fmaxf(prim_ext, boxes[i].hi[k] - boxes[i].lo[k]);
    }
    if (prim_ext > LB_BIG_FRACTION * scene_ext) code |= 1ull << 63;     // large primitives: their own subtree under the root
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
                              int* __restrict__ parent_internal, int* __restrict__ parent_leaf, int* __restrict__ rng_lo,
                              int* __restrict__ rng_hi) {
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
    rng_lo[i] = lo;          // the node covers sorted positions [lo, hi]
    rng_hi[i] = hi;
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
                             Box* __restrict__ node_boxes, int* __restrict__ arrive) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent_leaf[i];
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
    }
}

__global__ void depth_kernel(int n, const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf, int* __restrict__ max_depth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int d = 0;                                        // lanes past the end take part in the reduction with 0
    if (i < n)
        for (int node = parent_leaf[i]; node >= 0; node = parent_internal[node]) d++;
    d = __reduce_max_sync(0xffffffffu, d);
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, d);
}

// ---- 6a. binary traversal nodes (LT_TRAV_LBVH2, flag LT_BVH_KEEP_BINARY) -----------------------------------------------
// leaf refs point into the LEAF-ordered triangle arrays the collapse has produced (wpos: primitive -> position)
// The binary tree as traversed (LT_TRAV_LBVH2): only Karras nodes with more than `leaf_max` triangles below them become nodes,
// renumbered densely (`newidx` = exclusive scan of the used flags: the node array holds nothing that is never fetched);
// a smaller subtree is ONE leaf reference ~(first position | (count-1) << 28) into the leaf-ordered triangle array, where
// the 8-wide emit laid its triangles out contiguously (the same subtree is one leaf slot there).
LT_DEVINL bool binary_used(int i, const int* rng_lo, const int* rng_hi, int leaf_max) { return i == 0 || rng_hi[i] - rng_lo[i] + 1 > leaf_max; }

__global__ void mark_used_kernel(int n, const int* __restrict__ rng_lo, const int* __restrict__ rng_hi, int leaf_max, uint32_t* __restrict__ used) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) used[i] = binary_used(i, rng_lo, rng_hi, leaf_max) ? 1u : 0u;
}

__global__ void emit_kernel(const Box* __restrict__ prim_boxes, const uint32_t* __restrict__ sorted_prim, int n, const int* __restrict__ left,
                            const int* __restrict__ right, const int* __restrict__ rng_lo, const int* __restrict__ rng_hi, int leaf_max,
                            const uint32_t* __restrict__ newidx, const Box* __restrict__ node_boxes, const int32_t* __restrict__ wpos,
                            LbvhNode* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !binary_used(i, rng_lo, rng_hi, leaf_max)) return;
    int c0 = left[i], c1 = right[i];
    Box b0 = c0 >= 0 ? node_boxes[c0] : prim_boxes[sorted_prim[~c0]];
    Box b1 = c1 >= 0 ? node_boxes[c1] : prim_boxes[sorted_prim[~c1]];
    auto ref = [&](int c) -> int {
        if (c < 0) return ~wpos[sorted_prim[~c]];
        const int cnt = rng_hi[c] - rng_lo[c] + 1;
        if (cnt > leaf_max) return (int)newidx[c];
        return ~(wpos[sorted_prim[rng_lo[c]]] | ((cnt - 1) << 28));
    };
    LbvhNode nd;
    nd.x = make_float4(b0.lo[0], b1.lo[0], b0.hi[0], b1.hi[0]);
    nd.y = make_float4(b0.lo[1], b1.lo[1], b0.hi[1], b1.hi[1]);
    nd.z = make_float4(b0.lo[2], b1.lo[2], b0.hi[2], b1.hi[2]);
    nd.d = make_int4(ref(c0), ref(c1), 0, 0);
    out[newidx[i]] = nd;
}

// ---- 6b. collapse into the 8-wide hierarchy -----------------------------------------------------------------------
struct TreeView {
    const int *left, *right, *rng_lo, *rng_hi;
    const Box *node_boxes, *prim_boxes;
    const uint32_t* sorted_prim;
    int leaf_max;              // triangles per leaf slot (1..LT_WBVH_LEAF_MAX)
};
#define WB_EMPTY ((int)0x80000000)
#define LB_QUANT_SLACK 0.05      // quanta added to every quantised plane: covers the float32 decode of the traversal (wnode_step)
LT_DEVINL Box ref_box(const TreeView& t, int c) { return c >= 0 ? t.node_boxes[c] : t.prim_boxes[t.sorted_prim[~c]]; }
LT_DEVINL int ref_count(const TreeView& t, int c) { return c >= 0 ? t.rng_hi[c] - t.rng_lo[c] + 1 : 1; }
LT_DEVINL bool ref_is_leaf(const TreeView& t, int c) { return ref_count(t, c) <= t.leaf_max; }
__host__ __device__ inline float box_area(const Box& b) {
    const float x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
    return 2.0f * (x * y + y * z + z * x);
}

// One thread per wide node of the current level: choose its (up to) eight children, count what it will allocate.
__global__ void collapse_kernel(TreeView t, const int* __restrict__ wb, int level_start, int level_n, int* __restrict__ slots,
                                uint32_t* __restrict__ cnt_int, uint32_t* __restrict__ cnt_tri) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= level_n) return;
    const int w = level_start + k, b = wb[w];
    int c[8];
    float a[8];
    int n = 2;
    c[0] = t.left[b]; c[1] = t.right[b];
    for (int i = 0; i < 2; ++i) a[i] = ref_is_leaf(t, c[i]) ? -1.0f : box_area(ref_box(t, c[i]));
    while (n < 8) {
        int best = -1;
        float ba = -1.0f;
        for (int i = 0; i < n; ++i)
            if (a[i] > ba) { ba = a[i]; best = i; }
        if (best < 0) break;                               // only leaf slots left
        const int x = c[best];
        c[best] = t.left[x];
        c[n] = t.right[x];
        a[best] = ref_is_leaf(t, c[best]) ? -1.0f : box_area(ref_box(t, c[best]));
        a[n] = ref_is_leaf(t, c[n]) ? -1.0f : box_area(ref_box(t, c[n]));
        ++n;
    }
    uint32_t ni = 0, nt = 0;
    for (int i = 0; i < 8; ++i) {
        slots[8 * (size_t)w + i] = i < n ? c[i] : WB_EMPTY;
        if (i < n) {
            if (ref_is_leaf(t, c[i])) nt += (uint32_t)ref_count(t, c[i]);
            else ni += 1;
        }
    }
    cnt_int[k] = ni;
    cnt_tri[k] = nt;
}

// exclusive scan of two counter arrays in place (single CTA), totals[0..1] = their sums
__global__ void scan2_kernel(uint32_t* __restrict__ a, uint32_t* __restrict__ b, int total, uint32_t* __restrict__ totals) {
    __shared__ uint32_t wsum[2][32];
    __shared__ uint32_t carry[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (int base = 0; base < total; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const uint32_t va = i < total ? a[i] : 0u, vb = i < total ? b[i] : 0u;
        uint32_t xa = va, xb = vb;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
            if (lane >= o) { xa += ya; xb += yb; }
        }
        if (lane == 31) { wsum[0][wid] = xa; wsum[1][wid] = xb; }
        __syncthreads();
        uint32_t oa = 0, ob = 0;
        for (int w = 0; w < wid; ++w) { oa += wsum[0][w]; ob += wsum[1][w]; }
        const uint32_t ca = carry[0], cb2 = carry[1];
        if (i < total) { a[i] = ca + oa + xa - va; b[i] = cb2 + ob + xb - vb; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { carry[0] = ca + oa + xa; carry[1] = cb2 + ob + xb; }
        __syncthreads();
    }
    (void)nw;
    if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; }
}

// ---- the same scan over many CTAs: tile sums -> scan of the sums (the kernel above) -> per-tile scan with the tile's offset.
// `b` may be null (one array).  Used above SC_MIN elements; below that one CTA is faster than three launches.
#define SC_THREADS 256
#define SC_ITEMS 16
#define SC_TILE (SC_THREADS * SC_ITEMS)
#define SC_MIN 32768
__global__ void scan_tile_sums_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int total, uint32_t* __restrict__ sums, int nblk) {
    __shared__ uint32_t ws[2][SC_THREADS / 32];
    const int base = blockIdx.x * SC_TILE;
    uint32_t sa = 0, sb = 0;
    for (int k = threadIdx.x; k < SC_TILE; k += SC_THREADS) {
        const int i = base + k;
        if (i < total) { sa += a[i]; if (b) sb += b[i]; }
    }
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = sa; ws[1][threadIdx.x >> 5] = sb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t ta = 0, tb = 0;
        for (int w = 0; w < SC_THREADS / 32; ++w) { ta += ws[0][w]; tb += ws[1][w]; }
        sums[blockIdx.x] = ta;
        sums[nblk + blockIdx.x] = tb;
    }
}

__global__ void scan_tile_apply_kernel(uint32_t* __restrict__ a, uint32_t* __restrict__ b, int total, const uint32_t* __restrict__ sums, int nblk) {
    __shared__ uint32_t ws[2][SC_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;          // each thread owns SC_ITEMS consecutive elements
    uint32_t va[SC_ITEMS], vb[SC_ITEMS], sa = 0, sb = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int i = base + k;
        va[k] = i < total ? a[i] : 0u;
        vb[k] = (b && i < total) ? b[i] : 0u;
        sa += va[k]; sb += vb[k];
    }
    uint32_t xa = sa, xb = sb;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += ya; xb += yb; }
    }
    if (lane == 31) { ws[0][wid] = xa; ws[1][wid] = xb; }
    __syncthreads();
    uint32_t ra = sums[blockIdx.x] + xa - sa, rb = sums[nblk + blockIdx.x] + xb - sb;
    for (int w = 0; w < wid; ++w) { ra += ws[0][w]; rb += ws[1][w]; }
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int i = base + k;
        if (i < total) { a[i] = ra; ra += va[k]; if (b) { b[i] = rb; rb += vb[k]; } }
    }
}

// exclusive scan in place of a[0..total) (and b, if given) for total > SC_MIN; totals[0..1] = the sums.
// `sums`: 2 * ceil(total / SC_TILE) words of scratch.
static void scan_multi(uint32_t* a, uint32_t* b, int total, uint32_t* totals, uint32_t* sums, cudaStream_t st) {
    const int nblk = (total + SC_TILE - 1) / SC_TILE;
    scan_tile_sums_kernel<<<nblk, SC_THREADS, 0, st>>>(a, b, total, sums, nblk);
    scan2_kernel<<<1, 1024, 0, st>>>(sums, sums + nblk, nblk, totals);
    scan_tile_apply_kernel<<<nblk, SC_THREADS, 0, st>>>(a, b, total, sums, nblk);
}

struct RawTris {
    const double *v0, *v1, *v2;
    const int32_t* mat;
    const uint8_t* light;
};

LT_DEVINL void write_wtri(const RawTris& r, int prim, WTri* __restrict__ out) {
    WTri t;
    double cx[3], e1[3], e2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double a = r.v0[3 * (size_t)prim + c], b = r.v1[3 * (size_t)prim + c], cc = r.v2[3 * (size_t)prim + c];
        t.v0[c] = (float)a;
        e1[c] = b - a; e2[c] = cc - a;
        t.ab[c] = (float)e1[c];                  // float64 differences rounded afterwards, like TriRec<float>
        t.ac[c] = (float)e2[c];
    }
    cx[0] = e1[1] * e2[2] - e1[2] * e2[1]; cx[1] = e1[2] * e2[0] - e1[0] * e2[2]; cx[2] = e1[0] * e2[1] - e1[1] * e2[0];
    t.area2 = (float)sqrt(cx[0] * cx[0] + cx[1] * cx[1] + cx[2] * cx[2]);
    t.prim = prim;
    t.mat_light = r.mat[prim] | ((r.light && r.light[prim]) ? (int)0x80000000 : 0);
    *out = t;
}

// One thread per wide node of the current level: octant slot assignment, quantisation, node + triangle records,
// and the binary nodes of the next level's wide nodes.
__global__ void emit_wide_kernel(TreeView t, RawTris raw, int* __restrict__ wb, int level_start, int level_n, int next_start, int tri_start,
                                 const int* __restrict__ slots, const uint32_t* __restrict__ off_int, const uint32_t* __restrict__ off_tri,
                                 WNode* __restrict__ out, WTri* __restrict__ wtris, int32_t* __restrict__ leaf_prims, int32_t* __restrict__ wpos,
                                 double* __restrict__ sah) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    double cost = 0.0;
    if (k < level_n) {
        const int w = level_start + k, b = wb[w];
        const Box nb = t.node_boxes[b];
        int c[8];
        Box bx[8];
        int n = 0;
        for (int i = 0; i < 8; ++i) {
            c[i] = slots[8 * (size_t)w + i];
            if (c[i] != WB_EMPTY) { bx[i] = ref_box(t, c[i]); n = i + 1; }
        }
        // ---- slot assignment: child i goes to the free slot s whose diagonal (+-1, +-1, +-1) it lies furthest along, greedily
        // (a ray of octant `oct` then visits slots in descending s ^ (7 ^ oct): near children first)
        float ctr[3];
        for (int a = 0; a < 3; ++a) ctr[a] = 0.5f * (nb.lo[a] + nb.hi[a]);
        int child_at[8];
        for (int s2 = 0; s2 < 8; ++s2) child_at[s2] = -1;
        unsigned done_c = 0, used_s = 0;
        for (int r = 0; r < n; ++r) {
            float bestv = -FLT_MAX;
            int bi = 0, bs = 0;
            for (int i = 0; i < n; ++i) {
                if (done_c & (1u << i)) continue;
                const float dx = 0.5f * (bx[i].lo[0] + bx[i].hi[0]) - ctr[0], dy = 0.5f * (bx[i].lo[1] + bx[i].hi[1]) - ctr[1],
                            dz = 0.5f * (bx[i].lo[2] + bx[i].hi[2]) - ctr[2];
                for (int s2 = 0; s2 < 8; ++s2) {
                    if (used_s & (1u << s2)) continue;
                    const float v = ((s2 & 1) ? dx : -dx) + ((s2 & 2) ? dy : -dy) + ((s2 & 4) ? dz : -dz);
                    if (v > bestv) { bestv = v; bi = i; bs = s2; }
                }
            }
            child_at[bs] = bi;
            done_c |= 1u << bi;
            used_s |= 1u << bs;
        }
        // ---- quantisation grid: plane = p + q * 2^e, q in 0..255, boxes only ever grow (see lt_scene.cuh)
        float pf[3], scale[3];
        uint32_t eb[3];
        for (int a = 0; a < 3; ++a) {
            const float ext = nb.hi[a] - nb.lo[a];
            const float mag = fmaxf(fabsf(nb.lo[a]), fabsf(nb.hi[a]));
            const float ext_eff = fmaxf(fmaxf(ext, mag * (1.0f / 8192.0f)), 1e-30f);
            int ex;
            frexpf(ext_eff / 253.0f, &ex);                       // 2^ex >= ext_eff / 253
            int biased = min(max(ex + 127, 2), 250);
            for (;;) {
                scale[a] = __uint_as_float((uint32_t)biased << 23);
                pf[a] = __fsub_rd(nb.lo[a], scale[a]);
                const double top = ((double)nb.hi[a] - (double)pf[a]) / (double)scale[a] + LB_QUANT_SLACK;
                if (top <= 255.0 || biased >= 250) break;
                ++biased;
            }
            eb[a] = (uint32_t)biased;
        }
        uint32_t imask = 0, meta[8], q[3][2][8];
        const uint32_t child_base = (uint32_t)next_start + off_int[k], tri_base = (uint32_t)tri_start + off_tri[k];
        uint32_t n_int = 0, n_tri = 0;
        for (int s2 = 0; s2 < 8; ++s2) {
            meta[s2] = 0;
            for (int a = 0; a < 3; ++a) { q[a][0][s2] = 255; q[a][1][s2] = 0; }
            const int i = child_at[s2];
            if (i < 0) continue;
            for (int a = 0; a < 3; ++a) {
                const double xl = ((double)bx[i].lo[a] - (double)pf[a]) / (double)scale[a] - LB_QUANT_SLACK;
                const double xh = ((double)bx[i].hi[a] - (double)pf[a]) / (double)scale[a] + LB_QUANT_SLACK;
                q[a][0][s2] = (uint32_t)min(max((int)floor(xl), 0), 255);
                q[a][1][s2] = (uint32_t)min(max((int)ceil(xh), 0), 255);
            }
            const int ref = c[i];
            const int cnt = ref_count(t, ref);
            if (cnt <= t.leaf_max) {
                meta[s2] = (((1u << cnt) - 1u) << 5) | n_tri;
                const int first = ref >= 0 ? t.rng_lo[ref] : ~ref;
                for (int j = 0; j < cnt; ++j) {
                    const int prim = (int)t.sorted_prim[first + j];
                    write_wtri(raw, prim, wtris + tri_base + n_tri + j);
                    leaf_prims[tri_base + n_tri + j] = prim;
                    wpos[prim] = (int32_t)(tri_base + n_tri + j);
                }
                n_tri += (uint32_t)cnt;
                cost += (double)box_area(bx[i]) * cnt;
            } else {
                meta[s2] = 0x20u | (24u + (uint32_t)s2);
                imask |= 1u << s2;
                wb[child_base + n_int] = ref;
                n_int += 1;
            }
        }
        cost += (double)box_area(nb);
        WNode nd;
        nd.w0 = make_uint4(__float_as_uint(pf[0]), __float_as_uint(pf[1]), __float_as_uint(pf[2]), eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24));
        auto pack4 = [](const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
        nd.w1 = make_uint4(child_base, tri_base, pack4(meta), pack4(meta + 4));
        nd.w2 = make_uint4(pack4(q[0][0]), pack4(q[0][0] + 4), pack4(q[0][1]), pack4(q[0][1] + 4));
        nd.w3 = make_uint4(pack4(q[1][0]), pack4(q[1][0] + 4), pack4(q[1][1]), pack4(q[1][1] + 4));
        nd.w4 = make_uint4(pack4(q[2][0]), pack4(q[2][0] + 4), pack4(q[2][1]), pack4(q[2][1] + 4));
        out[w] = nd;
    }
    // surface-area cost of the level: sum over nodes of area(node) + sum over leaf slots of area(slot) * triangles
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if ((threadIdx.x & 31) == 0 && cost != 0.0) atomicAdd(sah, cost);
}

__global__ void single_tri_kernel(RawTris raw, WTri* __restrict__ wtris, int32_t* __restrict__ leaf_prims) {
    write_wtri(raw, 0, wtris);
    leaf_prims[0] = 0;
}

template <typename T>
struct Tmp {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return lt_dev_malloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    ~Tmp() { if (p) lt_dev_free(p); }
};
struct Events {                       // RAII: every early return of lt_bvh_build releases them
    cudaEvent_t e[5] = {};
    cudaError_t create() {
        for (auto& x : e) { cudaError_t r = cudaEventCreate(&x); if (r != cudaSuccess) return r; }
        return cudaSuccess;
    }
    ~Events() { for (auto x : e) if (x) cudaEventDestroy(x); }
};

}  // namespace lt

using namespace lt;

static void free_hierarchy(lt_scene* s) {
    if (s->lbvh_nodes) lt_dev_free(s->lbvh_nodes);
    if (s->wnodes) lt_dev_free(s->wnodes);
    if (s->wtris) lt_dev_free(s->wtris);
    if (s->wleaf_prims) lt_dev_free(s->wleaf_prims);
    s->lbvh_nodes = nullptr; s->wnodes = nullptr; s->wtris = nullptr; s->wleaf_prims = nullptr;
    s->n_lbvh_nodes = 0; s->n_wnodes = 0; s->wbvh_depth = 0; s->lbvh2_depth = 0;
}

static int bvh_build_body(lt_scene* s, int flags, lt_bvh_stats* out);

extern "C" int lt_bvh_build(lt_scene* s, int flags, lt_bvh_stats* out) {
    if (!s) { set_error("lt_bvh_build: null scene"); return LT_E_INVALID; }
    const int rc = bvh_build_body(s, flags, out);
    if (rc != LT_OK) free_hierarchy(s);          // never leave half a hierarchy behind (the traversals check for its parts)
    return rc;
}

static int bvh_build_body(lt_scene* s, int flags, lt_bvh_stats* out) {
    lt_ctx* ctx = s->ctx;
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = s->n_tris;
    lt_bvh_stats stats;
    memset(&stats, 0, sizeof stats);
    stats.n_prims = n;
    free_hierarchy(s);
    const RawTris raw{s->d_v0, s->d_v1, s->d_v2, s->d_mat, s->d_light};
    LT_CUDA_OK(lt_dev_malloc(&s->wtris, (size_t)n * sizeof(WTri)));
    LT_CUDA_OK(lt_dev_malloc(&s->wleaf_prims, (size_t)n * sizeof(int32_t)));
    if (n < 2) {   // a single primitive needs no hierarchy; the traversal tests it directly
        single_tri_kernel<<<1, 1, 0, st>>>(raw, s->wtris, s->wleaf_prims);
        LT_CUDA_OK(cudaGetLastError());
        LT_CUDA_OK(cudaStreamSynchronize(st));
        if (out) *out = stats;
        return LT_OK;
    }
    const int blocks = (n + LB_THREADS - 1) / LB_THREADS;
    const int tiles = (n + RS_TILE - 1) / RS_TILE;
    Tmp<Box> prim_boxes, node_boxes;
    Tmp<float> block_cb, cb;
    Tmp<uint64_t> keys_a, keys_b;
    Tmp<uint32_t> vals_a, vals_b, hist, cnt_int, cnt_tri, totals, sums;
    Tmp<int> left, right, par_i, par_l, arrive, maxd, rng_lo, rng_hi, wb, slots, wpos;
    Tmp<WNode> wtmp;
    Tmp<double> sah;
    LT_CUDA_OK(prim_boxes.alloc(n)); LT_CUDA_OK(node_boxes.alloc(n));
    LT_CUDA_OK(block_cb.alloc(6 * (size_t)blocks)); LT_CUDA_OK(cb.alloc(6));
    LT_CUDA_OK(keys_a.alloc(n)); LT_CUDA_OK(keys_b.alloc(n)); LT_CUDA_OK(vals_a.alloc(n)); LT_CUDA_OK(vals_b.alloc(n));
    LT_CUDA_OK(hist.alloc(256 * (size_t)tiles));
    LT_CUDA_OK(left.alloc(n)); LT_CUDA_OK(right.alloc(n)); LT_CUDA_OK(par_i.alloc(n)); LT_CUDA_OK(par_l.alloc(n));
    LT_CUDA_OK(rng_lo.alloc(n)); LT_CUDA_OK(rng_hi.alloc(n));
    LT_CUDA_OK(arrive.alloc(n)); LT_CUDA_OK(maxd.alloc(1));
    LT_CUDA_OK(wb.alloc(n)); LT_CUDA_OK(wpos.alloc(n)); LT_CUDA_OK(slots.alloc(8 * (size_t)n)); LT_CUDA_OK(cnt_int.alloc(n)); LT_CUDA_OK(cnt_tri.alloc(n));
    LT_CUDA_OK(totals.alloc(2)); LT_CUDA_OK(wtmp.alloc(n)); LT_CUDA_OK(sah.alloc(1));
    LT_CUDA_OK(sums.alloc(2 * (std::max<size_t>((size_t)n, 256 * (size_t)tiles) / SC_TILE + 2)));
    LT_CUDA_OK(cudaMemsetAsync(arrive.p, 0, (size_t)n * sizeof(int), st));
    LT_CUDA_OK(cudaMemsetAsync(maxd.p, 0, sizeof(int), st));
    LT_CUDA_OK(cudaMemsetAsync(sah.p, 0, sizeof(double), st));
    LT_CUDA_OK(cudaMemsetAsync(wb.p, 0, sizeof(int), st));              // wide node 0 collapses binary node 0
    Events ev;
    LT_CUDA_OK(ev.create());
    cudaEventRecord(ev.e[0], st);
    prim_boxes_kernel<<<blocks, LB_THREADS, 0, st>>>(s->d_v0, s->d_v1, s->d_v2, n, prim_boxes.p, block_cb.p);
    reduce_cb_kernel<<<1, LB_THREADS, 0, st>>>(block_cb.p, blocks, cb.p);
    morton_kernel<<<blocks, LB_THREADS, 0, st>>>(prim_boxes.p, n, cb.p, keys_a.p, vals_a.p);
    cudaEventRecord(ev.e[1], st);
    uint64_t *ki = keys_a.p, *ko = keys_b.p;
    uint32_t *vi = vals_a.p, *vo = vals_b.p;
    for (int pass = 0; pass < 8; ++pass) {           // 63-bit codes + the size-class bit: 8 digits of 8 bits
        int shift = pass * 8;
        rs_hist_kernel<<<tiles, LB_THREADS, 0, st>>>(ki, n, shift, hist.p);
        if (256 * tiles <= SC_MIN) rs_scan_kernel<<<1, LB_THREADS, 0, st>>>(hist.p, 256 * tiles);
        else scan_multi(hist.p, nullptr, 256 * tiles, totals.p, sums.p, st);
        rs_scatter_kernel<<<tiles, LB_THREADS, 0, st>>>(ki, vi, n, shift, hist.p, ko, vo);
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    cudaEventRecord(ev.e[2], st);
    const int iblocks = (n - 1 + LB_THREADS - 1) / LB_THREADS;
    karras_kernel<<<iblocks, LB_THREADS, 0, st>>>(ki, n, left.p, right.p, par_i.p, par_l.p, rng_lo.p, rng_hi.p);
    refit_kernel<<<blocks, LB_THREADS, 0, st>>>(prim_boxes.p, vi, n, left.p, right.p, par_i.p, par_l.p, node_boxes.p, arrive.p);
    depth_kernel<<<blocks, LB_THREADS, 0, st>>>(n, par_i.p, par_l.p, maxd.p);
    cudaEventRecord(ev.e[3], st);
    LT_CUDA_OK(cudaGetLastError());
    // ---- collapse, breadth first: the wide nodes of a level are contiguous, children in slot order
    const int leaf_max = ((flags >> 8) & 3) ? ((flags >> 8) & 3) : LT_WBVH_LEAF_MAX;
    const TreeView tv{left.p, right.p, rng_lo.p, rng_hi.p, node_boxes.p, prim_boxes.p, vi, leaf_max};
    int level_start = 0, level_n = 1, tri_cursor = 0, levels = 0;
    while (level_n > 0) {
        const int lb = (level_n + LB_THREADS - 1) / LB_THREADS;
        collapse_kernel<<<lb, LB_THREADS, 0, st>>>(tv, wb.p, level_start, level_n, slots.p, cnt_int.p, cnt_tri.p);
        if (level_n <= SC_MIN) scan2_kernel<<<1, 1024, 0, st>>>(cnt_int.p, cnt_tri.p, level_n, totals.p);
        else scan_multi(cnt_int.p, cnt_tri.p, level_n, totals.p, sums.p, st);
        emit_wide_kernel<<<lb, LB_THREADS, 0, st>>>(tv, raw, wb.p, level_start, level_n, level_start + level_n, tri_cursor, slots.p, cnt_int.p,
                                                    cnt_tri.p, wtmp.p, s->wtris, s->wleaf_prims, wpos.p, sah.p);
        uint32_t tot[2];
        LT_CUDA_OK(cudaMemcpyAsync(tot, totals.p, sizeof tot, cudaMemcpyDeviceToHost, st));
        LT_CUDA_OK(cudaStreamSynchronize(st));
        level_start += level_n;
        level_n = (int)tot[0];
        tri_cursor += (int)tot[1];
        ++levels;
        if (levels > 4 * LT_LBVH_STACK) { free_hierarchy(s); set_error("wide-BVH collapse did not terminate"); return LT_E_OVERFLOW; }
    }
    const int n_wide = level_start;
    if (tri_cursor != n) { free_hierarchy(s); set_error("wide-BVH collapse placed %d of %d triangles", tri_cursor, n); return LT_E_OVERFLOW; }
    int n_binary = 0;
    if (flags & LT_BVH_KEEP_BINARY) {
        if (n >= (1 << 28)) { free_hierarchy(s); set_error("LT_BVH_KEEP_BINARY: at most 2^28 - 1 triangles"); return LT_E_OVERFLOW; }
        const int bleaf = ((flags >> 10) & 3) ? ((flags >> 10) & 3) : std::min(leaf_max, LT_LBVH2_LEAF_MAX);   // never above the slots' size
        mark_used_kernel<<<iblocks, LB_THREADS, 0, st>>>(n, rng_lo.p, rng_hi.p, std::min(bleaf, leaf_max), cnt_int.p);
        if (n - 1 <= SC_MIN) {
            LT_CUDA_OK(cudaMemsetAsync(cnt_tri.p, 0, (size_t)(n - 1) * sizeof(uint32_t), st));
            scan2_kernel<<<1, 1024, 0, st>>>(cnt_int.p, cnt_tri.p, n - 1, totals.p);
        } else {
            scan_multi(cnt_int.p, nullptr, n - 1, totals.p, sums.p, st);
        }
        uint32_t tot[2];
        LT_CUDA_OK(cudaMemcpyAsync(tot, totals.p, sizeof tot, cudaMemcpyDeviceToHost, st));
        LT_CUDA_OK(cudaStreamSynchronize(st));
        n_binary = (int)tot[0];
        LT_CUDA_OK(lt_dev_malloc(&s->lbvh_nodes, (size_t)n_binary * sizeof(LbvhNode)));
        emit_kernel<<<iblocks, LB_THREADS, 0, st>>>(prim_boxes.p, vi, n, left.p, right.p, rng_lo.p, rng_hi.p, std::min(bleaf, leaf_max), cnt_int.p,
                                                    node_boxes.p, wpos.p, s->lbvh_nodes);
        s->n_lbvh_nodes = n_binary;
    }
    LT_CUDA_OK(lt_dev_malloc(&s->wnodes, (size_t)n_wide * sizeof(WNode)));
    LT_CUDA_OK(cudaMemcpyAsync(s->wnodes, wtmp.p, (size_t)n_wide * sizeof(WNode), cudaMemcpyDeviceToDevice, st));
    cudaEventRecord(ev.e[4], st);
    LT_CUDA_OK(cudaGetLastError());
    LT_CUDA_OK(cudaStreamSynchronize(st));
    float ms_all = 0, ms_sort = 0, ms_collapse = 0;
    cudaEventElapsedTime(&ms_all, ev.e[0], ev.e[4]);
    cudaEventElapsedTime(&ms_sort, ev.e[1], ev.e[2]);
    cudaEventElapsedTime(&ms_collapse, ev.e[3], ev.e[4]);
    int depth = 0;
    Box root;
    double cost = 0.0;
    LT_CUDA_OK(cudaMemcpy(&depth, maxd.p, sizeof(int), cudaMemcpyDeviceToHost));
    LT_CUDA_OK(cudaMemcpy(&root, node_boxes.p, sizeof(Box), cudaMemcpyDeviceToHost));
    LT_CUDA_OK(cudaMemcpy(&cost, sah.p, sizeof(double), cudaMemcpyDeviceToHost));
    s->n_wnodes = n_wide;
    s->wbvh_depth = levels;
    s->lbvh2_depth = depth;
    stats.n_nodes = n_wide;
    stats.max_depth = levels;
    stats.binary_depth = depth;
    stats.build_ms = ms_all;
    stats.sort_ms = ms_sort;
    stats.collapse_ms = ms_collapse;
    stats.bytes = (int64_t)n_wide * (int64_t)sizeof(WNode) + (int64_t)n * (int64_t)sizeof(WTri);
    stats.binary_bytes = (flags & LT_BVH_KEEP_BINARY) ? (int64_t)n_binary * (int64_t)sizeof(LbvhNode) + (int64_t)n * (int64_t)sizeof(WTri) : 0;
    for (int k = 0; k < 3; ++k) { stats.root_bmin[k] = root.lo[k]; stats.root_bmax[k] = root.hi[k]; }
    const double ra = (double)box_area(root);
    stats.sah_cost = ra > 0 ? cost / ra : 0.0;
    if (out) *out = stats;
    if (levels > LT_WBVH_MAX_DEPTH || ((flags & LT_BVH_KEEP_BINARY) && depth > LT_LBVH_STACK)) {
        free_hierarchy(s);
        set_error("hierarchy depth %d (binary %d) exceeds the traversal stack (%d / %d)", levels, depth, LT_WBVH_MAX_DEPTH, LT_LBVH_STACK);
        return LT_E_OVERFLOW;
    }
    return LT_OK;
}
