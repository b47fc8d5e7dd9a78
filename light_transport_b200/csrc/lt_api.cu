// lt_api.cu — the C-ABI (include/lt_b200.h): handles, host<->device staging, launches.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include <mutex>
#include <algorithm>
#include <cmath>

#include "lt_megakernel.cuh"
#include "lt_wavefront.cuh"
#include "lt_record.cuh"
#include "lt_host.h"

namespace lt {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

// ---- pointer staging: caller data may live on the host or on the device -------------------------
// ---- device memory (see lt_host.h) ---------------------------------------------------------------
static cudaStream_t g_alloc_stream[64] = {};
static std::once_flag g_alloc_once[64];
static cudaStream_t alloc_stream() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    std::call_once(g_alloc_once[dev], [dev] {              // distinct handles may be used from distinct threads
        cudaStreamCreateWithFlags(&g_alloc_stream[dev], cudaStreamNonBlocking);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    });
    return g_alloc_stream[dev];
}
cudaError_t lt_dev_malloc_impl(void** p, size_t bytes) {
    cudaStream_t st = alloc_stream();
    cudaError_t e = cudaMallocAsync(p, bytes, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}
cudaError_t lt_dev_free_impl(void* p) { return p ? cudaFreeAsync(p, alloc_stream()) : cudaSuccess; }

static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static bool is_pinned_host_ptr(const void* p) {        // cudaHostAlloc'ed or lt_host_register'ed: the DMA engine can write it directly
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

template <typename T>
struct InBuf {   // read-only input
    const T* dev = nullptr;
    T* owned = nullptr;
    int init(const T* p, size_t n, cudaStream_t s) {
        if (!p || n == 0) return LT_OK;
        if (is_device_ptr(p)) { dev = p; return LT_OK; }
        LT_CUDA_OK(lt_dev_malloc(&owned, n * sizeof(T)));
        LT_CUDA_OK(cudaMemcpyAsync(owned, p, n * sizeof(T), cudaMemcpyHostToDevice, s));
        dev = owned;
        return LT_OK;
    }
    ~InBuf() { if (owned) lt_dev_free(owned); }
};
template <typename T>
struct OutBuf {  // output; `accumulate`: host destination receives += instead of =
    T* dev = nullptr;
    T* owned = nullptr;
    T* host = nullptr;
    size_t n = 0;
    bool accumulate = false;
    int init(T* p, size_t n_, bool acc, cudaStream_t s) {
        if (!p || n_ == 0) return LT_OK;
        n = n_;
        accumulate = acc;
        if (is_device_ptr(p)) { dev = p; return LT_OK; }
        host = p;
        LT_CUDA_OK(lt_dev_malloc(&owned, n * sizeof(T)));
        if (acc) LT_CUDA_OK(cudaMemsetAsync(owned, 0, n * sizeof(T), s));
        dev = owned;
        return LT_OK;
    }
    int finish(cudaStream_t s) {
        if (!owned) return LT_OK;
        if (!accumulate) {
            LT_CUDA_OK(cudaMemcpyAsync(host, owned, n * sizeof(T), cudaMemcpyDeviceToHost, s));
            LT_CUDA_OK(cudaStreamSynchronize(s));
        } else {
            std::vector<T> tmp(n);
            LT_CUDA_OK(cudaMemcpyAsync(tmp.data(), owned, n * sizeof(T), cudaMemcpyDeviceToHost, s));
            LT_CUDA_OK(cudaStreamSynchronize(s));
            for (size_t i = 0; i < n; ++i) host[i] += tmp[i];
        }
        return LT_OK;
    }
    ~OutBuf() { if (owned) lt_dev_free(owned); }
};

template <typename T>
static int upload(T** dst, const std::vector<T>& v) {
    if (*dst) { lt_dev_free(*dst); *dst = nullptr; }
    if (v.empty()) return LT_OK;
    LT_CUDA_OK(lt_dev_malloc(dst, v.size() * sizeof(T)));
    LT_CUDA_OK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return LT_OK;
}

template <typename R>
static void fill_view(const lt_scene* s, SceneView<R>& v) {
    const SceneArrays<R>& a = s->arrays<R>();
    v.tris = a.tris; v.mats = a.mats; v.lights = a.lights; v.cnodes = a.cnodes;
    v.cleaves = a.cleaves;
    v.nleaves = a.nleaves;
    v.n_nleaves = a.nleaves ? s->n_nleaves : 0;
    v.lbvh = s->lbvh_nodes;
    v.wnodes = s->wnodes; v.wtris = s->wtris; v.wleaf_prims = s->wleaf_prims; v.wbvh_depth = s->wbvh_depth; v.wmagic = 0x47000000u;
    v.n_tris = s->n_tris; v.n_mats = s->n_mats; v.n_lights = s->n_lights;
    v.n_cnodes = s->n_cnodes; v.n_cleaves = s->n_cleaves; v.n_lbvh_nodes = s->n_lbvh_nodes;
    v.n_cflat = s->n_cflat;
    for (int c = 0; c < 3; ++c) { v.bbox_min[c] = s->bbox_min[c]; v.bbox_max[c] = s->bbox_max[c]; }
}

template <typename R> static int build_compat_arrays(lt_scene* s);

// Triangle records are formed ON THE DEVICE from the raw float64 arrays (uploaded once, also the LBVH
// builder's input): a 1 M-triangle mesh never goes through a host loop.  ab/ac are float64 differences
// rounded to R afterwards, exactly as R/src/intersects.py:64-65 forms them per call.
template <typename R>
__global__ void build_tris_kernel(const double* __restrict__ v0, const double* __restrict__ v1, const double* __restrict__ v2,
                                  const double* __restrict__ nn, const int32_t* __restrict__ mat, const uint8_t* __restrict__ light, int n,
                                  TriRec<R>* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TriRec<R> t;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double a = v0[3 * i + c], b = v1[3 * i + c], cc = v2[3 * i + c];
        t.v0[c] = (R)a;
        t.ab[c] = (R)__dsub_rn(b, a);
        t.ac[c] = (R)__dsub_rn(cc, a);
        t.n[c] = (R)nn[3 * i + c];
    }
    t.mat = mat[i];
    t.is_light = light ? light[i] : 0;
    t.pad[0] = t.pad[1] = 0;
    out[i] = t;
}

template <typename T>
static int upload_raw(T** dst, const T* src, size_t n, cudaStream_t st) {
    LT_CUDA_OK(lt_dev_malloc(dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) LT_CUDA_OK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return LT_OK;
}

// build the R-typed device arrays of a scene on first use (float32 at creation, float64 when a float64 call needs it)
template <typename R>
static int ensure_arrays(lt_scene* s) {
    SceneArrays<R>& a = s->arrays<R>();
    if (a.tris && a.complete) return LT_OK;
    a.release();                                       // a half-built set from a failed earlier attempt is rebuilt from scratch
    cudaStream_t st = s->ctx->stream;
    const int n = s->n_tris;
    LT_CUDA_OK(lt_dev_malloc(&a.tris, (size_t)n * sizeof(TriRec<R>)));
    build_tris_kernel<R><<<(n + 255) / 256, 256, 0, st>>>(s->d_v0, s->d_v1, s->d_v2, s->d_n, s->d_mat, s->d_light, n, a.tris);
    LT_CUDA_OK(cudaGetLastError());
    std::vector<MatRec<R>> mats(s->n_mats);
    for (int i = 0; i < s->n_mats; ++i) {
        MatRec<R>& m = mats[i];
        memset(&m, 0, sizeof m);
        for (int c = 0; c < 3; ++c) m.kd[c] = (R)s->h_mat_kd[3 * i + c];
        m.emission = (R)s->h_mat_emission[i];
        m.ior = (R)s->h_mat_ior[i];
        m.transmission = (R)s->h_mat_transmission[i];
        m.flags = s->h_mat_flags[i];
    }
    std::vector<LightRec<R>> lights(s->n_lights);
    for (int i = 0; i < s->n_lights; ++i) {
        LightRec<R>& l = lights[i];
        memset(&l, 0, sizeof l);
        for (int c = 0; c < 3; ++c) {
            l.pos[c] = (R)s->h_light_pos[3 * i + c];
            l.n[c] = (R)s->h_light_n[3 * i + c];
            l.lekd[c] = (R)s->h_light_lekd[3 * i + c];
        }
        l.area = (R)s->h_light_area[i];
    }
    int rc;
    if ((rc = upload(&a.mats, mats))) return rc;
    if ((rc = upload(&a.lights, lights))) return rc;
    if (s->h_nodes.n_nodes > 0 && (rc = build_compat_arrays<R>(s))) return rc;
    if (!s->h_small_v.empty()) {
        // NEAREST on a small scene: bounds of every triangle, so the all-primitive scan can cull like COMPAT does
        // (lt_intersect.cuh).  float32 bounds are rounded OUTWARDS from the float64 vertices: culling must stay conservative.
        std::vector<CompatLeaf<R>> nl(n);
        for (int i = 0; i < n; ++i) {
            CompatLeaf<R>& l = nl[i];
            memset(&l, 0, sizeof l);
            for (int c = 0; c < 3; ++c) {
                double lo = INFINITY, hi = -INFINITY;
                for (int v = 0; v < 3; ++v) { const double x = s->h_small_v[((size_t)i * 3 + v) * 3 + c]; lo = std::min(lo, x); hi = std::max(hi, x); }
                R rlo = (R)lo, rhi = (R)hi;
                if (sizeof(R) == 4) {
                    if ((double)rlo > lo) rlo = (R)nextafterf((float)rlo, -INFINITY);
                    if ((double)rhi < hi) rhi = (R)nextafterf((float)rhi, INFINITY);
                }
                l.bmin[c] = rlo; l.bmax[c] = rhi;
            }
            l.prim_off = i; l.n_prims = 1; l.kill = 0; l.flat_base = i;
        }
        if ((rc = upload(&a.nleaves, nl))) return rc;
        s->n_nleaves = n;
    }
    LT_CUDA_OK(cudaStreamSynchronize(st));
    a.complete = true;
    return LT_OK;
}

template <typename R>
static int build_compat_arrays(lt_scene* s) {
    const HostNodes* d = &s->h_nodes;
    std::vector<CompatNode<R>> nodes(d->n_nodes);
    std::vector<CompatLeaf<R>> leaves;
    std::vector<int> leafpos(d->n_nodes, -1);
    for (int i = 0; i < d->n_nodes; ++i) {
        CompatNode<R>& n = nodes[i];
        for (int c = 0; c < 3; ++c) { n.bmin[c] = (R)d->bmin[3 * i + c]; n.bmax[c] = (R)d->bmax[3 * i + c]; }
        n.prim_off = d->prim_off[i]; n.second = d->second[i]; n.n_prims = d->n_prims[i]; n.axis = d->axis[i];
    }
    for (int i = d->n_nodes - 1; i >= 0; --i)
        if (nodes[i].n_prims > 0) {
            CompatLeaf<R> l;
            memset(&l, 0, sizeof l);
            for (int c = 0; c < 3; ++c) { l.bmin[c] = nodes[i].bmin[c]; l.bmax[c] = nodes[i].bmax[c]; }
            l.prim_off = nodes[i].prim_off; l.n_prims = nodes[i].n_prims;
            leafpos[i] = (int)leaves.size();
            nodes[i].axis = leafpos[i];            // node record of a leaf: its position in `leaves`
            leaves.push_back(l);
        }
    if (leaves.size() > 64) { set_error("COMPAT traversal supports at most 64 leaves (got %zu)", leaves.size()); return LT_E_INVALID; }
    int flat = 0;
    for (auto& l : leaves) { l.flat_base = flat; flat += l.n_prims; }
    s->n_cflat = flat;
    for (auto& l : leaves)                          // R/src/bvh_new.py:436: visited[leaf_idx] = True on the NODE list
        for (int k = 0; k < l.n_prims; ++k) {
            int p = l.prim_off + k;
            if (p < d->n_nodes && leafpos[p] >= 0) l.kill |= 1ull << leafpos[p];
        }
    SceneArrays<R>& a = s->arrays<R>();
    int rc;
    if ((rc = upload(&a.cnodes, nodes))) return rc;
    if ((rc = upload(&a.cleaves, leaves))) return rc;
    s->n_cnodes = d->n_nodes;
    s->n_cleaves = (int)leaves.size();
    return LT_OK;
}

static size_t staged_bytes(const lt_scene* s, int trav, int real) {
    size_t r = real == LT_REAL_F64 ? 8 : 4;
    auto al = [](size_t x) { return (x + 15) / 16 * 16; };
    size_t tri = r == 8 ? sizeof(TriRec<double>) : sizeof(TriPad);          // float32 triangles are staged with the padded stride
    size_t mat = r == 8 ? sizeof(MatRec<double>) : sizeof(MatRec<float>);
    size_t nod = r == 8 ? sizeof(CompatNode<double>) : sizeof(CompatNode<float>);
    size_t lef = r == 8 ? sizeof(CompatLeaf<double>) : sizeof(CompatLeaf<float>);
    if (trav == LT_TRAV_LBVH) return (size_t)std::max(s->wbvh_depth, 1) * LT_BLOCK * sizeof(uint2);   // traversal stacks
    if (trav == LT_TRAV_LBVH2) return 0;
    size_t b = al(tri * s->n_tris) + al(mat * s->n_mats);
    if (trav == LT_TRAV_COMPAT) b += al(nod * s->n_cnodes) + al(lef * s->n_cleaves);
    if (trav == LT_TRAV_COMPAT) b += al((size_t)((s->n_cleaves + 1) / 2) * 96 + 16);          // slab_pairs_bytes (float64: the filter)
    if (trav == LT_TRAV_COMPAT && r == 4) b += 8 * 4 * 12 * 4 + 32;                           // LT_CHAIN_BYTES
    if (trav == LT_TRAV_COMPAT) b += 32;                                                     // flattened primitive table
    if (trav == LT_TRAV_NEAREST && s->n_nleaves > 0)                                          // per-triangle bounds + the same tables
        b += al(lef * s->n_nleaves) + al((size_t)((s->n_nleaves + 1) / 2) * 96 + 16) + 32;
    return b;
}

static int check_traversal(const lt_scene* s, int trav) {
    if (trav == LT_TRAV_COMPAT) {
        if (s->n_cnodes == 0) { set_error("COMPAT traversal needs lt_scene_set_compat_bvh"); return LT_E_NO_BVH; }
    } else if (trav == LT_TRAV_LBVH) {
        if (!s->wtris || !s->wleaf_prims || (!s->wnodes && s->n_tris > 1)) { set_error("LBVH traversal needs (a successful) lt_bvh_build"); return LT_E_NO_BVH; }
    } else if (trav == LT_TRAV_LBVH2) {
        if (!s->wtris || !s->wleaf_prims || (!s->lbvh_nodes && s->n_tris > 1)) { set_error("LBVH2 traversal needs lt_bvh_build with LT_BVH_KEEP_BINARY"); return LT_E_NO_BVH; }
    } else if (trav != LT_TRAV_NEAREST) {
        set_error("unknown traversal %d", trav);
        return LT_E_INVALID;
    }
    return LT_OK;
}

template <typename R>
static int do_intersect(lt_scene* s, int traversal, const double* o, const double* d, int64_t n, int32_t* prim_out, double* t_out,
                        lt_counters* c) {
    { int rc0 = ensure_arrays<R>(s); if (rc0) return rc0; }
    lt_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    InBuf<double> bo, bd;
    OutBuf<int32_t> bp;
    OutBuf<double> bt;
    int rc;
    if ((rc = bo.init(o, (size_t)n * 3, st)) || (rc = bd.init(d, (size_t)n * 3, st)) || (rc = bp.init(prim_out, (size_t)n, false, st)) ||
        (rc = bt.init(t_out, (size_t)n, false, st)))
        return rc;
    LT_CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(lt_counters), st));
    IntersectParams<R> P;
    fill_view<R>(s, P.S);
    P.o = bo.dev; P.d = bd.dev; P.n = n; P.prim = bp.dev; P.t = bt.dev;
    P.counters = reinterpret_cast<unsigned long long*>(ctx->d_counters);
    P.eps = 0.000001;
    P.traversal = traversal;
    int grid = (int)std::min<int64_t>((n + LT_BLOCK - 1) / LT_BLOCK, (int64_t)ctx->prop.multiProcessorCount * 16);
    if (grid < 1) grid = 1;
    LT_CUDA_OK(cudaEventRecord(ctx->ev0, st));
    if (n > 0 && (rc = launch_intersect<R>(P, grid, st))) return rc;
    LT_CUDA_OK(cudaEventRecord(ctx->ev1, st));
    if ((rc = bp.finish(st)) || (rc = bt.finish(st))) return rc;
    LT_CUDA_OK(cudaStreamSynchronize(st));
    if (c) {
        LT_CUDA_OK(cudaMemcpy(c, ctx->d_counters, sizeof(lt_counters), cudaMemcpyDeviceToHost));
        c->rays_ext = (uint64_t)n;
        float ms = 0;
        LT_CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        c->kernel_ns = (uint64_t)(ms * 1e6);
        c->launches = 1;
    }
    return LT_OK;
}

template <typename R>
static int do_render(lt_scene* s, const lt_render_params* p, const double* rand0, const double* rand1, const int32_t* light_idx,
                     double* film_sum, double* samples_out, lt_counters* c) {
    { int rc0 = ensure_arrays<R>(s); if (rc0) return rc0; }
    lt_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const int W = p->width, H = p->height, D = p->max_depth;
    // only (0, 0) means "the whole film"; any other pixel1 <= pixel0 is an EMPTY share (a pixel-sharded caller with nothing to do)
    const bool whole = p->pixel0 == 0 && p->pixel1 == 0;
    const int64_t pix0 = p->pixel0, pix1 = whole ? (int64_t)W * H : std::max<int64_t>(p->pixel1, p->pixel0);
    const int ns = p->s1 - p->s0;
    const size_t npix_all = (size_t)W * H;
    InBuf<double> b0, b1;
    InBuf<int32_t> bl;
    OutBuf<double> bf, bs;
    int rc;
    if ((rc = b0.init(rand0, npix_all * p->spp * D, st)) || (rc = b1.init(rand1, npix_all * p->spp * D, st)) ||
        (rc = bl.init(light_idx, npix_all * p->spp * (size_t)p->n_light_idx, st)) || (rc = bf.init(film_sum, npix_all * 3, true, st)) ||
        (rc = bs.init(samples_out, npix_all * ns * 3, false, st)))
        return rc;
    if (bs.owned) LT_CUDA_OK(cudaMemsetAsync(bs.owned, 0, npix_all * ns * 3 * sizeof(double), st));
    LT_CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(lt_counters), st));

    WalkParams<R> P;
    memset(&P, 0, sizeof P);
    fill_view<R>(s, P.S);
    P.width = W; P.height = H; P.spp = p->spp; P.s0 = p->s0; P.ns = ns; P.max_depth = D;
    P.rr_start = p->rr_start; P.n_light_idx = p->n_light_idx; P.semantics = p->semantics; P.traversal = p->traversal;
    P.pixel0 = (int32_t)pix0;
    P.total_paths = (pix1 - pix0) * ns;
    P.rr_floor = p->rr_floor; P.eps = p->eps;
    for (int k = 0; k < 3; ++k) P.cam[k] = p->camera[k];
    P.f_distance = p->f_distance;
    // Scene.top/bottom/left/right (R/src/scene.py:60-64) and numba's np.linspace step
    const double aspect = (double)W / (double)H;
    P.top = 1 / aspect; P.bottom = -1 / aspect;
    P.step_y = H > 1 ? (P.bottom - P.top) / (double)(H - 1) : 0.0;
    P.step_x = W > 1 ? (1.0 - (-1.0)) / (double)(W - 1) : 0.0;
    P.seed = p->seed;
    P.rand0 = b0.dev; P.rand1 = b1.dev; P.light_idx = bl.dev;
    P.film = bf.dev; P.samples = bs.dev;
    P.counters = reinterpret_cast<unsigned long long*>(ctx->d_counters);

    const bool wavefront = p->scheduler == LT_SCHED_WAVEFRONT;
    if (p->traversal == LT_TRAV_LBVH2 && wavefront) { set_error("LT_TRAV_LBVH2 runs on the megakernel scheduler only"); return LT_E_INVALID; }
    const size_t staged = (wavefront && p->traversal == LT_TRAV_LBVH) ? 0 : staged_bytes(s, p->traversal, p->real);
    const size_t smem = wavefront ? wavefront_smem_bytes<R>(staged) : staged;
    if (smem > (size_t)ctx->prop.sharedMemPerBlockOptin) {
        set_error("scene needs %zu B of shared memory for traversal %d; use LT_TRAV_LBVH for large scenes", smem, p->traversal);
        return LT_E_OVERFLOW;
    }
    // phase-aligned iterations pay where most extension hits spawn a shadow query (flags bit 16 / 17 force off / on: A/B runs)
    P.phased = (p->flags & (1 << 17)) ? 1 : (p->flags & (1 << 16)) ? 0 : (s->diffuse_area_share >= 0.5 ? 1 : 0);
    int regs = 0, local_bytes = 0, bps = 0;
    if (wavefront) { if ((rc = wavefront_kernel_attrs<R>(p->traversal, p->semantics, &regs, &bps, smem))) return rc; }
    else if ((rc = walk_kernel_attrs<R>(p->traversal, p->semantics, P.phased, &regs, &local_bytes, &bps, smem))) return rc;
    if (bps < 1) bps = 1;
    if (((p->flags >> 20) & 0xf) && !wavefront) bps = std::min(bps, (p->flags >> 20) & 0xf);     // A/B: fewer resident CTAs per SM
    const int block = wavefront ? WfCfg<R>::NT : LT_BLOCK;
    int grid = ctx->prop.multiProcessorCount * bps;
    int64_t ppw = p->paths_per_warp;
    if (wavefront) {
        // granule = consecutive indices pulled by one CTA at a time
        if (ppw <= 0) {
            ppw = P.total_paths / ((int64_t)grid * 12);
            ppw = std::max<int64_t>(block * 4, std::min<int64_t>(ppw, 1 << 16));
        }
        int64_t granules = (P.total_paths + ppw - 1) / ppw;
        if (granules < grid) grid = (int)std::max<int64_t>(1, granules);
    } else {
        int64_t nwarps = (int64_t)grid * block / 32;
        if (ppw <= 0) {
            ppw = P.total_paths / (nwarps * 6);
            ppw = std::max<int64_t>(64, std::min<int64_t>(ppw, 4096));
            ppw = (ppw + 31) / 32 * 32;
        }
        int64_t granules = (P.total_paths + ppw - 1) / ppw;
        if (granules < nwarps) grid = (int)std::max<int64_t>(1, (granules * 32 + block - 1) / block);
    }
    P.paths_per_warp = (int32_t)ppw;
    P.lbvh_chunk = ((p->flags >> 8) & 0xff) ? ((p->flags >> 8) & 0xff) : (p->traversal == LT_TRAV_LBVH ? 16 : 24);
    // suspended-frame storage for REFERENCE walks: a frame at stack depth k has bounce >= k and can still
    // recurse at bounce D-1, so up to D suspended frames; the megakernel keeps the first LT_LOCAL_FRAMES locally
    P.arena = nullptr; P.arena_frames = 0; P.arena_threads = (int64_t)grid * block;
    const int local_frames = wavefront ? 0 : LT_LOCAL_FRAMES;
    if (p->semantics == LT_SEM_REFERENCE && D > local_frames) {
        P.arena_frames = D - local_frames;
        size_t bytes = (size_t)P.arena_frames * LT_ARENA_VALUES * (size_t)P.arena_threads * sizeof(R);
        if (bytes > ctx->arena_bytes) {
            if (ctx->arena) lt_dev_free(ctx->arena);
            ctx->arena = nullptr; ctx->arena_bytes = 0;
            LT_CUDA_OK(lt_dev_malloc(&ctx->arena, bytes));
            ctx->arena_bytes = bytes;
        }
        P.arena = reinterpret_cast<R*>(ctx->arena);
    }
    if (wavefront) LT_CUDA_OK(cudaMemsetAsync(ctx->d_granule, 0, sizeof(unsigned long long), st));
    // Hierarchy walks, opt-in (flags bit 19): pin the node array in the L2 with a persisting access-policy window on the stream.
    // Measured on C3 (1 M triangles: 64 MB of binary nodes + 48 MB of triangle records against a 126 MB L2): 2.87 Grays/s with the
    // window, 3.03 without — what the window keeps, the triangle records lose; the 8-wide hierarchy (11 MB of nodes) does not care.
    bool l2_window = false;
    if ((p->traversal == LT_TRAV_LBVH || p->traversal == LT_TRAV_LBVH2) && (p->flags & (1 << 19)) && ctx->prop.persistingL2CacheMaxSize > 0) {
        const void* base = p->traversal == LT_TRAV_LBVH ? (const void*)s->wnodes : (const void*)s->lbvh_nodes;
        const size_t bytes = p->traversal == LT_TRAV_LBVH ? (size_t)s->n_wnodes * sizeof(WNode) : (size_t)s->n_lbvh_nodes * sizeof(LbvhNode);
        if (base && bytes > 0) {
            const size_t carve = (size_t)ctx->prop.persistingL2CacheMaxSize;
            if (!ctx->l2_limit_set) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve); ctx->l2_limit_set = true; }
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            av.accessPolicyWindow.base_ptr = const_cast<void*>(base);
            av.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)ctx->prop.accessPolicyMaxWindowSize);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)av.accessPolicyWindow.num_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            l2_window = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
            if (!l2_window) cudaGetLastError();
        }
    }
    LT_CUDA_OK(cudaEventRecord(ctx->ev0, st));
    if (P.total_paths > 0) {
        rc = wavefront ? launch_wavefront<R>(P, ctx->d_granule, grid, smem, st) : launch_walk<R>(P, grid, smem, st);
        if (rc) return rc;
    }
    LT_CUDA_OK(cudaEventRecord(ctx->ev1, st));
    if (l2_window) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        av.accessPolicyWindow.num_bytes = 0;                     // window off; persisting lines return to normal
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
    }
    if ((rc = bf.finish(st)) || (rc = bs.finish(st))) return rc;
    LT_CUDA_OK(cudaStreamSynchronize(st));
    if (l2_window) cudaCtxResetPersistingL2Cache();
    lt_counters hc;
    LT_CUDA_OK(cudaMemcpy(&hc, ctx->d_counters, sizeof hc, cudaMemcpyDeviceToHost));
    float ms = 0;
    LT_CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    hc.kernel_ns = (uint64_t)(ms * 1e6);
    hc.launches = 1;
    if (c) *c = hc;
    if (hc.frame_overflow) { set_error("frame stack overflow on %llu pushes", (unsigned long long)hc.frame_overflow); return LT_E_OVERFLOW; }
    return LT_OK;
}

template <typename R>
static int do_records(lt_scene* s, const lt_render_params* p, const double* rand0, const double* rand1, const double* logit0,
                      const double* logit1, const int32_t* light_idx, double* samples_out, double* record_log_pdf,
                      double* scene_rlp, int8_t* bounce_record, double* direct_list, double* indirect_list, double* rlp_minus,
                      double* rlp_plus, lt_counters* c) {
    { int rc0 = ensure_arrays<R>(s); if (rc0) return rc0; }
    lt_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const int W = p->width, H = p->height, D = p->max_depth, spp = p->spp;
    // pixel sub-range [pixel0, pixel1): every array then holds (pixel1 - pixel0) * spp rows (0, 0 = the whole film)
    const bool whole = p->pixel0 == 0 && p->pixel1 == 0;
    const int64_t pix0 = whole ? 0 : p->pixel0, pix1 = whole ? (int64_t)W * H : p->pixel1;
    if (pix0 < 0 || pix1 > (int64_t)W * H || pix1 < pix0) { set_error("lt_render_records: bad pixel range"); return LT_E_INVALID; }
    const size_t paths = (size_t)(pix1 - pix0) * spp;
    const bool grads = rlp_minus && rlp_plus;
    InBuf<double> b0, b1, bl0, bl1;
    InBuf<int32_t> bl;
    OutBuf<double> o_s, o_rlp, o_srlp, o_dl, o_il, o_m, o_p;
    OutBuf<int8_t> o_br;
    int rc;
    if ((rc = b0.init(rand0, paths * D, st)) || (rc = b1.init(rand1, paths * D, st)) ||
        (rc = bl.init(light_idx, paths * (size_t)p->n_light_idx, st)) || (rc = o_s.init(samples_out, paths * 3, false, st)) ||
        (rc = o_rlp.init(record_log_pdf, paths * D, false, st)) || (rc = o_srlp.init(scene_rlp, paths * D, false, st)) ||
        (rc = o_br.init(bounce_record, paths * D, false, st)) || (rc = o_dl.init(direct_list, paths * 3 * D, false, st)) ||
        (rc = o_il.init(indirect_list, paths * 3 * D, false, st)))
        return rc;
    if (grads && ((rc = bl0.init(logit0, paths * D, st)) || (rc = bl1.init(logit1, paths * D, st)) ||
                  (rc = o_m.init(rlp_minus, paths * 2 * D * D, false, st)) || (rc = o_p.init(rlp_plus, paths * 2 * D * D, false, st))))
        return rc;
    LT_CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(lt_counters), st));
    WalkParams<R> P;
    memset(&P, 0, sizeof P);
    fill_view<R>(s, P.S);
    P.width = W; P.height = H; P.spp = spp; P.s0 = 0; P.ns = spp; P.max_depth = D;
    P.rr_start = p->rr_start; P.n_light_idx = p->n_light_idx; P.semantics = LT_SEM_REFERENCE; P.traversal = p->traversal;
    P.rr_floor = p->rr_floor; P.eps = p->eps;
    for (int k = 0; k < 3; ++k) P.cam[k] = p->camera[k];
    P.f_distance = p->f_distance;
    const double aspect = (double)W / (double)H;
    P.top = 1 / aspect; P.bottom = -1 / aspect;
    P.step_y = H > 1 ? (P.bottom - P.top) / (double)(H - 1) : 0.0;
    P.step_x = W > 1 ? (1.0 - (-1.0)) / (double)(W - 1) : 0.0;
    P.seed = p->seed;
    P.rand0 = b0.dev; P.rand1 = b1.dev; P.light_idx = bl.dev;
    P.samples = o_s.dev;
    P.counters = reinterpret_cast<unsigned long long*>(ctx->d_counters);
    P.pixel0 = (int32_t)pix0;
    P.row_base = pix0 * spp;
    RecParams RP;
    memset(&RP, 0, sizeof RP);
    RP.row_base = pix0 * spp;
    RP.gate = (p->flags & (1 << 18)) ? 0 : 1;
    RP.rlp_a = o_rlp.dev; RP.scene_rlp = o_srlp.dev; RP.bounce_record = reinterpret_cast<signed char*>(o_br.dev);
    RP.direct = (o_dl.dev && o_il.dev) ? o_dl.dev : nullptr; RP.indirect = RP.direct ? o_il.dev : nullptr;
    RP.grad = 0; RP.mult = 1;
    P.total_paths = (int64_t)paths;
    auto grid_for = [&](int64_t items) { return (int)std::max<int64_t>(1, std::min<int64_t>((items + LT_BLOCK - 1) / LT_BLOCK, (int64_t)ctx->prop.multiProcessorCount * 8)); };
    LT_CUDA_OK(cudaEventRecord(ctx->ev0, st));
    if ((rc = launch_record<R>(P, RP, grid_for(P.total_paths), st))) return rc;
    int launches = 1;
    if (grads) {       // calculate_gradients: 4*D gated replays per sample, only their returned arrays are kept
        RP.grad = 1; RP.mult = 4 * D;
        RP.rlp_a = o_m.dev; RP.rlp_b = o_p.dev; RP.logit0 = bl0.dev; RP.logit1 = bl1.dev;
        RP.direct = RP.indirect = nullptr;
        P.samples = nullptr; P.light_idx = nullptr;      // light picks do not influence record_log_pdf
        P.total_paths = (int64_t)paths * RP.mult;
        if ((rc = launch_record<R>(P, RP, grid_for(P.total_paths), st))) return rc;
        launches = 2;
    }
    LT_CUDA_OK(cudaEventRecord(ctx->ev1, st));
    if ((rc = o_s.finish(st)) || (rc = o_rlp.finish(st)) || (rc = o_srlp.finish(st)) || (rc = o_br.finish(st)) || (rc = o_dl.finish(st)) ||
        (rc = o_il.finish(st)) || (rc = o_m.finish(st)) || (rc = o_p.finish(st)))
        return rc;
    LT_CUDA_OK(cudaStreamSynchronize(st));
    lt_counters hc;
    LT_CUDA_OK(cudaMemcpy(&hc, ctx->d_counters, sizeof hc, cudaMemcpyDeviceToHost));
    float ms = 0;
    LT_CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    hc.kernel_ns = (uint64_t)(ms * 1e6);
    hc.launches = launches;
    if (c) *c = hc;
    if (hc.frame_overflow) { set_error("record kernel: frame stack overflow"); return LT_E_OVERFLOW; }
    return LT_OK;
}

__global__ void finalize_kernel(const double* film, int64_t n, double inv_spp, double* image) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = film[i] * inv_spp;   // color / spp; np.clip(color, 0, 1)
        image[i] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    }
}
__global__ void finalize_div_kernel(const double* film, int64_t n, double spp, double* image) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = film[i] / spp;       // R/src/path_tracing.py:304: color/scene.number_of_samples
        image[i] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    }
}

__global__ void fma_peak_kernel(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 0.999f, b = 1e-3f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
        a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void fma64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999, b = 1e-3;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// L2 -> SM bandwidth of the access pattern a hierarchy traversal makes: every thread gathers `granule` bytes (a multiple of
// 16, as 128-bit ld.global.cg loads that bypass L1) from pseudo-random granule-aligned offsets of a buffer that fits the L2.
// mode 0: consecutive lanes read consecutive granules (streaming); mode 1: every lane its own random granule (incoherent).
__global__ void l2_gather_kernel(const uint4* __restrict__ buf, uint32_t n_granules, int words, int iters, int mode, uint4* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    uint32_t x = tid * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        uint32_t g;
        if (mode == 0) g = (uint32_t)(((uint64_t)it * nthreads + tid) % n_granules);
        else { x ^= x << 13; x ^= x >> 17; x ^= x << 5; g = x % n_granules; }
        const uint4* p = buf + (size_t)g * words;
        for (int w = 0; w < words; ++w) {
            const uint4 v = __ldcg(p + w);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) sink[tid & 1023] = acc;     // keeps the loads alive
}

}  // namespace lt

using namespace lt;

extern "C" {

const char* lt_last_error(void) { return g_err; }
const char* lt_version(void) { return "lt_b200 0.1 (sm_100a)"; }

int lt_init(int device, lt_ctx** out) {
    if (!out) { set_error("lt_init: out is NULL"); return LT_E_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        return LT_E_CUDA;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(device));
    lt_ctx* c = new lt_ctx();
    c->device = device;
    const int rc = [c, device]() -> int {
        LT_CUDA_OK(cudaGetDeviceProperties(&c->prop, device));
        LT_CUDA_OK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        LT_CUDA_OK(lt_dev_malloc(&c->d_counters, sizeof(lt_counters)));
        LT_CUDA_OK(lt_dev_malloc(&c->d_granule, sizeof(unsigned long long)));
        LT_CUDA_OK(cudaEventCreate(&c->ev0));
        LT_CUDA_OK(cudaEventCreate(&c->ev1));
        return LT_OK;
    }();
    if (rc) { lt_destroy(c); return rc; }                  // a half-built context is released, not leaked
    *out = c;
    return LT_OK;
}

void lt_destroy(lt_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->arena) lt_dev_free(c->arena);
    if (c->film) lt_dev_free(c->film);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->d_counters) lt_dev_free(c->d_counters);
    if (c->d_granule) lt_dev_free(c->d_granule);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int lt_set_stream(lt_ctx* c, void* cuda_stream) {
    if (!c) { set_error("null ctx"); return LT_E_INVALID; }
    c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return LT_OK;
}

int lt_device_info(lt_ctx* c, int32_t* sm_count, int32_t* clock_khz, int64_t* global_mem, int32_t* cc_major, int32_t* cc_minor) {
    if (!c) { set_error("null ctx"); return LT_E_INVALID; }
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (clock_khz) *clock_khz = khz;
    if (global_mem) *global_mem = (int64_t)c->prop.totalGlobalMem;
    if (cc_major) *cc_major = c->prop.major;
    if (cc_minor) *cc_minor = c->prop.minor;
    return LT_OK;
}

int lt_scene_create(lt_ctx* ctx, const lt_scene_desc* d, lt_scene** out) {
    if (!ctx || !d || !out) { set_error("lt_scene_create: null argument"); return LT_E_INVALID; }
    *out = nullptr;
    if (d->n_tris <= 0 || d->n_mats <= 0 || d->n_lights < 0) { set_error("scene needs >=1 triangle and material"); return LT_E_INVALID; }
    for (int i = 0; i < d->n_tris; ++i)
        if (d->tri_mat[i] < 0 || d->tri_mat[i] >= d->n_mats) { set_error("tri_mat[%d]=%d out of range", i, d->tri_mat[i]); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    lt_scene* s = new lt_scene();
    s->ctx = ctx;
    s->n_tris = d->n_tris; s->n_mats = d->n_mats; s->n_lights = d->n_lights;
    const size_t n3 = 3 * (size_t)d->n_tris;
    // scene bounds for the all-primitive scan's early miss (it only makes sense for small scenes)
    for (int c = 0; c < 3; ++c) { s->bbox_min[c] = -INFINITY; s->bbox_max[c] = INFINITY; }
    if (d->n_tris <= 65536 && d->n_tris > 0) {
        for (int c = 0; c < 3; ++c) { s->bbox_min[c] = INFINITY; s->bbox_max[c] = -INFINITY; }
        const double* vs[3] = {d->tri_v0, d->tri_v1, d->tri_v2};
        for (int a = 0; a < 3; ++a)
            for (size_t i = 0; i < n3; ++i) {
                const int c = (int)(i % 3);
                s->bbox_min[c] = std::min(s->bbox_min[c], vs[a][i]);
                s->bbox_max[c] = std::max(s->bbox_max[c], vs[a][i]);
            }
    }
    if (d->n_tris <= 65536 && d->n_tris > 0) {      // how much of the surface spawns shadow rays: picks the walk's phase policy
        double all = 0, dif = 0;
        for (int i = 0; i < d->n_tris; ++i) {
            double ab[3], ac[3];
            for (int c = 0; c < 3; ++c) { ab[c] = d->tri_v1[3 * i + c] - d->tri_v0[3 * i + c]; ac[c] = d->tri_v2[3 * i + c] - d->tri_v0[3 * i + c]; }
            const double cx = ab[1] * ac[2] - ab[2] * ac[1], cy = ab[2] * ac[0] - ab[0] * ac[2], cz = ab[0] * ac[1] - ab[1] * ac[0];
            const double area = 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz);
            const int m = d->tri_mat[i];
            all += area;
            if (m >= 0 && m < d->n_mats && (d->mat_flags[m] & 1)) dif += area;
        }
        s->diffuse_area_share = all > 0 ? dif / all : 1.0;
    }
    if (d->n_tris <= LT_NEAREST_BOXES) {
        s->h_small_v.resize((size_t)d->n_tris * 9);
        const double* vs[3] = {d->tri_v0, d->tri_v1, d->tri_v2};
        for (int i = 0; i < d->n_tris; ++i)
            for (int v = 0; v < 3; ++v)
                for (int c = 0; c < 3; ++c) s->h_small_v[((size_t)i * 3 + v) * 3 + c] = vs[v][3 * (size_t)i + c];
    }
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = upload_raw(&s->d_v0, d->tri_v0, n3, st)) || (rc = upload_raw(&s->d_v1, d->tri_v1, n3, st)) ||
        (rc = upload_raw(&s->d_v2, d->tri_v2, n3, st)) || (rc = upload_raw(&s->d_n, d->tri_n, n3, st)) ||
        (rc = upload_raw(&s->d_mat, d->tri_mat, (size_t)d->n_tris, st))) { lt_scene_destroy(s); return rc; }
    if (d->tri_light && (rc = upload_raw(&s->d_light, d->tri_light, (size_t)d->n_tris, st))) { lt_scene_destroy(s); return rc; }
    s->h_mat_kd.assign(d->mat_kd, d->mat_kd + 3 * (size_t)d->n_mats);
    s->h_mat_emission.assign(d->mat_emission, d->mat_emission + d->n_mats);
    s->h_mat_ior.assign(d->mat_ior, d->mat_ior + d->n_mats);
    s->h_mat_transmission.assign(d->mat_transmission, d->mat_transmission + d->n_mats);
    s->h_mat_flags.assign(d->mat_flags, d->mat_flags + d->n_mats);
    s->h_light_pos.assign(d->light_pos, d->light_pos + 3 * (size_t)d->n_lights);
    s->h_light_n.assign(d->light_n, d->light_n + 3 * (size_t)d->n_lights);
    s->h_light_lekd.assign(d->light_lekd, d->light_lekd + 3 * (size_t)d->n_lights);
    s->h_light_area.assign(d->light_area, d->light_area + d->n_lights);
    if ((rc = ensure_arrays<float>(s))) { lt_scene_destroy(s); return rc; }     // float64 records are built on first float64 use
    *out = s;
    return LT_OK;
}

void lt_scene_destroy(lt_scene* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    s->f32.release();
    s->f64.release();
    if (s->lbvh_nodes) lt_dev_free(s->lbvh_nodes);
    if (s->wnodes) lt_dev_free(s->wnodes);
    if (s->wtris) lt_dev_free(s->wtris);
    if (s->wleaf_prims) lt_dev_free(s->wleaf_prims);
    if (s->d_v0) lt_dev_free(s->d_v0);
    if (s->d_v1) lt_dev_free(s->d_v1);
    if (s->d_v2) lt_dev_free(s->d_v2);
    if (s->d_n) lt_dev_free(s->d_n);
    if (s->d_mat) lt_dev_free(s->d_mat);
    if (s->d_light) lt_dev_free(s->d_light);
    delete s;
}

int lt_scene_set_compat_bvh(lt_scene* s, const lt_node_desc* d) {
    if (!s || !d || d->n_nodes <= 0) { set_error("lt_scene_set_compat_bvh: bad argument"); return LT_E_INVALID; }
    if (d->n_nodes > 128) { set_error("COMPAT traversal supports at most 128 nodes (got %d); use LT_TRAV_LBVH", d->n_nodes); return LT_E_INVALID; }
    for (int i = 0; i < d->n_nodes; ++i) {
        if (d->n_prims[i] > 0) {
            if (d->prim_off[i] < 0 || d->prim_off[i] + d->n_prims[i] > s->n_tris) { set_error("node %d: primitive range outside the scene", i); return LT_E_INVALID; }
            // Q9b: visited[leaf_idx] with leaf_idx >= len(nodes) raises IndexError in the reference
            if (d->prim_off[i] + d->n_prims[i] > d->n_nodes) {
                set_error("COMPAT undefined: leaf %d holds primitive index %d >= node count %d (the reference raises IndexError)", i,
                          d->prim_off[i] + d->n_prims[i] - 1, d->n_nodes);
                return LT_E_COMPAT_UNDEFINED;
            }
        } else {
            if (d->axis[i] < 0 || d->axis[i] > 2) { set_error("interior node %d: axis %d", i, d->axis[i]); return LT_E_INVALID; }
            if (d->second[i] <= i || d->second[i] >= d->n_nodes || i + 1 >= d->n_nodes) { set_error("interior node %d: bad child offset", i); return LT_E_INVALID; }
        }
    }
    LT_CUDA_OK(cudaSetDevice(s->ctx->device));
    HostNodes& h = s->h_nodes;
    h.n_nodes = d->n_nodes;
    h.bmin.assign(d->bmin, d->bmin + 3 * (size_t)d->n_nodes); h.bmax.assign(d->bmax, d->bmax + 3 * (size_t)d->n_nodes);
    h.prim_off.assign(d->prim_off, d->prim_off + d->n_nodes); h.second.assign(d->second, d->second + d->n_nodes);
    h.n_prims.assign(d->n_prims, d->n_prims + d->n_nodes); h.axis.assign(d->axis, d->axis + d->n_nodes);
    int rc;
    if ((rc = build_compat_arrays<float>(s))) return rc;
    if (s->f64.tris && (rc = build_compat_arrays<double>(s))) return rc;
    return LT_OK;
}

int lt_intersect(lt_scene* s, int traversal, int real, const double* o, const double* d, int64_t n, int32_t* prim_out, double* t_out,
                 lt_counters* c) {
    if (!s || n < 0 || (n > 0 && (!o || !d || !prim_out || !t_out))) { set_error("lt_intersect: bad argument"); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(s->ctx->device));
    int rc = check_traversal(s, traversal);
    if (rc) return rc;
    if (real == LT_REAL_F64) return do_intersect<double>(s, traversal, o, d, n, prim_out, t_out, c);
    if (real == LT_REAL_F32) return do_intersect<float>(s, traversal, o, d, n, prim_out, t_out, c);
    set_error("unknown real %d", real);
    return LT_E_INVALID;
}

int lt_render(lt_scene* s, const lt_render_params* p, const double* rand0, const double* rand1, const int32_t* light_idx, double* film_sum,
              double* samples_out, lt_counters* c) {
    if (!s || !p || !film_sum) { set_error("lt_render: null argument"); return LT_E_INVALID; }
    if (p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->s0 < 0 || p->s1 > p->spp || p->s1 < p->s0 || p->max_depth < 0) {
        set_error("lt_render: bad film/sample range");
        return LT_E_INVALID;
    }
    if ((rand0 == nullptr) != (rand1 == nullptr)) { set_error("rand0 and rand1 must both be given or both be NULL"); return LT_E_INVALID; }
    if (light_idx && p->n_light_idx <= 0) { set_error("light_idx given but n_light_idx == 0"); return LT_E_INVALID; }
    if (s->n_lights <= 0) { set_error("scene has no Light (next-event estimation needs one)"); return LT_E_INVALID; }
    if (p->semantics != LT_SEM_REFERENCE && p->semantics != LT_SEM_ITERATIVE) { set_error("unknown semantics"); return LT_E_INVALID; }
    if (p->scheduler != LT_SCHED_MEGAKERNEL && p->scheduler != LT_SCHED_WAVEFRONT) { set_error("unknown scheduler %d", p->scheduler); return LT_E_INVALID; }
    if (p->pixel1 > p->width * p->height || p->pixel0 < 0 || p->pixel1 < 0 || (p->pixel1 < p->pixel0 && p->pixel1 != 0)) {
        set_error("bad pixel range [%d, %d) for a %d x %d film", p->pixel0, p->pixel1, p->width, p->height);
        return LT_E_INVALID;
    }
    LT_CUDA_OK(cudaSetDevice(s->ctx->device));
    int rc = check_traversal(s, p->traversal);
    if (rc) return rc;
    if (p->real == LT_REAL_F64) return do_render<double>(s, p, rand0, rand1, light_idx, film_sum, samples_out, c);
    if (p->real == LT_REAL_F32) return do_render<float>(s, p, rand0, rand1, light_idx, film_sum, samples_out, c);
    set_error("unknown real %d", p->real);
    return LT_E_INVALID;
}

int lt_host_register(lt_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx || !ptr || bytes == 0) { set_error("lt_host_register: null argument"); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    LT_CUDA_OK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return LT_OK;
}

int lt_host_unregister(lt_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) { set_error("lt_host_unregister: null argument"); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    LT_CUDA_OK(cudaHostUnregister(ptr));
    return LT_OK;
}

int lt_render_image(lt_scene* s, const lt_render_params* p, const double* rand0, const double* rand1, const int32_t* light_idx,
                    double* image_out, double* samples_out, lt_counters* c) {
    if (!s || !p || !image_out) { set_error("lt_render_image: null argument"); return LT_E_INVALID; }
    lt_ctx* ctx = s->ctx;
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)p->width * p->height * 3;
    if (n > ctx->film_values) {
        if (ctx->film) lt_dev_free(ctx->film);
        ctx->film = nullptr; ctx->film_values = 0;
        LT_CUDA_OK(lt_dev_malloc(&ctx->film, n * sizeof(double)));
        ctx->film_values = n;
    }
    LT_CUDA_OK(cudaMemsetAsync(ctx->film, 0, n * sizeof(double), ctx->stream));
    int rc = lt_render(s, p, rand0, rand1, light_idx, ctx->film, samples_out, c);
    if (rc) return rc;
    const int ns = p->s1 - p->s0;
    int grid = (int)std::min<int64_t>(((int64_t)n + 255) / 256, (int64_t)ctx->prop.multiProcessorCount * 8);
    const bool dev_out = is_device_ptr(image_out);
    // finalize in place (elementwise) unless the caller gave a device buffer; then ONE copy of the image
    double* dst = dev_out ? image_out : ctx->film;
    finalize_div_kernel<<<std::max(grid, 1), 256, 0, ctx->stream>>>(ctx->film, (int64_t)n, (double)std::max(ns, 1), dst);
    LT_CUDA_OK(cudaGetLastError());
    if (!dev_out) {
        // D2H through the context's pinned staging buffer (full-rate DMA), then one host copy into the caller's pageable array
        const size_t bytes = n * sizeof(double);
        // (a caller whose array is page-locked — lt_host_register — gets the DMA straight into it)
        const bool direct = is_pinned_host_ptr(image_out);
        if (!direct && bytes > ctx->pinned_bytes) {
            if (ctx->pinned) cudaFreeHost(ctx->pinned);
            ctx->pinned = nullptr; ctx->pinned_bytes = 0;
            if (cudaHostAlloc(&ctx->pinned, bytes, cudaHostAllocDefault) == cudaSuccess) ctx->pinned_bytes = bytes;
            else cudaGetLastError();                                    // no pinned memory to be had: copy straight into the caller's array
        }
        void* dst_h = (!direct && ctx->pinned_bytes >= bytes) ? ctx->pinned : (void*)image_out;
        LT_CUDA_OK(cudaMemcpyAsync(dst_h, ctx->film, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        LT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        if (dst_h != (void*)image_out) memcpy(image_out, dst_h, bytes);
    } else {
        LT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }
    if (c) c->launches += 1;
    return LT_OK;
}

int lt_render_records(lt_scene* s, const lt_render_params* p, const double* rand0, const double* rand1, const double* logit0,
                      const double* logit1, const int32_t* light_idx, double* samples_out, double* record_log_pdf,
                      double* scene_record_log_pdf, int8_t* bounce_record, double* direct_list, double* indirect_list,
                      double* rlp_minus, double* rlp_plus, lt_counters* c) {
    if (!s || !p || !rand0 || !rand1 || !record_log_pdf || !scene_record_log_pdf || !bounce_record) {
        set_error("lt_render_records: rand0/rand1 and the three [H][W][spp][D] outputs are required");
        return LT_E_INVALID;
    }
    if (p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->max_depth <= 0 || p->max_depth > LT_REC_FRAMES) {
        set_error("lt_render_records: need 1 <= max_depth <= %d", LT_REC_FRAMES);
        return LT_E_INVALID;
    }
    if ((rlp_minus == nullptr) != (rlp_plus == nullptr) || (rlp_minus && (!logit0 || !logit1))) {
        set_error("lt_render_records: rlp_minus/rlp_plus go together and need logit0/logit1");
        return LT_E_INVALID;
    }
    if ((direct_list == nullptr) != (indirect_list == nullptr)) { set_error("direct_list and indirect_list go together"); return LT_E_INVALID; }
    if (light_idx && p->n_light_idx <= 0) { set_error("light_idx given but n_light_idx == 0"); return LT_E_INVALID; }
    if (s->n_lights <= 0) { set_error("scene has no Light"); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(s->ctx->device));
    int rc = check_traversal(s, p->traversal);
    if (rc) return rc;
    if (p->real == LT_REAL_F64)
        return do_records<double>(s, p, rand0, rand1, logit0, logit1, light_idx, samples_out, record_log_pdf, scene_record_log_pdf,
                                  bounce_record, direct_list, indirect_list, rlp_minus, rlp_plus, c);
    if (p->real == LT_REAL_F32)
        return do_records<float>(s, p, rand0, rand1, logit0, logit1, light_idx, samples_out, record_log_pdf, scene_record_log_pdf,
                                 bounce_record, direct_list, indirect_list, rlp_minus, rlp_plus, c);
    set_error("unknown real %d", p->real);
    return LT_E_INVALID;
}

int lt_film_finalize(lt_ctx* ctx, const double* film_sum, int64_t n, int64_t spp_total, double* image_out) {
    if (!ctx || !film_sum || !image_out || n < 0 || spp_total <= 0) { set_error("lt_film_finalize: bad argument"); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    InBuf<double> in;
    OutBuf<double> out;
    int rc;
    if ((rc = in.init(film_sum, (size_t)n, st)) || (rc = out.init(image_out, (size_t)n, false, st))) return rc;
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->prop.multiProcessorCount * 8);
    if (n > 0) finalize_div_kernel<<<std::max(grid, 1), 256, 0, st>>>(in.dev, n, (double)spp_total, out.dev);
    LT_CUDA_OK(cudaGetLastError());
    if ((rc = out.finish(st))) return rc;
    LT_CUDA_OK(cudaStreamSynchronize(st));
    return LT_OK;
}

int lt_measure_fp64_peak(lt_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) { set_error("null argument"); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    const int blocks = ctx->prop.multiProcessorCount * 8, threads = 256, iters = 1 << 13;
    double* d = nullptr;
    LT_CUDA_OK(lt_dev_malloc(&d, (size_t)blocks * threads * sizeof(double)));
    fma64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, 1 << 8);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        fma64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters);
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        if (e != cudaSuccess) { lt_dev_free(d); LT_CUDA_OK(e); }
        best = std::max(best, 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    }
    lt_dev_free(d);
    *tflops = best;
    return LT_OK;
}

int lt_measure_l2_bandwidth(lt_ctx* ctx, int64_t buffer_bytes, int granule_bytes, int mode, double* gbs) {
    if (!ctx || !gbs || buffer_bytes < (1 << 20) || granule_bytes < 16 || granule_bytes % 16 || (mode != 0 && mode != 1)) {
        set_error("lt_measure_l2_bandwidth: bad argument");
        return LT_E_INVALID;
    }
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    const int words = granule_bytes / 16;
    const uint32_t n_granules = (uint32_t)(buffer_bytes / granule_bytes);
    uint4 *buf = nullptr, *sink = nullptr;
    LT_CUDA_OK(lt_dev_malloc(&buf, (size_t)n_granules * granule_bytes));
    if (lt_dev_malloc(&sink, 1024 * sizeof(uint4)) != cudaSuccess) { lt_dev_free(buf); set_error("out of memory"); return LT_E_CUDA; }
    const int blocks = ctx->prop.multiProcessorCount * 8, threads = 256;
    const int iters = (int)std::max<int64_t>(64, (int64_t)(4ll << 30) / ((int64_t)blocks * threads * granule_bytes));   // ~4 GB per launch
    cudaMemsetAsync(buf, 1, (size_t)n_granules * granule_bytes, ctx->stream);
    l2_gather_kernel<<<blocks, threads, 0, ctx->stream>>>(buf, n_granules, words, iters, mode, sink);   // warms the L2
    double best = 0;
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 3 && e == cudaSuccess; ++rep) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        l2_gather_kernel<<<blocks, threads, 0, ctx->stream>>>(buf, n_granules, words, iters, mode, sink);
        cudaEventRecord(ctx->ev1, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        if (e == cudaSuccess) best = std::max(best, (double)iters * blocks * threads * granule_bytes / (ms * 1e-3) / 1e9);
    }
    lt_dev_free(buf);
    lt_dev_free(sink);
    LT_CUDA_OK(e);
    *gbs = best;
    return LT_OK;
}

int lt_measure_fp32_peak(lt_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) { set_error("null argument"); return LT_E_INVALID; }
    LT_CUDA_OK(cudaSetDevice(ctx->device));
    const int blocks = ctx->prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    float* d = nullptr;
    LT_CUDA_OK(lt_dev_malloc(&d, (size_t)blocks * threads * sizeof(float)));
    fma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, 1 << 10);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        LT_CUDA_OK(cudaEventRecord(ctx->ev0, ctx->stream));
        fma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters);
        LT_CUDA_OK(cudaEventRecord(ctx->ev1, ctx->stream));
        LT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        LT_CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        best = std::max(best, tf);
    }
    lt_dev_free(d);
    *tflops = best;
    return LT_OK;
}

}  // extern "C"
