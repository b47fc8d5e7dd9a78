// lt_host.h — host-side handle definitions behind the opaque lt_ctx / lt_scene of include/lt_b200.h.
#pragma once
#include <vector>

#include "lt_scene.cuh"

namespace lt {

// Device memory comes from the device's stream-ordered pool with its release threshold lifted (lt_init): a scene that is
// created, rendered and destroyed per call (the reference-facing render_scene) reuses its blocks instead of mapping
// and unmapping hundreds of megabytes through the driver every time (1 M-triangle scene: the per-call overhead
// jittered between 60 and 650 ms).  Every API call synchronises its stream before it returns, so a block is idle
// when it is freed.
cudaError_t lt_dev_malloc_impl(void** p, size_t bytes);
cudaError_t lt_dev_free_impl(void* p);
template <typename T> inline cudaError_t lt_dev_malloc(T** p, size_t bytes) { return lt_dev_malloc_impl(reinterpret_cast<void**>(p), bytes); }
inline cudaError_t lt_dev_free(void* p) { return lt_dev_free_impl(p); }

// RAII device buffer for the temporaries of an API call: every early (error) return releases it
template <typename T>
struct DevBuf {
    T* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t n) { return lt_dev_malloc(&p, (n ? n : 1) * sizeof(T)); }
    ~DevBuf() { if (p) lt_dev_free(p); }
};

template <typename R>
struct SceneArrays {
    TriRec<R>* tris = nullptr;
    MatRec<R>* mats = nullptr;
    LightRec<R>* lights = nullptr;
    CompatNode<R>* cnodes = nullptr;
    CompatLeaf<R>* cleaves = nullptr;   // leaves in descending node order
    CompatLeaf<R>* nleaves = nullptr;   // NEAREST on small scenes: one bounds record per triangle, in primitive order
    bool complete = false;              // set by ensure_arrays once every upload has succeeded
    void release() {
        complete = false;
        if (tris) lt_dev_free(tris);
        if (mats) lt_dev_free(mats);
        if (lights) lt_dev_free(lights);
        if (cnodes) lt_dev_free(cnodes);
        if (cleaves) lt_dev_free(cleaves);
        if (nleaves) lt_dev_free(nleaves);
        tris = nullptr; mats = nullptr; lights = nullptr; cnodes = nullptr; cleaves = nullptr; nleaves = nullptr;
    }
};

}  // namespace lt

struct HostNodes {
    int n_nodes = 0;
    std::vector<double> bmin, bmax;
    std::vector<int32_t> prim_off, second, n_prims, axis;
};

struct lt_ctx {
    int device = 0;
    cudaDeviceProp prop;
    cudaStream_t stream = nullptr;       // stream in use
    cudaStream_t own_stream = nullptr;   // created by lt_init
    lt_counters* d_counters = nullptr;
    unsigned long long* d_granule = nullptr;   // wavefront scheduler: global granule counter
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    double* film = nullptr;              // device film of lt_render_image
    size_t film_values = 0;
    bool l2_limit_set = false;           // cudaLimitPersistingL2CacheSize raised once
    void* pinned = nullptr;              // pinned host staging for the image read-back of lt_render_image
    size_t pinned_bytes = 0;
};

struct lt_scene {
    lt_ctx* ctx = nullptr;
    int n_tris = 0, n_mats = 0, n_lights = 0;
    int n_cnodes = 0, n_cleaves = 0, n_cflat = 0;
    double bbox_min[3], bbox_max[3];   // bounds of all vertices (computed for small scenes, else +-inf)
    std::vector<double> h_small_v;     // [tri][vertex][3] of scenes with <= LT_NEAREST_BOXES triangles (per-triangle bounds, built per precision)
    int n_nleaves = 0;
    double diffuse_area_share = 1.0;   // share of the triangle area whose material is diffuse (scenes up to 65536 triangles; else 1)
    int n_lbvh_nodes = 0;
    lt::SceneArrays<float> f32;
    lt::SceneArrays<double> f64;
    lt::LbvhNode* lbvh_nodes = nullptr;        // binary nodes (only with LT_BVH_KEEP_BINARY)
    lt::WNode* wnodes = nullptr;               // 8-wide compressed hierarchy
    lt::WTri* wtris = nullptr;                 // float32 triangle records, leaf order
    int32_t* wleaf_prims = nullptr;            // leaf order -> primitive
    int n_wnodes = 0, wbvh_depth = 0, lbvh2_depth = 0;
    // raw float64 scene arrays on the device (uploaded once): input of the record builders and of the LBVH build
    double *d_v0 = nullptr, *d_v1 = nullptr, *d_v2 = nullptr, *d_n = nullptr;
    int32_t* d_mat = nullptr;
    uint8_t* d_light = nullptr;
    // small host copies for the lazily built float64 records
    std::vector<double> h_mat_kd, h_mat_emission, h_mat_ior, h_mat_transmission, h_light_pos, h_light_n, h_light_lekd, h_light_area;
    std::vector<int32_t> h_mat_flags;
    HostNodes h_nodes;
    template <typename R> lt::SceneArrays<R>& arrays();
    template <typename R> const lt::SceneArrays<R>& arrays() const;
};
template <> inline lt::SceneArrays<float>& lt_scene::arrays<float>() { return f32; }
template <> inline lt::SceneArrays<double>& lt_scene::arrays<double>() { return f64; }
template <> inline const lt::SceneArrays<float>& lt_scene::arrays<float>() const { return f32; }
template <> inline const lt::SceneArrays<double>& lt_scene::arrays<double>() const { return f64; }
