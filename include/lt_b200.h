/* lt_b200.h — C-ABI of liblt_b200.so: the B200-native replacement for the Monte Carlo path walk
 * of zhouyifan233/light-transport.
 *
 * The reference has no FFI seam; its boundary is the Python call
 *     render_scene(scene, primitives, bvh)            R/src/path_tracing.py:371-387
 *     trace_path(scene, primitives, bvh, ray, ...)    R/src/path_tracing.py:23-159
 *     intersect_bvh(ray, primitives, linear_bvh)      R/src/bvh_new.py:413-482
 * (R/ = LightTransportSimulator/light_transport/).  A maintainer binds the entry points below with
 * ctypes from that Python code (INTEGRATION.md shows the stub); light_transport_b200/host.py is
 * that binding, shipped.
 *
 * Conventions: every function returns 0 (LT_OK) or a negative LT_E_* code; lt_last_error() gives
 * the message of the calling thread's last failure.  No exceptions cross the boundary.  Pointers
 * are caller-owned and only need to stay valid for the duration of the call; data pointers of
 * lt_intersect/lt_render/lt_film_* may be HOST or DEVICE memory (detected per pointer).  Handles are
 * opaque, freed by the matching *_destroy, not thread-safe individually.  Calls are synchronous.
 */
#ifndef LT_B200_H
#define LT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LT_OK 0
#define LT_E_INVALID (-1)      /* bad argument / unsupported combination */
#define LT_E_CUDA (-2)         /* CUDA runtime failure (message has the CUDA error string) */
#define LT_E_COMPAT_UNDEFINED (-3) /* COMPAT traversal on a node list where the reference raises IndexError (Q9b) */
#define LT_E_NO_BVH (-4)       /* traversal needs a structure that was not set/built */
#define LT_E_OVERFLOW (-5)     /* a capacity was exceeded (details in message) */
#define LT_E_NCCL (-6)         /* NCCL unavailable, or a collective failed (synchronously or asynchronously); the communicator is aborted */

/* traversal semantics */
#define LT_TRAV_COMPAT 0       /* emulate R/src/bvh_new.py:413-482 over the reference's node list, defects included */
#define LT_TRAV_NEAREST 1      /* true closest hit, EPSILON < t, ties -> lowest primitive index; all-primitive scan */
#define LT_TRAV_LBVH 2         /* same semantics as NEAREST through the GPU-built hierarchy: 8-wide, compressed (lt_bvh_build) */
#define LT_TRAV_LBVH2 3        /* the same through the binary Karras tree it is collapsed from (needs LT_BVH_KEEP_BINARY);
                                  lt_intersect and the megakernel walk only */
/* integrator semantics */
#define LT_SEM_REFERENCE 0     /* R/src/path_tracing.py:23-159 (recurse-and-continue) */
#define LT_SEM_ITERATIVE 1     /* R/src/path_tracing_fix1.py:17-136 */
/* arithmetic */
#define LT_REAL_F32 0
#define LT_REAL_F64 1
/* path scheduling */
#define LT_SCHED_MEGAKERNEL 0  /* persistent warps, in-warp regeneration */
#define LT_SCHED_WAVEFRONT 1   /* generate / extend / shade / shadow queues */

typedef struct lt_ctx lt_ctx;
typedef struct lt_scene lt_scene;
typedef struct lt_comm lt_comm;

/* Scene as structure-of-arrays, float64 like the reference's jitclass fields (a11):
 * PreComputedTriangle R/src/primitives.py:85-113, Material R/src/material.py:18-37,
 * Light R/src/scene.py:6-17.  3-vectors (the reference's homogeneous w is dropped, Q11). */
typedef struct {
    int32_t n_tris, n_mats, n_lights, reserved;
    const double *tri_v0, *tri_v1, *tri_v2;  /* [n_tris][3] vertex_1..3 */
    const double *tri_n;                     /* [n_tris][3] geometric unit normal */
    const int32_t *tri_mat;                  /* [n_tris]    material index */
    const uint8_t *tri_light;                /* [n_tris]    is_light */
    const double *mat_kd;                    /* [n_mats][3] color.diffuse */
    const double *mat_emission, *mat_ior, *mat_transmission; /* [n_mats] */
    const int32_t *mat_flags;                /* [n_mats] bit0 is_diffuse, bit1 is_mirror */
    const double *light_pos, *light_n;       /* [n_lights][3] Light.source, Light.normal */
    const double *light_lekd;                /* [n_lights][3] material.emission * material.color.diffuse */
    const double *light_area;                /* [n_lights]    Light.total_area */
} lt_scene_desc;

/* The reference's flattened node list (LinearBVHNode, R/src/bvh_new.py:53-67), for COMPAT. */
typedef struct {
    int32_t n_nodes, reserved;
    const double *bmin, *bmax;               /* [n_nodes][3] bounds.min_point / max_point */
    const int32_t *prim_off;                 /* primitives_offset   (-1 where None) */
    const int32_t *second;                   /* second_child_offset (-1 where None) */
    const int32_t *n_prims;                  /* n_primitives (0 on interior nodes) */
    const int32_t *axis;                     /* axis (-1 where None) */
} lt_node_desc;

typedef struct {
    int32_t width, height;       /* film size; Scene.width/height R/src/scene.py:57-58 */
    int32_t spp;                 /* samples per pixel of the WHOLE job: path id = (i*W+j)*spp + s */
    int32_t s0, s1;              /* this call renders samples s in [s0, s1) of every pixel */
    int32_t max_depth;           /* Scene.max_depth */
    int32_t semantics, traversal, real, scheduler;
    int32_t rr_start;            /* roulette when bounce > rr_start: 3 (path_tracing.py:148) / 5 (fix1:126) */
    int32_t n_light_idx;         /* replay: light picks provided per path (light_idx[...][n_light_idx]) */
    double rr_floor;             /* 0.05, path_tracing.py:149 */
    double eps;                  /* EPSILON = 1e-6, constants.py:12 */
    double camera[3];            /* Scene.camera[:3] */
    double f_distance;           /* Scene.f_distance */
    uint64_t seed;               /* Philox4x32-10 key (used when rand0 == NULL) */
    int32_t pixel0, pixel1;      /* optional pixel sub-range [pixel0,pixel1) in row-major order; 0,0 = all */
    int32_t paths_per_warp;      /* megakernel work granule; 0 = auto */
    int32_t flags;               /* bits 8-15: LBVH traversal steps per megakernel iteration (0 = default 24);
                                    bit 16 / bit 17: force phase-aligned iterations off / on (default: from the scene);
                                    bit 18 (lt_render_records): gradient replays are NOT gated by bounce_record (calculate_grad=False, the
                                    "more samples" block R/src/path_tracing.py:347); bit 19: persisting-L2 window on the hierarchy nodes (A/B: measured slower); bits 20-23: cap on resident CTAs per SM
                                    (0 = what the kernel was built for); rest reserved */
} lt_render_params;

typedef struct {
    uint64_t paths, rays_ext, rays_shadow, box_tests, tri_tests, max_frames, light_idx_overflow,
        frame_overflow, lane_steps, lane_slots, kernel_ns, launches;
    /* rays_ext counts every closest-hit query the reference makes for these paths (R/src/utils.py:59).
     * rays_memo of them were answered from a memo instead of a traversal: after a callee's ray misses,
     * trace_path's caller re-traces the SAME ray object (Q1) and must miss again. */
    uint64_t rays_memo;
} lt_counters;

typedef struct {
    int32_t n_prims, n_nodes;    /* triangles; 8-wide nodes of the traversal hierarchy */
    int32_t max_depth;           /* levels of the 8-wide hierarchy (= traversal stack entries needed) */
    int32_t binary_depth;        /* depth of the Karras tree it was collapsed from */
    float build_ms;              /* device time of the whole build */
    float sort_ms;
    double root_bmin[3], root_bmax[3];
    double sah_cost;             /* (sum over wide nodes of area(node) + sum over leaf slots of area(slot) * triangles) / area(root) */
    float collapse_ms;           /* of build_ms: collapse + quantise + triangle records */
    float reserved;
    int64_t bytes;               /* nodes (80 B) + float32 triangle records (48 B): what a float32 traversal touches */
    int64_t binary_bytes;        /* LT_BVH_KEEP_BINARY: binary nodes (64 B) + the same triangle records; else 0 */
} lt_bvh_stats;
/* lt_bvh_build flags */
/* bits 8-9: triangles per leaf slot of the wide hierarchy, 1..3 (0 = default 3)
 * bits 10-11: triangles per leaf of the binary tree kept by LT_BVH_KEEP_BINARY, 1..3, capped at the slot size (0 = default) */
#define LT_BVH_KEEP_BINARY 1     /* also keep the binary (Karras) nodes for LT_TRAV_LBVH2 */

int lt_init(int device, lt_ctx **out);
void lt_destroy(lt_ctx *);
/* Run all subsequent work of this context on `cuda_stream` (a cudaStream_t, e.g. torch's current
 * stream) so callers can bracket it with their own events; NULL restores the context's own stream. */
int lt_set_stream(lt_ctx *, void *cuda_stream);
int lt_device_info(lt_ctx *, int32_t *sm_count, int32_t *clock_khz, int64_t *global_mem, int32_t *cc_major,
                   int32_t *cc_minor);

int lt_scene_create(lt_ctx *, const lt_scene_desc *, lt_scene **out);          /* H2D once */
void lt_scene_destroy(lt_scene *);
int lt_scene_set_compat_bvh(lt_scene *, const lt_node_desc *);                 /* replaces the bvh argument of render_scene */
int lt_bvh_build(lt_scene *, int flags, lt_bvh_stats *out);                    /* GPU LBVH: replaces build_bvh/flatten_bvh, bvh_new.py:148-300 */

/* closest hit for an explicit ray batch (replaces intersect_bvh / hit_object for a ray set).
 * o,d: [n][3] float64; prim_out: -1 on miss; t_out: +inf on miss.  real selects fp32/fp64 arithmetic. */
int lt_intersect(lt_scene *, int traversal, int real, const double *o, const double *d, int64_t n,
                 int32_t *prim_out, double *t_out, lt_counters *c);

/* The walk (a1-a10): ray generation, traversal, shading, roulette, tally.
 * rand0/rand1: [H][W][spp][max_depth] float64 (Scene.rand_0/rand_1) or NULL -> Philox(seed);
 * light_idx:   [H][W][spp][n_light_idx] int32 replayed light picks or NULL -> Philox stream 1;
 * film_sum:    [H][W][3] float64, += sum over the rendered samples of per-path radiance;
 * samples_out: optional [H][W][s1-s0][3] float64 per-path radiance (R/src/path_tracing.py:297). */
int lt_render(lt_scene *, const lt_render_params *, const double *rand0, const double *rand1,
              const int32_t *light_idx, double *film_sum, double *samples_out, lt_counters *c);

/* Page-lock a caller-owned host array (cudaHostRegister) so that read-backs into it (lt_render_image's image_out) are one
 * DMA instead of a DMA into the library's staging buffer plus a host copy.  The reference hands scene.image back from every
 * render_scene call (R/src/path_tracing_old.py:171); the host side registers that array once.  Unregister before freeing it. */
int lt_host_register(lt_ctx *, void *ptr, size_t bytes);
int lt_host_unregister(lt_ctx *, void *ptr);

/* lt_render + lt_film_finalize in one call with the film kept on the device: renders samples
 * [s0,s1), writes image_out[H][W][3] = clip(film / (s1-s0), 0, 1) (host or device pointer).  This is
 * the whole of the reference's render_scene for one GPU: scene in (lt_scene_create), image out. */
int lt_render_image(lt_scene *, const lt_render_params *, const double *rand0, const double *rand1,
                    const int32_t *light_idx, double *image_out, double *samples_out, lt_counters *c);

/* The per-sample outputs of render_scene_samples (R/src/path_tracing.py:251-303) for replayed random
 * numbers, REFERENCE semantics, whole sample range, max_depth <= 32:
 *   samples_out            [H][W][spp][3]     per-path radiance                                  (:297)
 *   record_log_pdf         [H][W][spp][D]     the array trace_path returns                        (:288,300)
 *   scene_record_log_pdf   [H][W][spp][D]     Scene.record_log_pdf side effect                    (:95-96)
 *   bounce_record          [H][W][spp][D] i8  Scene.bounce_record side effect                     (:50-52,77-79,151-153)
 *   direct_list, indirect_list [H][W][spp][3][D]  the root frame's lists, optional               (:99-100)
 *   rlp_minus, rlp_plus    [H][W][spp][2D][D] calculate_gradients' replays (:203-249), optional; need
 *                          logit0/logit1 [H][W][spp][D] = log(u/(1-u)) (Scene.rand_0_logit, scene.py:70-71).
 * With a pixel sub-range (params.pixel0 < params.pixel1) only those pixels are walked and every array holds
 * (pixel1 - pixel0) * spp rows instead of H * W * spp. */
int lt_render_records(lt_scene *, const lt_render_params *, const double *rand0, const double *rand1,
                      const double *logit0, const double *logit1, const int32_t *light_idx, double *samples_out,
                      double *record_log_pdf, double *scene_record_log_pdf, int8_t *bounce_record, double *direct_list,
                      double *indirect_list, double *rlp_minus, double *rlp_plus, lt_counters *c);

/* image = clip(film_sum / spp_total, 0, 1)   (R/src/path_tracing.py:304-305) */
int lt_film_finalize(lt_ctx *, const double *film_sum, int64_t n_values, int64_t spp_total, double *image_out);

/* ---- multi-GPU: photons shard across ranks by sample range (disjoint Philox counters); the ONLY exchange of a pass is one
 * sum-reduction of the per-rank films over NCCL (NVLink / NVSwitch).  The reference is single-process (numba.prange threads,
 * R/src/path_tracing_old.py:142-149) and has no counterpart.  One process per GPU: rank 0 calls lt_comm_unique_id, the host
 * hands the 128 bytes to every rank (torch.distributed / MPI / a file), every rank calls lt_comm_init.
 * lt_film_reduce: film_sum (DEVICE, [n] float64) <- sum over ranks, on rank `root` (root < 0: on every rank), on the context's
 * stream; with spp_total > 0 the receiving rank(s) also finalize in place: film_sum <- clip(film_sum / spp_total, 0, 1)
 * (R/src/path_tracing.py:304-305).  NCCL errors, including asynchronous ones (ncclCommGetAsyncError), abort the communicator and
 * return LT_E_NCCL instead of hanging. */
int lt_comm_unique_id(void *id128);
int lt_comm_init(lt_ctx *, int n_ranks, int rank, const void *id128, lt_comm **out);
void lt_comm_destroy(lt_comm *);
int lt_film_reduce(lt_ctx *, lt_comm *, double *film_sum, int64_t n_values, int root, int64_t spp_total);
int lt_nccl_version(int *version);

/* ---- C4b: many-bounce photon walk in a homogeneous scattering medium (BASELINE.json config 4).  The
 * reference only has the Henyey-Greenstein evaluation (R/src/medium_samples.py:14-16, lt_henyey_greenstein);
 * the walk itself is beyond the reference and is pinned by the oracle + analytic expectations. */
typedef struct {
    double albedo, g, radius;      /* single-scattering albedo, HG asymmetry, sphere radius in mean free paths (<= 0: infinite) */
    int32_t max_depth, mode;       /* max collisions; 0 analog absorption, 1 implicit capture + roulette */
    int32_t rr_start, n_rbins;     /* roulette when collision index > rr_start; radial bins of the absorbed-weight tally */
    double rr_floor;
    uint64_t seed, first_photon;   /* Philox key; global index of the first photon (ranks take disjoint ranges) */
    int32_t real, reserved;
} lt_medium_params;
typedef struct {
    double escaped, absorbed, truncated, rr_killed, rr_gained, sum_mu;
    uint64_t n_collisions, n_segments, n_scatters;
} lt_medium_totals;
/* hist[max_depth+1]: photons by collision count; rbins[n_rbins]: absorbed weight by radial shell (host pointers) */
int lt_medium_walk(lt_ctx *, const lt_medium_params *, uint64_t n_photons, uint64_t *hist, double *rbins,
                   lt_medium_totals *totals, float *kernel_ms);
double lt_henyey_greenstein(double cos_theta, double g);

/* roofline denominators measured on this device: FP32 / FP64 FMA throughput (TFLOP/s), and the L2 -> SM bandwidth (GB/s) of
 * 128-bit gathers of `granule_bytes` from a `buffer_bytes` buffer that fits the L2 (mode 0: streaming, 1: one random granule
 * per lane - the access pattern of an incoherent hierarchy traversal) */
int lt_measure_fp32_peak(lt_ctx *, double *tflops);
int lt_measure_fp64_peak(lt_ctx *, double *tflops);
int lt_measure_l2_bandwidth(lt_ctx *, int64_t buffer_bytes, int granule_bytes, int mode, double *gbs);

const char *lt_last_error(void);
const char *lt_version(void);

#ifdef __cplusplus
}
#endif
#endif
