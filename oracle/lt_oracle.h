/* lt_oracle.h — CPU oracle for the photon/path walk.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, float64 restatement of the reference's hot path
 * (R/ = /root/reference/LightTransportSimulator/light_transport/), function by function, with the
 * reference file:line each one follows cited in lt_oracle.c.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product (liblt_b200.so and
 * the light_transport_b200 package) never does.
 *
 * Parity status: PINNED — checked against the .npz fixtures in tests/golden/, which were produced by running the
 * unmodified reference (numba) in the build container (tests/golden/make_golden.py).
 */
#ifndef LT_ORACLE_H
#define LT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t n_tris, n_mats, n_lights, n_nodes;
    const double *tri_v0, *tri_v1, *tri_v2, *tri_n;     /* [n_tris][3] */
    const int32_t *tri_mat;                             /* [n_tris] */
    const uint8_t *tri_light;                           /* [n_tris] is_light */
    const double *mat_kd;                               /* [n_mats][3] */
    const double *mat_emission, *mat_ior, *mat_transmission; /* [n_mats] */
    const int32_t *mat_flags;                           /* bit0 is_diffuse, bit1 is_mirror */
    const double *light_pos, *light_n, *light_lekd;     /* [n_lights][3] */
    const double *light_area;                           /* [n_lights] */
    const double *node_bmin, *node_bmax;                /* [n_nodes][3]  reference LinearBVHNode list */
    const int32_t *node_prim_off, *node_second, *node_n_prims, *node_axis;
} lto_scene;

typedef struct {
    int32_t width, height, spp, s0, s1, max_depth;
    int32_t semantics;   /* 0 REFERENCE (R/src/path_tracing.py), 1 ITERATIVE (path_tracing_fix1.py) */
    int32_t traversal;   /* 0 COMPAT (intersect_bvh emulation), 1 NEAREST (brute force) */
    int32_t rr_start;    /* roulette when bounce > rr_start */
    int32_t n_light_idx; /* replayed light indices per path (replay mode) */
    int32_t n_threads;   /* OpenMP threads, 0 = default */
    int32_t pad_;
    double rr_floor, eps, camera[3], f_distance;
    uint64_t seed;       /* Philox key when rand0 == NULL */
    int32_t pixel0, pixel1;   /* lto_render_records only: pixel sub-range [pixel0, pixel1) (row-major; 0,0 = whole film); the arrays then
                                 hold (pixel1 - pixel0) * spp rows */
    int32_t grad_ungated;     /* lto_render_records only: the gradient replays run with calculate_grad=False (not gated by bounce_record):
                                 the "more samples" block, R/src/path_tracing.py:347-348 */
    int32_t pad2_;
} lto_params;

typedef struct {
    uint64_t paths, rays_ext, rays_shadow, box_tests, tri_tests, max_frames, light_idx_overflow, errors;
} lto_counters;

/* closest-hit query for explicit rays; prim = -1 / t = inf on miss.  Returns 0, or -1 if the
 * reference would have raised IndexError (Q9b). */
int lto_intersect(const lto_scene *S, int traversal, const double *o, const double *d, int64_t n,
                  int32_t *prim_out, double *t_out, lto_counters *c);

/* render samples [s0,s1) of every pixel.  rand0/rand1: [H][W][spp][D] or NULL (=> Philox(seed));
 * light_idx: [H][W][spp][n_light_idx] or NULL (=> Philox stream 1).
 * film_sum [H][W][3] += per-sample radiance in sample order; samples_out (optional)
 * [H][W][s1-s0][3]. */
int lto_render(const lto_scene *S, const lto_params *P, const double *rand0, const double *rand1,
               const int32_t *light_idx, double *film_sum, double *samples_out, lto_counters *c);

/* per-pixel sum and sum of squares of the per-sample radiance for a SUBSET of pixels (row-major indices) of the W x H film,
 * samples [s0,s1), Philox streams: single pixels of a film too large to hold (config C2) */
int lto_render_pixels(const lto_scene *S, const lto_params *P, const int32_t *pixels, int32_t n_pixels, double *sum,
                      double *sumsq, lto_counters *c);

/* render_scene_samples' per-sample outputs for replayed random numbers:
 * samples [H][W][spp][3]; record_log_pdf (the array trace_path RETURNS), scene_record_log_pdf and
 * bounce_record (the arrays it writes into Scene) [H][W][spp][D]; direct/indirect lists of the root frame
 * [H][W][spp][3][D]; and, when rlp_minus/rlp_plus are given, calculate_gradients' replays
 * [H][W][spp][2D][D] driven by logit0/logit1 = log(u/(1-u)).  Whole sample range only. */
int lto_render_records(const lto_scene *S, const lto_params *P, const double *rand0, const double *rand1,
                       const double *logit0, const double *logit1, const int32_t *light_idx, double *samples_out,
                       double *record_log_pdf, double *scene_record_log_pdf, int8_t *bounce_record, double *direct_list,
                       double *indirect_list, double *rlp_minus, double *rlp_plus, lto_counters *cnt);

/* ---- C4b: many-bounce walk in a homogeneous scattering medium (beyond the reference: it only has the
 * Henyey-Greenstein evaluation R/src/medium_samples.py:14-16).  Same record as lt_medium_params. */
typedef struct {
    double albedo, g, radius;      /* single-scattering albedo, HG asymmetry, sphere radius in mean free paths (<=0: infinite) */
    int32_t max_depth, mode;       /* max collisions; mode 0 analog absorption, 1 implicit capture + roulette */
    int32_t rr_start, n_rbins;     /* roulette when collision index > rr_start; radial bins of the absorbed-weight tally */
    double rr_floor;
    uint64_t seed, first_photon;
} lto_medium_params;
typedef struct {
    double escaped, absorbed, truncated, rr_killed, rr_gained, sum_mu;   /* weights; sum of sampled scattering cosines */
    uint64_t n_collisions, n_segments, n_scatters;
} lto_medium_totals;
/* hist[k], k=0..max_depth: photons that made exactly k collisions; rbins[n_rbins]: absorbed weight by shell */
int lto_medium_walk(const lto_medium_params *P, uint64_t n_photons, uint64_t *hist, double *rbins, lto_medium_totals *tot,
                    int n_threads);
double lto_henyey_greenstein(double cos_theta, double g);

/* Philox4x32-10 known-answer access for the tests */
void lto_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
