/* lt_oracle.c — CPU oracle (float64, plain C).  TEST INFRASTRUCTURE ONLY; see lt_oracle.h.
 *
 * Each function restates one reference function; R/ = /root/reference/LightTransportSimulator/
 * light_transport/.  The reference keeps homogeneous 4-vectors (points w=1, directions w=0, Q11);
 * with those w values every 4-component dot/norm equals its 3-component counterpart, so 3-vectors
 * are used here.  Comparisons are written in the reference's form so NaN/inf behave alike.
 */
#include "lt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define INV_PI (1.0 / 3.14159265358979323846)       /* R/src/constants.py:7  */
#define PI_OVER_2 (3.14159265358979323846 / 2.0)    /* :10 */
#define PI_OVER_4 (0.5 * PI_OVER_2)                 /* :11 */

typedef struct { double o[3], d[3]; } ray_t;        /* R/src/rays.py:9-18 (tmax is always inf) */

typedef struct {
    const lto_scene *S;
    const lto_params *P;
    const double *r0, *r1;        /* this path's rand_0_set / rand_1_set, or NULL => Philox */
    const int32_t *lidx;          /* this path's replayed light picks or NULL */
    uint64_t path_id;
    uint32_t nee_count;
    int depth_frames;
    int error;
    lto_counters c;
} walk_t;

/* np.dot inside numba = cblas_ddot; OpenBLAS's scalar tail is an FMA chain (probed bit-for-bit on
 * 3000 random 4-vectors against numba 0.65 + scipy-openblas in the build container):
 * fma(a3,b3, fma(a2,b2, fma(a1,b1, a0*b0))), and the w term adds an exact 0 on this path. */
static inline double dot3(const double *a, const double *b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }
/* np.linalg.norm inside numba = cblas_dnrm2; OpenBLAS's x86-64 kernel accumulates in x87 extended
 * precision and rounds once (probed: equals the correctly rounded sqrt of the exact sum). */
static inline double norm3(const double *v) {
    long double s = (long double)v[0] * v[0] + (long double)v[1] * v[1] + (long double)v[2] * v[2];
    return (double)sqrtl(s);
}
static inline void cross3(const double *a, const double *b, double *r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
/* R/src/vectors.py:5-7 */
static inline void normalize3(double *v) {
    double n = norm3(v);
    v[0] /= n; v[1] /= n; v[2] /= n;
}

/* ---------------- Philox4x32-10 (Salmon et al. 2011), keyed (seed), counter (path, bounce, stream) */
void lto_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u53(uint32_t a, uint32_t b) {   /* numpy's random_double recipe */
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}
static void get_rand(const walk_t *w, int bounce, double *a, double *b) {
    if (w->r0) { *a = w->r0[bounce]; *b = w->r1[bounce]; return; }
    uint32_t ctr[4] = {(uint32_t)w->path_id, (uint32_t)(w->path_id >> 32), (uint32_t)bounce, 0u};
    uint32_t key[2] = {(uint32_t)w->P->seed, (uint32_t)(w->P->seed >> 32)}, o[4];
    lto_philox4x32_10(ctr, key, o);
    *a = u53(o[0], o[1]); *b = u53(o[2], o[3]);
}
static int pick_light(walk_t *w) {                   /* stands in for np.random.choice, Q12 */
    uint32_t k = w->nee_count++;
    int n = w->S->n_lights;
    if (w->lidx) {
        if ((int)k >= w->P->n_light_idx) { w->c.light_idx_overflow++; return 0; }
        return w->lidx[k];
    }
    if (n == 1) return 0;
    uint32_t ctr[4] = {(uint32_t)w->path_id, (uint32_t)(w->path_id >> 32), k, 1u};
    uint32_t key[2] = {(uint32_t)w->P->seed, (uint32_t)(w->P->seed >> 32)}, o[4];
    lto_philox4x32_10(ctr, key, o);
    return (int)(((uint64_t)o[0] * (uint64_t)n) >> 32);
}

/* ---------------- R/src/intersects.py:45-104  triangle_intersect (Moeller-Trumbore, two-sided) */
static int triangle_intersect(const lto_scene *S, int k, const double *o, const double *d, double *t_out) {
    const double eps = 0.0000001;
    const double *a = S->tri_v0 + 3 * k, *b = S->tri_v1 + 3 * k, *c = S->tri_v2 + 3 * k, *n = S->tri_n + 3 * k;
    double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    double ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    double ray_dot_plane = dot3(d, n);
    if (fabs(ray_dot_plane) <= eps) return 0;
    double pvec[3]; cross3(d, ac, pvec);
    double det = dot3(ab, pvec);
    if (-eps < det && det < eps) return 0;
    double inv_det = 1.0 / det;
    double tvec[3] = {o[0] - a[0], o[1] - a[1], o[2] - a[2]};
    double u = dot3(tvec, pvec) * inv_det;
    if (u < 0 || u > 1) return 0;
    double qvec[3]; cross3(tvec, ab, qvec);
    double v = dot3(d, qvec) * inv_det;
    if (v < 0 || u + v > 1) return 0;
    double t = dot3(ac, qvec) * inv_det;
    if (t > eps) { *t_out = t; return 1; }
    return 0;
}

/* ---------------- R/src/intersects.py:178-196 intersect_bounds, :228-235 gamma (float32 eps, Q10) */
static int intersect_bounds(const double *bmin, const double *bmax, const double *o, const double *inv_dir) {
    const double eps = 5.9604644775390625e-08;      /* np.finfo(np.float32).eps * 0.5 */
    const double gamma3 = (3 * eps) / (1 - 3 * eps);
    double t0 = 0, t1 = INFINITY;                   /* ray.tmax is always inf on this path */
    for (int i = 0; i < 3; ++i) {
        double t_near = (bmin[i] - o[i]) * inv_dir[i];
        double t_far = (bmax[i] - o[i]) * inv_dir[i];
        if (t_near > t_far) { double s = t_near; t_near = t_far; t_far = s; }
        t_far *= 1 + 2 * gamma3;
        t0 = t_near > t0 ? t_near : t0;
        t1 = t_far < t1 ? t_far : t1;
        if (t0 > t1) return 0;
    }
    return 1;
}

/* ---------------- R/src/bvh_new.py:413-482 intersect_bvh, literally (Q8/Q9/Q9b) */
static int intersect_bvh(const lto_scene *S, const ray_t *ray, double eps, int *prim_out, double *t_out,
                         lto_counters *c) {
    int n = S->n_nodes, current = 0, triangle = -1;
    unsigned char visited_stack[512];
    unsigned char *visited = n <= 512 ? visited_stack : (unsigned char *)malloc((size_t)n);
    memset(visited, 0, (size_t)n);
    double inv_dir[3] = {1 / ray->d[0], 1 / ray->d[1], 1 / ray->d[2]};
    int dir_is_neg[3] = {inv_dir[0] < 0, inv_dir[1] < 0, inv_dir[2] < 0};
    double min_distance = INFINITY;
    int rc = 0;
    for (;;) {
        int rescan = 0;
        if (!visited[current]) {
            visited[current] = 1;
            c->box_tests++;
            if (intersect_bounds(S->node_bmin + 3 * current, S->node_bmax + 3 * current, ray->o, inv_dir)) {
                if (S->node_n_prims[current] > 0) {
                    for (int i = 0; i < S->node_n_prims[current]; ++i) {
                        int leaf_idx = S->node_prim_off[current] + i;
                        if (leaf_idx >= n) { rc = -1; goto done; }      /* IndexError in the reference */
                        visited[leaf_idx] = 1;                           /* :436, wrong list on purpose */
                        double t;
                        c->tri_tests++;
                        if (triangle_intersect(S, leaf_idx, ray->o, ray->d, &t) && eps < t && t < min_distance) {
                            min_distance = t;
                            triangle = leaf_idx;
                        }
                    }
                    if (current == 0) break;
                    rescan = 1;
                } else {
                    current = dir_is_neg[S->node_axis[current]] ? S->node_second[current] : current + 1;
                }
            } else {
                if (current == 0) break;
                rescan = 1;
            }
        } else {
            rescan = 1;
        }
        if (rescan) {                        /* :445-449 — no break in the scan: LAST unvisited wins */
            int all_visited = 1;
            for (int i = 0; i < n; ++i)
                if (!visited[i]) { all_visited = 0; current = i; }
            if (all_visited) break;
        }
    }
done:
    if (visited != visited_stack) free(visited);
    *prim_out = triangle;
    *t_out = min_distance;
    return rc;
}

/* NEAREST semantics = R/src/utils.py:34-46 intent: closest hit over all primitives, EPSILON < t,
 * first (lowest index) wins ties. */
static void intersect_nearest(const lto_scene *S, const ray_t *ray, double eps, int *prim_out, double *t_out,
                              lto_counters *c) {
    int best = -1;
    double bt = INFINITY;
    for (int k = 0; k < S->n_tris; ++k) {
        double t;
        c->tri_tests++;
        if (triangle_intersect(S, k, ray->o, ray->d, &t) && eps < t && t < bt) { bt = t; best = k; }
    }
    *prim_out = best;
    *t_out = bt;
}

static int closest_hit(walk_t *w, const ray_t *ray, int *prim, double *t) {
    if (w->P->traversal == 0) {
        if (intersect_bvh(w->S, ray, w->P->eps, prim, t, &w->c)) { w->error = 1; *prim = -1; }
    } else {
        intersect_nearest(w->S, ray, w->P->eps, prim, t, &w->c);
    }
    return *prim >= 0;
}

/* ---------------- R/src/utils.py:114-128 concentric_sample_disk, :71-80 create_orthonormal_system,
 *                  :131-161 cosine_weighted_hemisphere_sampling (Q3: world-z flip test) */
static double cosine_weighted_hemisphere_sampling(const double *n, const double *d_in, double u0, double u1,
                                                  double *out) {
    double inc_z = -d_in[2];
    double ox = 2.0 * u0 - 1, oy = 2.0 * u1 - 1, dx, dy;
    if (ox == 0 && oy == 0) { dx = 0; dy = 0; }
    else {
        double r, theta;
        if (fabs(ox) > fabs(oy)) { r = ox; theta = PI_OVER_4 * (oy / ox); }
        else { r = oy; theta = PI_OVER_2 - PI_OVER_4 * (ox / oy); }
        dx = r * cos(theta); dy = r * sin(theta);
    }
    double zz = 1 - dx * dx - dy * dy;
    double z = sqrt(zz > 0 ? zz : 0);
    double v2[3], v3[3];
    if (fabs(n[0]) > fabs(n[1])) {
        double s = sqrt(n[0] * n[0] + n[2] * n[2]);
        v2[0] = -n[2] / s; v2[1] = 0.0 / s; v2[2] = n[0] / s;
    } else {
        double s = sqrt(n[1] * n[1] + n[2] * n[2]);
        v2[0] = 0.0 / s; v2[1] = n[2] / s; v2[2] = -n[1] / s;
    }
    cross3(n, v2, v3);
    double oz = z;
    if (inc_z < 0) oz *= -1;
    double pdf = (inc_z * oz > 0) ? fabs(z) * INV_PI : 0;
    for (int c = 0; c < 3; ++c) out[c] = dx * v2[c] + dy * v3[c] + oz * n[c];
    return pdf;
}

/* ---------------- R/src/light_samples.py:35-61 cast_one_shadow_ray */
static void cast_one_shadow_ray(walk_t *w, int prim, const double *x, const double *n, double *contrib) {
    const lto_scene *S = w->S;
    contrib[0] = contrib[1] = contrib[2] = 0;
    int k = pick_light(w);
    const double *lp = S->light_pos + 3 * k;
    ray_t sr;
    double dv[3] = {lp[0] - x[0], lp[1] - x[1], lp[2] - x[2]};
    double mag = norm3(dv);
    for (int c = 0; c < 3; ++c) { sr.o[c] = x[c]; sr.d[c] = dv[c] / mag; }
    int hp; double min_distance;
    w->c.rays_shadow++;
    closest_hit(w, &sr, &hp, &min_distance);
    if (min_distance >= mag - w->P->eps) {
        const double *kd = S->mat_kd + 3 * S->tri_mat[prim];
        const double *lekd = S->light_lekd + 3 * k, *ln = S->light_n + 3 * k;
        double cos_theta = dot3(n, sr.d);
        double nd[3] = {-sr.d[0], -sr.d[1], -sr.d[2]};
        double cos_phi = dot3(ln, nd);
        double g = fabs(cos_theta * cos_phi) / (mag * mag);
        for (int c = 0; c < 3; ++c) contrib[c] += lekd[c] * (kd[c] * INV_PI) * g * S->light_area[k];
    }
}

/* R/src/brdf.py:7-9 */
static void reflected(const double *v, const double *n, double *r) {
    double k = 2 * dot3(v, n);
    for (int c = 0; c < 3; ++c) r[c] = v[c] - k * n[c];
    normalize3(r);
}

/* specular part shared by both integrators: R/src/path_tracing.py:103-141 / path_tracing_fix1.py:81-119.
 * returns 0 when the material is none of diffuse/mirror/transmissive ("error" branch). */
static int specular_bounce(const lto_scene *S, int m, int inside, const double *p, const double *n, double rand_0,
                           double eps, ray_t *ray) {
    if (S->mat_flags[m] & 2) {
        double r[3]; reflected(ray->d, n, r);
        for (int c = 0; c < 3; ++c) { ray->o[c] = p[c] + eps * n[c]; ray->d[c] = r[c]; }
        return 1;
    }
    if (S->mat_transmission[m] > 0.0) {
        double ior = S->mat_ior[m];
        double n1 = inside ? ior : 1, n2 = inside ? 1 : ior;
        double q = (n1 - n2) / (n1 + n2), R0 = q * q;
        double theta = dot3(ray->d, n);
        double om = 1 - cos(theta);                    /* Q5: cosine of a cosine */
        double om2 = om * om;
        double reflection_prob = R0 + (1 - R0) * (om2 * om2 * om);
        double Nr = ior;
        if (dot3(ray->d, n) > 0) Nr = 1 / Nr;          /* never true after the flip (Q5) */
        Nr = 1 / Nr;
        double cos_theta = -(dot3(ray->d, n));
        double _sqrt = 1 - (Nr * Nr) * (1 - cos_theta * cos_theta);
        if (_sqrt > 0 && rand_0 > reflection_prob) {
            double k = Nr * cos_theta - sqrt(_sqrt), td[3];
            for (int c = 0; c < 3; ++c) td[c] = ray->d[c] * Nr + n[c] * k;
            normalize3(td);
            for (int c = 0; c < 3; ++c) { ray->o[c] = p[c] + (-eps * n[c]); ray->d[c] = td[c]; }
        } else {
            double r[3]; reflected(ray->d, n, r);
            for (int c = 0; c < 3; ++c) { ray->o[c] = p[c] + eps * n[c]; ray->d[c] = r[c]; }
        }
        return 1;
    }
    return 0;
}

/* ---------------- R/src/path_tracing.py:23-159 trace_path (Q1: recurse AND continue) */
static void trace_path_reference(walk_t *w, ray_t *ray, int bounce, double *light) {
    const lto_scene *S = w->S;
    const lto_params *P = w->P;
    double T[3] = {1, 1, 1};
    light[0] = light[1] = light[2] = 0;
    if (++w->depth_frames > (int)w->c.max_frames) w->c.max_frames = (uint64_t)w->depth_frames;
    for (;;) {
        if (bounce >= P->max_depth) break;
        double rand_0, rand_1;
        get_rand(w, bounce, &rand_0, &rand_1);
        int prim; double t;
        w->c.rays_ext++;
        if (!closest_hit(w, ray, &prim, &t)) break;                    /* R/src/utils.py:52-68 */
        double p[3] = {ray->o[0] + t * ray->d[0], ray->o[1] + t * ray->d[1], ray->o[2] + t * ray->d[2]};
        int m = S->tri_mat[prim];
        double n[3] = {S->tri_n[3 * prim], S->tri_n[3 * prim + 1], S->tri_n[3 * prim + 2]};
        if (S->tri_light[prim] && bounce == 0)
            for (int c = 0; c < 3; ++c) light[c] += S->mat_emission[m] * T[c];
        int inside = 0;
        if (dot3(n, ray->d) > 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; inside = 1; }
        if (S->mat_flags[m] & 1) {
            double so[3] = {p[0] + P->eps * n[0], p[1] + P->eps * n[1], p[2] + P->eps * n[2]};
            double direct[3], dir[3], tp[3];
            cast_one_shadow_ray(w, prim, so, n, direct);
            double pdf = cosine_weighted_hemisphere_sampling(n, ray->d, rand_0, rand_1, dir);
            if (pdf == 0) break;
            for (int c = 0; c < 3; ++c) { ray->o[c] = p[c] + P->eps * dir[c]; ray->d[c] = dir[c]; }
            double cos_theta = dot3(dir, n);
            for (int c = 0; c < 3; ++c) T[c] *= (S->mat_kd[3 * m + c] * INV_PI) * cos_theta / pdf;
            trace_path_reference(w, ray, bounce + 1, tp);
            for (int c = 0; c < 3; ++c) light[c] += direct[c] + T[c] * tp[c];
        } else if (!specular_bounce(S, m, inside, p, n, rand_0, P->eps, ray)) {
            break;
        }
        if (bounce > P->rr_start) {
            double r_r = 1 - T[1] > P->rr_floor ? 1 - T[1] : P->rr_floor;
            if (rand_0 < r_r) break;
            for (int c = 0; c < 3; ++c) T[c] /= 1 - r_r;
        }
        bounce += 1;
    }
    w->depth_frames--;
}

static void add_counters(lto_counters *a, const lto_counters *b);

/* ---------------- R/src/path_tracing.py:23-159 again, now with every side output
 * and the calculate_grad gate (rank 2).  `rlp` is the frame's record_log_pdf array (returned value),
 * dl/il the frame's direct/indirect lists [3][D] (may be NULL below the root: the caller drops them). */
typedef struct {
    double *scene_rlp;      /* scene.record_log_pdf[i,j,s,:]  (written when !grad) */
    int8_t *bounce_record;  /* scene.bounce_record[i,j,s,:]   (written when !grad, read when grad) */
    const double *r0, *r1;  /* rand sets of THIS replay (may differ from the walk's base rows) */
    int grad;               /* a gradient replay: no side effects on the scene arrays */
    int gate;               /* replays stop where bounce_record is 0 (calculate_grad=True, :37-39) */
} rec_t;

static void trace_path_reference_rec(walk_t *w, const rec_t *rc, ray_t *ray, int bounce, double *light, double *rlp,
                                     double *dl, double *il) {
    const lto_scene *S = w->S;
    const lto_params *P = w->P;
    const int D = P->max_depth;
    double T[3] = {1, 1, 1};
    light[0] = light[1] = light[2] = 0;
    for (int b = 0; b < D; ++b) rlp[b] = 0;
    if (dl) for (int k = 0; k < 3 * D; ++k) { dl[k] = 0; il[k] = 0; }
    for (;;) {
        /* :33-34 read rand_set[bounce] before any test (index D is read but unused there) */
        if (rc->grad && rc->gate && bounce < D && rc->bounce_record[bounce] == 0) break;       /* :37-39 (calculate_grad is True) */
        if (bounce >= D) break;
        double rand_0 = rc->r0[bounce], rand_1 = rc->r1[bounce];
        int prim; double t;
        w->c.rays_ext++;
        if (!closest_hit(w, ray, &prim, &t)) {
            if (!rc->grad) for (int b = bounce; b < D; ++b) rc->bounce_record[b] = 0;   /* :50-52 */
            break;
        }
        double p[3] = {ray->o[0] + t * ray->d[0], ray->o[1] + t * ray->d[1], ray->o[2] + t * ray->d[2]};
        int m = S->tri_mat[prim];
        double n[3] = {S->tri_n[3 * prim], S->tri_n[3 * prim + 1], S->tri_n[3 * prim + 2]};
        if (S->tri_light[prim] && bounce == 0)
            for (int c = 0; c < 3; ++c) light[c] += S->mat_emission[m] * T[c];
        int inside = 0;
        if (dot3(n, ray->d) > 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; inside = 1; }
        if (S->mat_flags[m] & 1) {
            double so[3] = {p[0] + P->eps * n[0], p[1] + P->eps * n[1], p[2] + P->eps * n[2]};
            double direct[3], dir[3], tp[3];
            cast_one_shadow_ray(w, prim, so, n, direct);
            double pdf = cosine_weighted_hemisphere_sampling(n, ray->d, rand_0, rand_1, dir);
            if (pdf == 0) {
                if (!rc->grad) for (int b = bounce + 1; b < D; ++b) rc->bounce_record[b] = 0;  /* :77-79 */
                break;
            }
            for (int c = 0; c < 3; ++c) { ray->o[c] = p[c] + P->eps * dir[c]; ray->d[c] = dir[c]; }
            double cos_theta = dot3(dir, n);
            for (int c = 0; c < 3; ++c) T[c] *= (S->mat_kd[3 * m + c] * INV_PI) * cos_theta / pdf;
            double child_rlp[D > 0 ? D : 1];
            trace_path_reference_rec(w, rc, ray, bounce + 1, tp, child_rlp, 0, 0);
            for (int b = 0; b < D; ++b) rlp[b] = child_rlp[b];          /* :93 the callee's array REPLACES ours */
            rlp[bounce] = log(pdf);                                      /* :94 */
            if (!rc->grad) rc->scene_rlp[bounce] = log(pdf);             /* :95-96 */
            if (dl) for (int c = 0; c < 3; ++c) { dl[c * D + bounce] = direct[c]; il[c * D + bounce] = T[c] * tp[c]; }
            for (int c = 0; c < 3; ++c) light[c] += direct[c] + T[c] * tp[c];
        } else if (!specular_bounce(S, m, inside, p, n, rand_0, P->eps, ray)) {
            break;
        }
        if (bounce > P->rr_start) {
            double r_r = 1 - T[1] > P->rr_floor ? 1 - T[1] : P->rr_floor;
            if (rand_0 < r_r) {
                if (!rc->grad) for (int b = bounce + 1; b < D; ++b) rc->bounce_record[b] = 0;  /* :151-153 */
                break;
            }
            for (int c = 0; c < 3; ++c) T[c] /= 1 - r_r;
        }
        bounce += 1;
    }
}

/* R/src/path_tracing.py:19-21 (Q16: this is sigma(-x)) */
static inline double sigmoid_ref(double x) { return 1 / (1 + exp(x)); }

/* render_scene_samples' per-sample body (R/src/path_tracing.py:270-303) for replayed random numbers:
 * the base walk with its side outputs, then calculate_gradients (:203-249) when minus/plus are given. */
int lto_render_records(const lto_scene *S, const lto_params *P, const double *rand0, const double *rand1,
                       const double *logit0, const double *logit1, const int32_t *light_idx, double *samples_out,
                       double *record_log_pdf, double *scene_record_log_pdf, int8_t *bounce_record, double *direct_list,
                       double *indirect_list, double *rlp_minus, double *rlp_plus, lto_counters *cnt) {
    const int W = P->width, H = P->height, D = P->max_depth, spp = P->spp;
    if (!rand0 || !rand1 || P->s0 != 0 || P->s1 != spp || D <= 0) return -2;
    const double aspect = (double)W / (double)H, top = 1 / aspect, bottom = -1 / aspect;
    const double step_y = H > 1 ? (bottom - top) / (double)(H - 1) : 0;
    const double step_x = W > 1 ? (1.0 - (-1.0)) / (double)(W - 1) : 0;
    lto_counters total;
    memset(&total, 0, sizeof total);
    int err = 0;
#pragma omp parallel
    {
        walk_t w;
        memset(&w, 0, sizeof w);
        w.S = S; w.P = P;
        double *r0s = (double *)malloc(sizeof(double) * 2 * (size_t)D), *r1s = r0s + D;
        const int64_t pix_lo = (P->pixel0 == 0 && P->pixel1 == 0) ? 0 : P->pixel0, pix_hi = (P->pixel0 == 0 && P->pixel1 == 0) ? (int64_t)W * H : P->pixel1;
#pragma omp for schedule(dynamic, 16)
        for (int64_t pix = pix_lo; pix < pix_hi; ++pix) {
            int i = (int)(pix / W), j = (int)(pix % W);
            double y = (H > 1 && i == H - 1) ? bottom : top + i * step_y;
            double x = (W > 1 && j == W - 1) ? 1.0 : -1.0 + j * step_x;
            for (int s = 0; s < spp; ++s) {
                uint64_t path = (uint64_t)(pix - pix_lo) * (uint64_t)spp + (uint64_t)s;      /* row of the (sub-range) arrays */
                const double *b0 = rand0 + path * (uint64_t)D, *b1 = rand1 + path * (uint64_t)D;
                w.path_id = (uint64_t)pix * (uint64_t)spp + (uint64_t)s;
                w.r0 = b0; w.r1 = b1;
                w.lidx = light_idx ? light_idx + path * (uint64_t)P->n_light_idx : 0;
                w.nee_count = 0;
                ray_t cam_ray, ray;
                double end[3] = {x + b0[0] / W, y + b0[0] / H, P->f_distance};
                for (int c = 0; c < 3; ++c) { cam_ray.o[c] = P->camera[c]; cam_ray.d[c] = end[c] - P->camera[c]; }
                normalize3(cam_ray.d);
                rec_t rc;
                rc.scene_rlp = scene_record_log_pdf + path * (uint64_t)D;
                rc.bounce_record = bounce_record + path * (uint64_t)D;
                for (int b = 0; b < D; ++b) { rc.scene_rlp[b] = 0; rc.bounce_record[b] = 1; }   /* R/src/scene.py:72-73 */
                rc.r0 = b0; rc.r1 = b1; rc.grad = 0; rc.gate = P->grad_ungated ? 0 : 1;
                double L[3];
                ray = cam_ray;
                trace_path_reference_rec(&w, &rc, &ray, 0, L, record_log_pdf + path * (uint64_t)D,
                                         direct_list ? direct_list + path * 3 * (uint64_t)D : 0,
                                         indirect_list ? indirect_list + path * 3 * (uint64_t)D : 0);
                w.c.paths++;
                if (samples_out) for (int c = 0; c < 3; ++c) samples_out[path * 3 + c] = L[c];
                if (rlp_minus && rlp_plus) {                      /* calculate_gradients, :203-249 */
                    const double *l0 = logit0 + path * (uint64_t)D, *l1 = logit1 + path * (uint64_t)D;
                    rc.grad = 1;
                    for (int which = 0; which < 2; ++which)
                        for (int ri = 0; ri < D; ++ri)
                            for (int sign = 0; sign < 2; ++sign) {
                                for (int b = 0; b < D; ++b) { r0s[b] = sigmoid_ref(l0[b]); r1s[b] = sigmoid_ref(l1[b]); }
                                const double sstep = 0.01;
                                double *tgt = which == 0 ? r0s : r1s;
                                const double *lg = which == 0 ? l0 : l1;
                                double pert = lg[ri];
                                if (sign == 0) pert -= sstep; else pert += sstep;    /* logits[rand_i] -= / += _s */
                                tgt[ri] = sigmoid_ref(pert);
                                rc.r0 = r0s; rc.r1 = r1s;
                                ray = cam_ray;
                                double Lg[3];
                                double *out = (sign == 0 ? rlp_minus : rlp_plus) + (path * 2 * (uint64_t)D + (uint64_t)(which * D + ri)) * (uint64_t)D;
                                trace_path_reference_rec(&w, &rc, &ray, 0, Lg, out, 0, 0);
                            }
                }
            }
        }
        free(r0s);
#pragma omp critical
        {
            add_counters(&total, &w.c);
            if (w.error) err = -1;
        }
    }
    if (cnt) *cnt = total;
    return err;
}

/* ---------------- R/src/path_tracing_fix1.py:17-136 trace_path (iterative reformulation) */
static void trace_path_iterative(walk_t *w, ray_t *ray, int bounce, double *light) {
    const lto_scene *S = w->S;
    const lto_params *P = w->P;
    double T[3] = {1, 1, 1};
    light[0] = light[1] = light[2] = 0;
    w->c.max_frames = 1;
    for (;;) {
        if (bounce >= P->max_depth) break;
        double rand_0, rand_1;
        get_rand(w, bounce, &rand_0, &rand_1);
        int prim; double t;
        w->c.rays_ext++;
        if (!closest_hit(w, ray, &prim, &t)) break;
        double p[3] = {ray->o[0] + t * ray->d[0], ray->o[1] + t * ray->d[1], ray->o[2] + t * ray->d[2]};
        int m = S->tri_mat[prim];
        double n[3] = {S->tri_n[3 * prim], S->tri_n[3 * prim + 1], S->tri_n[3 * prim + 2]};
        if (S->tri_light[prim])
            for (int c = 0; c < 3; ++c) light[c] += S->mat_emission[m] * T[c];
        int inside = 0;
        if (dot3(n, ray->d) > 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; inside = 1; }
        if (S->mat_flags[m] & 1) {
            double so[3] = {p[0] + P->eps * n[0], p[1] + P->eps * n[1], p[2] + P->eps * n[2]};
            double direct[3], dir[3];
            cast_one_shadow_ray(w, prim, so, n, direct);
            for (int c = 0; c < 3; ++c) light[c] += T[c] * direct[c];
            double pdf = cosine_weighted_hemisphere_sampling(n, ray->d, rand_0, rand_1, dir);
            if (pdf == 0) break;
            double cos_theta = dot3(dir, n);
            for (int c = 0; c < 3; ++c) T[c] *= (S->mat_kd[3 * m + c] * INV_PI) * cos_theta / pdf;
            for (int c = 0; c < 3; ++c) { ray->o[c] = p[c] + P->eps * dir[c]; ray->d[c] = dir[c]; }
        } else if (!specular_bounce(S, m, inside, p, n, rand_0, P->eps, ray)) {
            break;
        }
        if (bounce > P->rr_start) {
            double r_r = 1 - T[1] > P->rr_floor ? 1 - T[1] : P->rr_floor;
            if (rand_0 < r_r) break;
            for (int c = 0; c < 3; ++c) T[c] /= 1 - r_r;
        }
        bounce += 1;
    }
}

static void add_counters(lto_counters *a, const lto_counters *b) {
    a->paths += b->paths; a->rays_ext += b->rays_ext; a->rays_shadow += b->rays_shadow;
    a->box_tests += b->box_tests; a->tri_tests += b->tri_tests;
    a->light_idx_overflow += b->light_idx_overflow; a->errors += b->errors;
    if (b->max_frames > a->max_frames) a->max_frames = b->max_frames;
}

int lto_intersect(const lto_scene *S, int traversal, const double *o, const double *d, int64_t n, int32_t *prim_out,
                  double *t_out, lto_counters *cnt) {
    int rc = 0;
    lto_counters total;
    memset(&total, 0, sizeof total);
#pragma omp parallel
    {
        lto_counters c;
        memset(&c, 0, sizeof c);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n; ++r) {
            ray_t ray;
            for (int k = 0; k < 3; ++k) { ray.o[k] = o[3 * r + k]; ray.d[k] = d[3 * r + k]; }
            int prim; double t;
            if (traversal == 0) {
                if (intersect_bvh(S, &ray, 0.000001, &prim, &t, &c)) {
#pragma omp atomic write
                    rc = -1;
                }
            } else {
                intersect_nearest(S, &ray, 0.000001, &prim, &t, &c);
            }
            prim_out[r] = prim;
            t_out[r] = t;
        }
#pragma omp critical
        add_counters(&total, &c);
    }
    if (cnt) *cnt = total;
    return rc;
}

/* ---------------- R/src/path_tracing.py:263-305: ray generation (a1) and the film tally (a10) */
int lto_render(const lto_scene *S, const lto_params *P, const double *rand0, const double *rand1,
               const int32_t *light_idx, double *film_sum, double *samples_out, lto_counters *cnt) {
    const int W = P->width, H = P->height, D = P->max_depth, spp = P->spp, ns = P->s1 - P->s0;
    /* numba's np.linspace (numba/np/arrayobj.py): arr[i] = start + i*step, arr[-1] = stop */
    const double aspect = (double)W / (double)H, top = 1 / aspect, bottom = -1 / aspect;
    const double step_y = H > 1 ? (bottom - top) / (double)(H - 1) : 0;
    const double step_x = W > 1 ? (1.0 - (-1.0)) / (double)(W - 1) : 0;
    lto_counters total;
    memset(&total, 0, sizeof total);
    int err = 0;
#ifdef _OPENMP
    if (P->n_threads > 0) omp_set_num_threads(P->n_threads);
#endif
#pragma omp parallel
    {
        walk_t w;
        memset(&w, 0, sizeof w);
        w.S = S; w.P = P;
#pragma omp for schedule(dynamic, 16)
        for (int64_t pix = 0; pix < (int64_t)W * H; ++pix) {
            int i = (int)(pix / W), j = (int)(pix % W);
            double y = (H > 1 && i == H - 1) ? bottom : top + i * step_y;
            double x = (W > 1 && j == W - 1) ? 1.0 : -1.0 + j * step_x;
            double color[3] = {0, 0, 0};
            for (int s = P->s0; s < P->s1; ++s) {
                uint64_t path = (uint64_t)pix * (uint64_t)spp + (uint64_t)s;
                w.path_id = path;
                w.r0 = rand0 ? rand0 + path * (uint64_t)D : 0;
                w.r1 = rand1 ? rand1 + path * (uint64_t)D : 0;
                w.lidx = light_idx ? light_idx + path * (uint64_t)P->n_light_idx : 0;
                w.nee_count = 0;
                w.depth_frames = 0;
                double u0, u1;
                get_rand(&w, 0, &u0, &u1);
                ray_t ray;
                double end[3] = {x + u0 / W, y + u0 / H, P->f_distance};      /* Q13: same number twice */
                for (int c = 0; c < 3; ++c) { ray.o[c] = P->camera[c]; ray.d[c] = end[c] - P->camera[c]; }
                normalize3(ray.d);
                double L[3];
                if (P->semantics == 0) trace_path_reference(&w, &ray, 0, L);
                else trace_path_iterative(&w, &ray, 0, L);
                w.c.paths++;
                for (int c = 0; c < 3; ++c) color[c] += L[c];
                if (samples_out)
                    for (int c = 0; c < 3; ++c) samples_out[((uint64_t)pix * ns + (s - P->s0)) * 3 + c] = L[c];
            }
            if (film_sum)
                for (int c = 0; c < 3; ++c) film_sum[pix * 3 + c] += color[c];
        }
#pragma omp critical
        {
            add_counters(&total, &w.c);
            if (w.error) err = -1;
        }
    }
    if (cnt) *cnt = total;
    return err;
}

/* Per-pixel statistics of a pixel SUBSET of a (possibly huge) film: mean[n][3] and m2[n][3] = sum of squares of the
 * per-sample radiance over samples [s0, s1) of each listed pixel (row-major pixel indices of the W x H film).  Same
 * ray generation and walk as lto_render (R/src/path_tracing.py:263-305), Philox streams only.  Lets the tests
 * compare single pixels of config C2 (1024 x 1024 x 1024 spp) without a whole-film array. */
int lto_render_pixels(const lto_scene *S, const lto_params *P, const int32_t *pixels, int32_t n_pixels, double *sum,
                      double *sumsq, lto_counters *cnt) {
    const int W = P->width, H = P->height, spp = P->spp;
    const double aspect = (double)W / (double)H, top = 1 / aspect, bottom = -1 / aspect;
    const double step_y = H > 1 ? (bottom - top) / (double)(H - 1) : 0;
    const double step_x = W > 1 ? (1.0 - (-1.0)) / (double)(W - 1) : 0;
    lto_counters total;
    memset(&total, 0, sizeof total);
    int err = 0;
#ifdef _OPENMP
    if (P->n_threads > 0) omp_set_num_threads(P->n_threads);
#endif
#pragma omp parallel
    {
        walk_t w;
        memset(&w, 0, sizeof w);
        w.S = S; w.P = P;
#pragma omp for schedule(dynamic, 1)
        for (int32_t k = 0; k < n_pixels; ++k) {
            const int64_t pix = pixels[k];
            int i = (int)(pix / W), j = (int)(pix % W);
            double y = (H > 1 && i == H - 1) ? bottom : top + i * step_y;
            double x = (W > 1 && j == W - 1) ? 1.0 : -1.0 + j * step_x;
            double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
            for (int s = P->s0; s < P->s1; ++s) {
                w.path_id = (uint64_t)pix * (uint64_t)spp + (uint64_t)s;
                w.r0 = 0; w.r1 = 0; w.lidx = 0;
                w.nee_count = 0;
                w.depth_frames = 0;
                double u0, u1;
                get_rand(&w, 0, &u0, &u1);
                ray_t ray;
                double end[3] = {x + u0 / W, y + u0 / H, P->f_distance};
                for (int c = 0; c < 3; ++c) { ray.o[c] = P->camera[c]; ray.d[c] = end[c] - P->camera[c]; }
                normalize3(ray.d);
                double L[3];
                if (P->semantics == 0) trace_path_reference(&w, &ray, 0, L);
                else trace_path_iterative(&w, &ray, 0, L);
                w.c.paths++;
                for (int c = 0; c < 3; ++c) { a[c] += L[c]; b[c] += L[c] * L[c]; }
            }
            for (int c = 0; c < 3; ++c) { sum[3 * k + c] = a[c]; sumsq[3 * k + c] = b[c]; }
        }
#pragma omp critical
        {
            add_counters(&total, &w.c);
            if (w.error) err = -1;
        }
    }
    if (cnt) *cnt = total;
    return err;
}


/* ================================================================================================
 * C4b — photon random walk in a homogeneous medium (sigma_t = 1): free flight, collision, absorption
 * (analog or implicit capture + the reference's roulette rule R/src/path_tracing.py:148-155 applied to
 * the photon weight), Henyey-Greenstein scattering.  Philox4x32-10: key = seed, counter =
 * (photon lo, photon hi, collision index, stream 2 | 3).
 * ================================================================================================ */
/* R/src/medium_samples.py:14-16 */
double lto_henyey_greenstein(double cosTheta, double g) {
    const double inv_4_pi = 0.25 * INV_PI;
    double denom = 1 + g * g + 2 * g * cosTheta;
    return inv_4_pi * (1 - g * g) / (denom * sqrt(denom));
}

static void medium_rand4(const lto_medium_params *P, uint64_t photon, uint32_t k, double u[4]) {
    uint32_t key[2] = {(uint32_t)P->seed, (uint32_t)(P->seed >> 32)}, o[4];
    uint32_t c2[4] = {(uint32_t)photon, (uint32_t)(photon >> 32), k, 2u};
    lto_philox4x32_10(c2, key, o);
    u[0] = u53(o[0], o[1]); u[1] = u53(o[2], o[3]);
    uint32_t c3[4] = {(uint32_t)photon, (uint32_t)(photon >> 32), k, 3u};
    lto_philox4x32_10(c3, key, o);
    u[2] = u53(o[0], o[1]); u[3] = u53(o[2], o[3]);
}

int lto_medium_walk(const lto_medium_params *P, uint64_t n_photons, uint64_t *hist, double *rbins, lto_medium_totals *tot,
                    int n_threads) {
    const int D = P->max_depth, NB = P->n_rbins;
    const double a = P->albedo, g = P->g, Rs = P->radius;
    lto_medium_totals T;
    memset(&T, 0, sizeof T);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        lto_medium_totals t;
        memset(&t, 0, sizeof t);
        uint64_t *h = (uint64_t *)calloc((size_t)D + 1, sizeof(uint64_t));
        double *rb = (double *)calloc((size_t)(NB > 0 ? NB : 1), sizeof(double));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n_photons; ++i) {
            const uint64_t id = P->first_photon + (uint64_t)i;
            double x[3] = {0, 0, 0}, d[3], w = 1.0, u[4];
            medium_rand4(P, id, 0xffffffffu, u);                /* isotropic point source at the centre */
            {
                double mu = 1 - 2 * u[0], sn = sqrt(1 - mu * mu > 0 ? 1 - mu * mu : 0), ph = 2 * 3.14159265358979323846 * u[1];
                d[0] = sn * cos(ph); d[1] = sn * sin(ph); d[2] = mu;
            }
            int k = 0;                                          /* collisions so far */
            for (;;) {
                medium_rand4(P, id, (uint32_t)k, u);
                double s = -log(1.0 - u[0]);                    /* free path, sigma_t = 1 */
                t.n_segments++;
                for (int c = 0; c < 3; ++c) x[c] += s * d[c];
                double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
                if (Rs > 0 && r2 >= Rs * Rs) { t.escaped += w; break; }
                k++;
                t.n_collisions++;
                int shell = NB > 0 ? (Rs > 0 ? (int)(sqrt(r2) / Rs * NB) : 0) : -1;
                if (shell >= NB) shell = NB - 1;
                if (P->mode == 0) {                             /* analog: absorbed with probability 1 - albedo */
                    if (u[1] >= a) { t.absorbed += w; if (shell >= 0) rb[shell] += w; break; }
                } else {                                        /* implicit capture, then roulette on the weight */
                    double dep = w * (1 - a);
                    t.absorbed += dep;
                    if (shell >= 0) rb[shell] += dep;
                    w *= a;
                    if (k - 1 > P->rr_start) {
                        double q = 1 - w > P->rr_floor ? 1 - w : P->rr_floor;
                        if (u[1] < q) { t.rr_killed += w; break; }
                        t.rr_gained += w / (1 - q) - w;
                        w /= 1 - q;
                    }
                }
                if (k >= D) { t.truncated += w; break; }
                /* Henyey-Greenstein scattering about the flight direction */
                double mu;
                if (fabs(g) < 1e-3) mu = 1 - 2 * u[2];
                else { double sq = (1 - g * g) / (1 - g + 2 * g * u[2]); mu = (1 + g * g - sq * sq) / (2 * g); }
                if (mu > 1) mu = 1;
                if (mu < -1) mu = -1;
                t.sum_mu += mu;
                t.n_scatters++;
                double sn = sqrt(1 - mu * mu), ph = 2 * 3.14159265358979323846 * u[3], cp = cos(ph), sp = sin(ph);
                double nd[3];
                if (fabs(d[2]) > 0.99999) {
                    nd[0] = sn * cp; nd[1] = sn * sp; nd[2] = d[2] > 0 ? mu : -mu;
                } else {
                    double den = sqrt(1 - d[2] * d[2]);
                    nd[0] = sn * (d[0] * d[2] * cp - d[1] * sp) / den + d[0] * mu;
                    nd[1] = sn * (d[1] * d[2] * cp + d[0] * sp) / den + d[1] * mu;
                    nd[2] = -sn * cp * den + d[2] * mu;
                }
                double nn = sqrt(nd[0] * nd[0] + nd[1] * nd[1] + nd[2] * nd[2]);
                for (int c = 0; c < 3; ++c) d[c] = nd[c] / nn;
            }
            h[k < D ? k : D]++;
        }
#pragma omp critical
        {
            T.escaped += t.escaped; T.absorbed += t.absorbed; T.truncated += t.truncated; T.rr_killed += t.rr_killed;
            T.rr_gained += t.rr_gained; T.sum_mu += t.sum_mu; T.n_collisions += t.n_collisions; T.n_segments += t.n_segments; T.n_scatters += t.n_scatters;
            for (int k = 0; k <= D; ++k) hist[k] += h[k];
            for (int b = 0; b < NB; ++b) rbins[b] += rb[b];
        }
        free(h);
        free(rb);
    }
    if (tot) *tot = T;
    return 0;
}
