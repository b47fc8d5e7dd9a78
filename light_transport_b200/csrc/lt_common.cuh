// lt_common.cuh — small vector math, Philox4x32-10, error plumbing shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/lt_b200.h"

#define LT_DEVINL __device__ __forceinline__

namespace lt {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
#define LT_CUDA_OK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            lt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            cudaGetLastError(); /* a recoverable error must not resurface in a later call's launch check */ \
            return LT_E_CUDA;                                                                     \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------- vectors
template <typename R>
struct V3 {
    R x, y, z;
};
template <typename R> LT_DEVINL V3<R> mk(R x, R y, R z) { return V3<R>{x, y, z}; }
template <typename R> LT_DEVINL V3<R> operator+(V3<R> a, V3<R> b) { return mk<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename R> LT_DEVINL V3<R> operator-(V3<R> a, V3<R> b) { return mk<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename R> LT_DEVINL V3<R> operator-(V3<R> a) { return mk<R>(-a.x, -a.y, -a.z); }
template <typename R> LT_DEVINL V3<R> operator*(V3<R> a, R s) { return mk<R>(a.x * s, a.y * s, a.z * s); }
template <typename R> LT_DEVINL V3<R> operator*(V3<R> a, V3<R> b) { return mk<R>(a.x * b.x, a.y * b.y, a.z * b.z); }

// The reference's np.dot runs through OpenBLAS ddot whose scalar tail is an FMA chain
// fma(a2,b2, fma(a1,b1, a0*b0)) (probed).  Writing the dot this way makes the
// f64 kernels resolve exact geometric ties (rays along box edges, Q13 diagonal pixels) like the
// reference, and is also the fastest form in f32.
LT_DEVINL float dot(V3<float> a, V3<float> b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x))); }
LT_DEVINL double dot(V3<double> a, V3<double> b) { return __fma_rn(a.z, b.z, __fma_rn(a.y, b.y, __dmul_rn(a.x, b.x))); }

// np.cross: plain products and differences, no contraction
LT_DEVINL V3<double> cross(V3<double> a, V3<double> b) {
    return mk<double>(__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
                      __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
LT_DEVINL V3<float> cross(V3<float> a, V3<float> b) {
    return mk<float>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// np.linalg.norm runs through OpenBLAS dnrm2 (x87 extended accumulate, one rounding; probed to equal
// the correctly rounded sqrt of the exact sum).  f64: sum of squares in double-double, then one
// Newton correction of the rounded sqrt.
LT_DEVINL double norm(V3<double> v) {
    double p0 = __dmul_rn(v.x, v.x), e0 = __fma_rn(v.x, v.x, -p0);
    double p1 = __dmul_rn(v.y, v.y), e1 = __fma_rn(v.y, v.y, -p1);
    double p2 = __dmul_rn(v.z, v.z), e2 = __fma_rn(v.z, v.z, -p2);
    // two-sum p0+p1
    double s = __dadd_rn(p0, p1), bb = __dsub_rn(s, p0);
    double err = __dadd_rn(__dsub_rn(p0, __dsub_rn(s, bb)), __dsub_rn(p1, bb));
    double s2 = __dadd_rn(s, p2), bb2 = __dsub_rn(s2, s);
    double err2 = __dadd_rn(__dsub_rn(s, __dsub_rn(s2, bb2)), __dsub_rn(p2, bb2));
    double lo = __dadd_rn(__dadd_rn(err, err2), __dadd_rn(__dadd_rn(e0, e1), e2));
    double r = __dsqrt_rn(s2);
    if (!(r > 0.0) || isinf(r)) return r;
    // residual (s2 + lo - r*r) / (2r), r*r computed exactly with fma
    double rr = __dmul_rn(r, r), re = __fma_rn(r, r, -rr);
    double res = __dadd_rn(__dsub_rn(__dsub_rn(s2, rr), re), lo);
    return __dadd_rn(r, __ddiv_rn(res, __dmul_rn(2.0, r)));
}
LT_DEVINL float norm(V3<float> v) { return sqrtf(dot(v, v)); }

template <typename R> LT_DEVINL V3<R> normalize(V3<R> v) {
    R n = norm(v);
    return mk<R>(v.x / n, v.y / n, v.z / n);
}
// fp32 fast path: one reciprocal-sqrt instead of a sqrt and three divides
LT_DEVINL V3<float> normalize_fast(V3<float> v) {
    float inv = rsqrtf(dot(v, v));
    return mk<float>(v.x * inv, v.y * inv, v.z * inv);
}
LT_DEVINL V3<double> normalize_fast(V3<double> v) { return normalize(v); }

template <typename R> LT_DEVINL R rabs(R x);
template <> LT_DEVINL float rabs<float>(float x) { return fabsf(x); }
template <> LT_DEVINL double rabs<double>(double x) { return fabs(x); }
template <typename R> LT_DEVINL R real_sqrt(R x);
template <> LT_DEVINL float real_sqrt<float>(float x) { return sqrtf(x); }
template <> LT_DEVINL double real_sqrt<double>(double x) { return __dsqrt_rn(x); }
template <typename R> LT_DEVINL R comp(V3<R> v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

template <typename R> struct Lim;
template <> struct Lim<float> { static LT_DEVINL float inf() { return __int_as_float(0x7f800000); } };
template <> struct Lim<double> { static LT_DEVINL double inf() { return __longlong_as_double(0x7ff0000000000000LL); } };

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11)
// key = 64-bit seed; counter = (path_lo, path_hi, bounce | nee index, stream).  Stream 0 feeds
// rand_0/rand_1 of a bounce (replaces Scene.rand_0/rand_1, R/src/scene.py:68-69), stream 1 the light
// pick of the k-th next-event estimation of a path (replaces np.random.choice, light_samples.py:38).
LT_DEVINL void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c1 = lo1; c3 = lo0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 53-bit uniform in [0,1) from two words (NumPy's random_double recipe); the fp32 variant keeps the
// top 24 bits of the same first word, so an f32 walk follows its f64 twin closely.
LT_DEVINL void u01(uint32_t a, uint32_t b, double& u) { u = ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0); }
LT_DEVINL void u01(uint32_t a, uint32_t, float& u) { u = (float)(a >> 8) * (1.0f / 16777216.0f); }

}  // namespace lt
