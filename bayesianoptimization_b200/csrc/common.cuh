// common.cuh - shared device helpers for the b200bo kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/b200bo.h"

namespace b200bo {

constexpr int kPad = 128;  // training-set size is padded to a multiple of this

// ---- branch-free fp64 primitives for the covariance functions -----------------------------
// The kernel-matrix builders evaluate sqrt and exp for every (training point, candidate) pair
// with only a few resident warps, so data-dependent slow-path branches (libm special cases) and
// their code size hurt more than the arithmetic.  Both routines are within 2 ulp of libm on their
// domain (checked against numpy over 3e5 random arguments), far inside the 1e-5 parity bar.

// Polynomial / reduction constants live in constant memory: an fp64 immediate costs two uniform
// moves every time it is used, a constant-bank operand is free.
__constant__ double kExpC[14] = {
    1.6059043836821613e-10, 2.0876756987868100e-09, 2.5052108385441720e-08, 2.7557319223985893e-07,
    2.7557319223985888e-06, 2.4801587301587302e-05, 1.9841269841269841e-04, 1.3888888888888889e-03,
    8.3333333333333332e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0};
__constant__ double kExpR[4] = {1.4426950408889634074, -6.93147180369123816490e-01,
                                -1.90821492927058770002e-10, 700.0};

// sqrt(x) for x >= 0 (arguments below 1e-30 are treated as 1e-30: |error| <= 1e-15 absolute).
__device__ __forceinline__ double sqrt_pos(double x) {
    const double xc = fmax(x, 1e-30);
    double y = (double)rsqrtf((float)xc);  // 2^-23 seed
    const double h = 0.5 * xc;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    double s = xc * y;
    s = fma(fma(-s, s, xc), 0.5 * y, s);
    return s;
}

// exp(-k) for k >= 0 (k > 700 is clamped: the result, < 1e-304, is irrelevant at fp64 scale).
__device__ __forceinline__ double exp_neg(double k) {
    k = fmin(k, kExpR[3]);
    const double n = rint(-k * kExpR[0]);
    double r = fma(n, kExpR[1], -k);
    r = fma(n, kExpR[2], r);
    // degree-13 Taylor polynomial (|r| <= ln2/2: truncation 4e-18) in Estrin form: dependent
    // depth 4 instead of 13, so the few resident warps keep the fp64 pipe fed.  a_k = kExpC[13-k].
    const double r2 = r * r;
    const double b0 = fma(kExpC[12], r, kExpC[13]), b1 = fma(kExpC[10], r, kExpC[11]);
    const double b2 = fma(kExpC[8], r, kExpC[9]), b3 = fma(kExpC[6], r, kExpC[7]);
    const double b4 = fma(kExpC[4], r, kExpC[5]), b5 = fma(kExpC[2], r, kExpC[3]);
    const double b6 = fma(kExpC[0], r, kExpC[1]);
    const double r4 = r2 * r2;
    const double c0 = fma(b1, r2, b0), c1 = fma(b3, r2, b2), c2 = fma(b5, r2, b4);
    const double r8 = r4 * r4;
    const double d0 = fma(c1, r4, c0), d1 = fma(b6, r4, c2);
    const double p = fma(d1, r8, d0);
    const long long e = ((long long)n + 1023ll) << 52;  // 2^n, n in [-1010, 0]
    return p * __longlong_as_double(e);
}

// ---- covariance functions -------------------------------------------------------------
// COV codes: 0 = Matern nu=0.5, 1 = nu=1.5, 2 = nu=2.5, 3 = RBF / Matern nu=inf.
// Follows SK/gaussian_process/kernels.py:1722-1731 (Matern) and :1549/:1553 (RBF) operation by
// operation: dists = sqrt(r2); nu=2.5: K = dists*sqrt(5); (1 + K + K^2/3) * exp(-K).
template <int COV>
__device__ __forceinline__ double cov_eval(double r2) {
    if (COV == 3) return exp_neg(0.5 * r2);
    const double dist = sqrt_pos(r2);
    if (COV == 2) {
        const double k = dist * 2.23606797749978969641;  // math.sqrt(5)
        return (1.0 + k + k * k / 3.0) * exp_neg(k);
    }
    if (COV == 1) {
        const double k = dist * 1.73205080756887729353;  // math.sqrt(3)
        return (1.0 + k) * exp_neg(k);
    }
    return exp_neg(dist);
}

__host__ __device__ __forceinline__ int cov_code(int family, int nu) {
    return (family == B200BO_KERNEL_RBF || nu == B200BO_NU_INF) ? 3 : nu;
}

__device__ __forceinline__ double cov_from_r2(double r2, int family, int nu) {
    switch (cov_code(family, nu)) {
        case 0: return cov_eval<0>(r2);
        case 1: return cov_eval<1>(r2);
        case 2: return cov_eval<2>(r2);
        default: return cov_eval<3>(r2);
    }
}

// scipy.special.ndtr (cephes ndtr.c), which scipy.stats.norm.cdf evaluates
// (SP/stats/_continuous_distns.py:370-371).
__device__ __forceinline__ double ndtr(double a) {
    if (isnan(a)) return a;
    const double x = a * 0.70710678118654752440;
    const double z = fabs(x);
    if (z < 0.70710678118654752440) return 0.5 + 0.5 * erf(x);
    double y = 0.5 * erfc(z);
    if (x > 0) y = 1.0 - y;
    return y;
}

// norm.pdf: exp(-x^2/2)/sqrt(2*pi) (SP/stats/_continuous_distns.py:362-363)
__device__ __forceinline__ double norm_pdf(double x) {
    return exp(-(x * x) / 2.0) / 2.50662827463100050242;
}

// frozen norm(loc, scale).cdf(b) as scipy evaluates it: NaN unless scale > 0
// (rv_continuous.cdf argcheck), else ndtr((b - loc)/scale).
__device__ __forceinline__ double norm_cdf_loc_scale(double b, double loc, double scale) {
    if (!(scale > 0.0) || isnan(loc)) return CUDART_NAN;
    return ndtr((b - loc) / scale);
}

// order-preserving map double -> uint64 (for (value,index) selection keys)
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    if (v == 0.0) v = 0.0;  // -0.0 == +0.0 for np.argmin / np.argsort
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
// np.argmin: NaN is the minimum (first NaN wins)
__device__ __forceinline__ unsigned long long key_nan_first(double v) {
    return isnan(v) ? 0ull : ordered_bits(v);
}
// np.argsort: NaN sorts last
__device__ __forceinline__ unsigned long long key_nan_last(double v) {
    return isnan(v) ? 0xFFFFFFFFFFFFFFFFull : ordered_bits(v);
}

// ---- cp.async helpers -------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace b200bo
