// common.cuh - shared device helpers for the b200bo kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/b200bo.h"

namespace b200bo {

constexpr int kPad = 128;  // training-set size is padded to a multiple of this

// ---- covariance functions -------------------------------------------------------------
// Follows SK/gaussian_process/kernels.py:1722-1731 (Matern) and :1549/:1553 (RBF) operation by
// operation: dists = sqrt(r2); nu=2.5: K = dists*sqrt(5); (1 + K + K^2/3) * exp(-K).
__device__ __forceinline__ double cov_from_r2(double r2, int family, int nu) {
    if (family == B200BO_KERNEL_RBF) return exp(-0.5 * r2);
    const double dist = sqrt(r2);
    if (nu == B200BO_NU_25) {
        const double k = dist * 2.23606797749978969641;  // math.sqrt(5)
        return (1.0 + k + k * k / 3.0) * exp(-k);
    }
    if (nu == B200BO_NU_15) {
        const double k = dist * 1.73205080756887729353;  // math.sqrt(3)
        return (1.0 + k) * exp(-k);
    }
    if (nu == B200BO_NU_05) return exp(-dist);
    return exp(-(dist * dist) / 2.0);  // nu = inf
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
