// potrf_block.cuh - 64x64 diagonal block of the blocked Cholesky: factor + explicit inverse,
// written as barrier-separated PHASES of a 256-thread CTA.  Every phase is a plain function of
// the thread id over three 64x65 "shared" arrays, so the same source is (a) the body of
// potrf_diag_kernel and (b) compiled by g++ into a sequential emulation (tools/potrf_emul.cpp)
// that checks the index logic on a machine without a GPU.
//
// Algorithm: right-looking with 8-column panels.  Per panel: the 8x8 diagonal block is factored in
// one warp's REGISTERS (every lane redundantly: no shuffles, full ILP), the rows below are solved
// by substitution (one thread per row), the trailing lower triangle gets the rank-8 update.  The
// floating-point operation ORDER per entry is exactly that of the unblocked column-by-column
// algorithm (subtract l_ik*l_jk for k ascending, then scale by the pivot's reciprocal square root).
// Square roots and divisions sit on the kernel's critical path (64 dependent pivots), so each
// pivot costs ONE rsqrt: l_kk = p * rsqrt(p), column entries are multiplied by rsqrt(p) - a few
// ulp from the sqrt/divide formulation, far inside the 1e-5 parity tolerance.  The inverse of the factor is built from the eight 8x8 diagonal
// inverses by recursive doubling: inv([[A,0],[C,B]]) = [[A^-1,0],[-B^-1 C A^-1, B^-1]].
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#define PB_UNROLL _Pragma("unroll")
#if defined(__CUDA_ARCH__)
#define PB_SYNCWARP() __syncwarp()
#define PB_RSQRT(x) rsqrt(x)
#else
#define PB_SYNCWARP()
#define PB_RSQRT(x) (1.0 / sqrt(x))
#endif
#else
#define PB_HD inline
#define PB_UNROLL
#define PB_SYNCWARP()
#define PB_RSQRT(x) (1.0 / sqrt(x))
#endif

namespace b200bo {
namespace potrf {

constexpr int kB = 64;    // block edge
constexpr int kLd = 65;   // row stride of the shared arrays (bank-conflict-free column walks)
constexpr int kPw = 8;    // panel width
constexpr int kThreads = 256;

// ---- phase D (one warp, all lanes identical): factor the 8x8 diagonal block at c0 ------------
// Reads S[c0..c0+8)^2 (lower part), writes the strict lower part of the factor back to S, the
// diagonal of the factor to diag[c0..c0+8) and its reciprocal to rdiag[c0..c0+8).  Returns 0 or the 1-based local index (c0+k+1) of the
// first non-positive pivot; a bad pivot is replaced by 1 so the arithmetic stays finite.
PB_HD int diag_factor(double* S, double* diag, double* rdiag, int c0) {
    double rk[kPw];
    double a[kPw][kPw];
    PB_UNROLL
    for (int r = 0; r < kPw; ++r) {
        PB_UNROLL
        for (int c = 0; c < kPw; ++c) a[r][c] = (c <= r) ? S[(c0 + r) * kLd + c0 + c] : 0.0;
    }
    int bad_at = 0;
    PB_UNROLL
    for (int k = 0; k < kPw; ++k) {
        double piv = a[k][k];
        const bool bad = !(piv > 0.0);
        if (bad) {
            piv = 1.0;
            if (bad_at == 0) bad_at = c0 + k + 1;
        }
        rk[k] = PB_RSQRT(piv);
        a[k][k] = piv * rk[k];
        PB_UNROLL
        for (int r = k + 1; r < kPw; ++r) a[r][k] = a[r][k] * rk[k];
        PB_UNROLL
        for (int r = k + 1; r < kPw; ++r) {
            PB_UNROLL
            for (int c = k + 1; c <= r; ++c) a[r][c] = fma(-a[r][k], a[c][k], a[r][c]);
        }
    }
    PB_SYNCWARP();  // every lane has read the block before any lane overwrites it
    PB_UNROLL
    for (int r = 0; r < kPw; ++r) {
        diag[c0 + r] = a[r][r];
        rdiag[c0 + r] = rk[r];
        PB_UNROLL
        for (int c = 0; c < r; ++c) S[(c0 + r) * kLd + c0 + c] = a[r][c];
    }
    return bad_at;
}

// ---- phase P: rows below the diagonal block, one thread per row -------------------------------
// l_ik = (a_ik - sum_{m<k} l_im l_km) * (1/l_kk)  for the panel's 8 columns, k ascending.
PB_HD void panel_solve(int tid, double* S, const double* rdiag, int c0) {
    const int i = c0 + kPw + tid;
    if (i >= kB) return;
    double l[kPw];
    PB_UNROLL
    for (int k = 0; k < kPw; ++k) {
        double v = S[i * kLd + c0 + k];
        PB_UNROLL
        for (int m = 0; m < k; ++m) v = fma(-l[m], S[(c0 + k) * kLd + c0 + m], v);
        l[k] = v * rdiag[c0 + k];
    }
    PB_UNROLL
    for (int k = 0; k < kPw; ++k) S[i * kLd + c0 + k] = l[k];
}

// ---- phase U: rank-8 update of the trailing lower triangle, 16x16 thread grid -----------------
PB_HD void trailing_update(int tid, double* S, int c0) {
    const int c1 = c0 + kPw;
    for (int i = c1 + (tid >> 4); i < kB; i += 16) {
        double li[kPw];
        PB_UNROLL
        for (int k = 0; k < kPw; ++k) li[k] = S[i * kLd + c0 + k];
        for (int j = c1 + (tid & 15); j <= i; j += 16) {
            double s = S[i * kLd + j];
            PB_UNROLL
            for (int k = 0; k < kPw; ++k) s = fma(-li[k], S[j * kLd + c0 + k], s);
            S[i * kLd + j] = s;
        }
    }
}

// ---- phase I0: inverse of the 8x8 diagonal block b (one warp per block, lanes identical) ------
// V (zero-initialised) receives the lower-triangular inverse; S holds the strict lower factor and
// rdiag[] the reciprocal of its diagonal.
PB_HD void diag_inverse(const double* S, const double* rdiag, double* V, int b) {
    const int c0 = b * kPw;
    double l[kPw][kPw], w[kPw][kPw];
    PB_UNROLL
    for (int r = 0; r < kPw; ++r) {
        PB_UNROLL
        for (int c = 0; c < kPw; ++c) {
            l[r][c] = (c < r) ? S[(c0 + r) * kLd + c0 + c] : 0.0;
            w[r][c] = 0.0;
        }
    }
    PB_UNROLL
    for (int c = 0; c < kPw; ++c) {
        w[c][c] = rdiag[c0 + c];
        PB_UNROLL
        for (int r = c + 1; r < kPw; ++r) {
            double s = 0.0;
            PB_UNROLL
            for (int m = c; m < r; ++m) s = fma(l[r][m], w[m][c], s);
            w[r][c] = -s * rdiag[c0 + r];
        }
    }
    PB_UNROLL
    for (int r = 0; r < kPw; ++r) {
        PB_UNROLL
        for (int c = 0; c <= r; ++c) V[(c0 + r) * kLd + c0 + c] = w[r][c];
    }
}

// ---- phases I1/I2 at doubling level s (8, 16, 32) ----------------------------------------------
// pairs of diagonal blocks of size s at offset o = 2s*p:  A^-1 = V[o.., o..], B^-1 = V[o+s.., o+s..],
// C = S[o+s.., o..].   I1:  T[o+s+i][o+j] = sum_k C[i][k] A^-1[k][j]   (k >= j: A^-1 is lower)
//                      I2:  V[o+s+i][o+j] = -sum_k B^-1[i][k] T[o+s+k][o+j]   (k <= i)
// 32*s entries per phase, (32*s)/256 per thread, fixed summation order.
PB_HD void inverse_level_t(int tid, const double* S, const double* V, double* T, int s, int log2s) {
    const int entries = 32 * s;
    for (int idx = tid; idx < entries; idx += kThreads) {
        const int p = idx >> (2 * log2s), rem = idx & (s * s - 1);
        const int i = rem >> log2s, j = rem & (s - 1);
        const int o = 2 * s * p;
        const double* Crow = S + (o + s + i) * kLd + o;
        double acc = 0.0;
        for (int k = j; k < s; ++k) acc = fma(Crow[k], V[(o + k) * kLd + o + j], acc);
        T[(o + s + i) * kLd + o + j] = acc;
    }
}
PB_HD void inverse_level_w(int tid, double* V, const double* T, int s, int log2s) {
    const int entries = 32 * s;
    for (int idx = tid; idx < entries; idx += kThreads) {
        const int p = idx >> (2 * log2s), rem = idx & (s * s - 1);
        const int i = rem >> log2s, j = rem & (s - 1);
        const int o = 2 * s * p;
        const double* Brow = V + (o + s + i) * kLd + o + s;
        double acc = 0.0;
        for (int k = 0; k <= i; ++k) acc = fma(Brow[k], T[(o + s + k) * kLd + o + j], acc);
        V[(o + s + i) * kLd + o + j] = -acc;
    }
}

}  // namespace potrf
}  // namespace b200bo
