// fit_kernels.cuh - kernel-matrix build, blocked Cholesky, triangular inverse, LML gradient.
// Replaces (on device) the arithmetic of GaussianProcessRegressor.fit's tail and
// log_marginal_likelihood (SK/gaussian_process/_gpr.py:349-367, :584-651).
#pragma once
#include "common.cuh"
#include "potrf_block.cuh"

namespace b200bo {

// ---------------------------------------------------------------------------------------
// Xs = transform(X) / length_scale  (rows >= n are zero padding)
// sklearn divides before differencing: cdist(X / length_scale, Y / length_scale)
// (SK/gaussian_process/kernels.py:1716-1720).
// ---------------------------------------------------------------------------------------
__global__ void scale_x_kernel(const double* __restrict__ X, const double* __restrict__ ls,
                               const int* __restrict__ xform, double* __restrict__ Xs, int n,
                               int np, int d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)np * d) return;
    const int i = (int)(idx / d), j = (int)(idx % d);
    double v = 0.0;
    if (i < n) {
        v = X[idx];
        if (xform && xform[j] == B200BO_XFORM_ROUND) v = rint(v);  // np.round: half-to-even
        v = v / ls[j];
    }
    Xs[idx] = v;
}

// ---------------------------------------------------------------------------------------
// K(X,X): full symmetric np x np matrix; K_ii = const + alpha; padding = identity.
// (SK/gaussian_process/kernels.py:1716,1740-1743 + _gpr.py:350)
// ---------------------------------------------------------------------------------------
// 32x32 output tile per CTA (32x8 threads, 4 rows each); the two 32-row slabs of Xs are staged in
// shared memory with coalesced loads (row stride 65: conflict-free when lanes walk different rows).
__global__ void __launch_bounds__(256)
kbuild_kernel(const double* __restrict__ Xs, double* __restrict__ K, int n, int np, int d, int family,
              int nu, double constv, double jitter) {
    __shared__ double xi[32][B200BO_MAX_DIM + 1], xj[32][B200BO_MAX_DIM + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int idx = tid; idx < 32 * d; idx += 256) {
        const int r = idx / d, t = idx - r * d;
        xi[r][t] = Xs[(size_t)i0 * d + idx];
        xj[r][t] = Xs[(size_t)j0 * d + idx];
    }
    __syncthreads();
    const int j = j0 + tx;
    for (int rr = ty; rr < 32; rr += 8) {
        const int i = i0 + rr;
        double v;
        if (i >= n || j >= n) {
            v = (i == j) ? 1.0 : 0.0;
        } else if (i == j) {
            v = constv + jitter;
        } else {
            // (a-b)^2 is symmetric, so K is too (pdist evaluates each pair once)
            double r2 = 0.0;
            for (int t = 0; t < d; ++t) {
                const double df = xi[rr][t] - xj[tx][t];
                r2 += df * df;
            }
            v = constv * cov_from_r2(r2, family, nu);
        }
        K[(size_t)i * np + j] = v;
    }
}

// ---------------------------------------------------------------------------------------
// Generic fp64 GEMM on 64x64 tiles (fit-side building block; inner product on the fp64 tensor
// path, mma.sync m8n8k4 - the trailing SYRK/GEMM updates of the blocked Cholesky, the triangular
// inverse recursion and K^-1 all run through it):
//   C[m][n] = beta*C[m][n] + alpha * sum_k opA(m,k) * opB(k,n)
//   opA(m,k) = TA ? A[k*lda+m] : A[m*lda+k];  opB(k,n) = TB ? B[n*ldb+k] : B[k*ldb+n]
// M, N multiples of 64; K multiple of 16.  lower_only: skip tiles strictly above the diagonal.
// kmode: 0 = full K range; 1 = k in [0, m0+64)   (A or B lower-triangular in (m,k)/(k,n) sense:
//        caller guarantees contributions with k >= m0+64 vanish);
//        2 = k in [n0, K)                       (contributions with k < n0 vanish)
//        3 = k in [max(m0,n0), K)
// Batched over blockIdx.z with element strides sA/sB/sC.
// ---------------------------------------------------------------------------------------
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
dgemm64_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int lda,
               long long sA, const double* __restrict__ B, int ldb, long long sB, double beta,
               double* C, int ldc, long long sC, int lower_only, int kmode) {
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    if (lower_only && n0 > m0) return;
    A += (long long)blockIdx.z * sA;
    B += (long long)blockIdx.z * sB;
    C += (long long)blockIdx.z * sC;
    // k-major tiles, row stride 68 doubles = 8 words (mod 32): the 4 k-rows x 4 columns an LDS.64
    // half-warp touches for an m8n8k4 fragment fall on disjoint banks
    __shared__ double As[16][68];
    __shared__ double Bs[16][68];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;  // warp tile: rows wm*16..+16, cols wn*32..+32
    const int g = lane >> 2, t4 = lane & 3;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int kbeg = 0, kend = K;
    if (kmode == 1) kend = min(K, m0 + 64);
    if (kmode == 2) kbeg = n0;
    if (kmode == 3) kbeg = max(m0, n0);

    for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int idx = tid + t * 256;
            if (TA) {
                const int m = idx & 63, kk = idx >> 6;
                As[kk][m] = A[(size_t)(k0 + kk) * lda + m0 + m];
            } else {
                const int kk = idx & 15, m = idx >> 4;
                As[kk][m] = A[(size_t)(m0 + m) * lda + k0 + kk];
            }
            if (TB) {
                const int kk = idx & 15, n = idx >> 4;
                Bs[kk][n] = B[(size_t)(n0 + n) * ldb + k0 + kk];
            } else {
                const int n = idx & 63, kk = idx >> 6;
                Bs[kk][n] = B[(size_t)(k0 + kk) * ldb + n0 + n];
            }
        }
        __syncthreads();
        // fp64 tensor path: mma.sync m8n8k4 (DMMA); 2 x 4 fragments per warp and k-step
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = As[k4 * 4 + t4][wm * 16 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k4 * 4 + t4][wn * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                                 : "d"(a[i]), "d"(b[j]));
        }
        __syncthreads();
    }
    // C fragment: row = g, columns 2*t4 + {0,1}
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double* c = C + (size_t)(m0 + wm * 16 + i * 8 + g) * ldc + n0 + wn * 32 + j * 8 + 2 * t4 + e;
                const double v = alpha * acc[i][j][e];
                *c = (beta == 0.0) ? v : fma(beta, *c, v);
            }
}

// ---------------------------------------------------------------------------------------
// Diagonal 64x64 block: in-place Cholesky (lower) + inverse of the factor written to Dinv.
// Single CTA, 256 threads, phases in potrf_block.cuh (8-column panels: 3 barriers per panel
// instead of 2 per column, one rsqrt per pivot instead of sqrt + divisions; inverse by recursive
// doubling instead of 64 dependent substitution steps).  Agrees with the unblocked kernel below to
// round-off.
// info: 0 or the 1-based index of the first non-positive pivot (LAPACK dpotrf convention,
// SP/linalg/_decomp_cholesky.py:58 raises LinAlgError on it).
// ---------------------------------------------------------------------------------------
constexpr int kPotrfSmemBytes = 3 * potrf::kB * potrf::kLd * 8;
__global__ void __launch_bounds__(256)
potrf_diag_kernel(double* A, int ld, int j0, double* Dinv, int ldd, int* info) {
    using namespace potrf;
    extern __shared__ __align__(16) double potrf_smem[];
    double* S = potrf_smem;
    double* V = potrf_smem + kB * kLd;
    double* T = potrf_smem + 2 * kB * kLd;
    __shared__ double diag[kB], rdiag[kB];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < kB * kB; idx += kThreads) {
        const int r = idx >> 6, c = idx & 63;
        S[r * kLd + c] = A[(size_t)(j0 + r) * ld + j0 + c];
        V[r * kLd + c] = 0.0;
    }
    __syncthreads();
    for (int c0 = 0; c0 < kB; c0 += kPw) {
        if (warp == 0) {
            const int bad = diag_factor(S, diag, rdiag, c0);
            if (bad != 0 && tid == 0 && *info == 0) *info = j0 + bad;
        }
        __syncthreads();
        if (c0 + kPw < kB) {
            panel_solve(tid, S, rdiag, c0);
            __syncthreads();
            trailing_update(tid, S, c0);
            __syncthreads();
        }
    }
    // factor back to global memory (strict upper part of the block zeroed)
    for (int idx = tid; idx < kB * kB; idx += kThreads) {
        const int r = idx >> 6, c = idx & 63;
        A[(size_t)(j0 + r) * ld + j0 + c] = (c < r) ? S[r * kLd + c] : (c == r ? diag[r] : 0.0);
    }
    diag_inverse(S, rdiag, V, warp);  // 8 warps <-> 8 diagonal 8x8 blocks
    __syncthreads();
    for (int s = kPw, l2 = 3; s < kB; s *= 2, ++l2) {
        inverse_level_t(tid, S, V, T, s, l2);
        __syncthreads();
        inverse_level_w(tid, V, T, s, l2);
        __syncthreads();
    }
    for (int idx = tid; idx < kB * kB; idx += kThreads) {
        const int r = idx >> 6, c = idx & 63;
        Dinv[(size_t)r * ldd + c] = V[r * kLd + c];
    }
}

// ---------------------------------------------------------------------------------------
// Diagonal 64x64 block, first version (kept for A/B runs: B200BO_POTRF=legacy): unblocked
// column-by-column Cholesky (two barriers per column) + inverse by forward substitution.
// Single CTA, 256 threads.  info: 0 or the 1-based index of the first non-positive pivot
// (LAPACK dpotrf convention, SP/linalg/_decomp_cholesky.py:58 raises LinAlgError on it).
// ---------------------------------------------------------------------------------------
constexpr int kPotrfLegacySmemBytes = 2 * 64 * 65 * 8;
__global__ void __launch_bounds__(256)
potrf_diag_legacy_kernel(double* A, int ld, int j0, double* Dinv, int ldd, int* info) {
    extern __shared__ __align__(16) double potrf_smem[];
    double (*S)[65] = reinterpret_cast<double (*)[65]>(potrf_smem);
    double (*V)[65] = reinterpret_cast<double (*)[65]>(potrf_smem + 64 * 65);
    const int tid = threadIdx.x;
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        S[r][c] = A[(size_t)(j0 + r) * ld + j0 + c];
    }
    __syncthreads();
    // right-looking, two barriers per column: every thread derives the pivot itself (the square
    // root goes to diag[], S[k][k] keeps the pivot so late readers still see it), the first 64
    // threads scale the column, then all 256 apply the rank-1 update to the trailing lower triangle
    __shared__ double diag[64];
    for (int k = 0; k < 64; ++k) {
        double piv = S[k][k];
        const bool bad = !(piv > 0.0);
        if (bad) piv = 1.0;
        const double lkk = sqrt(piv);
        if (tid == k) {
            diag[k] = lkk;
            if (bad && *info == 0) *info = j0 + k + 1;
        } else if (tid > k && tid < 64) {
            S[tid][k] = S[tid][k] / lkk;
        }
        __syncthreads();
        // trailing block (k+1..63)^2, lower part: S[i][j] -= S[i][k]*S[j][k]; 16x16 thread grid,
        // no integer divisions on the critical path
        for (int i = k + 1 + (tid >> 4); i < 64; i += 16) {
            const double lik = S[i][k];
            for (int j = k + 1 + (tid & 15); j <= i; j += 16) S[i][j] = fma(-lik, S[j][k], S[i][j]);
        }
        __syncthreads();
    }
    // write factor back (zero the strict upper part of the block)
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        A[(size_t)(j0 + r) * ld + j0 + c] = (c < r) ? S[r][c] : (c == r ? diag[r] : 0.0);
    }
    // inverse of the lower-triangular block: column c by forward substitution, 4 threads per
    // column share each dot product (lanes 4c..4c+3, combined with two shuffles, fixed order)
    {
        const int c = tid >> 2, q = tid & 3;
        for (int i = 0; i < 64; ++i) {
            double part = 0.0;
            for (int k = c + q; k < i; k += 4) part = fma(S[i][k], V[k][c], part);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (q == 0) {
                double v = 0.0;
                if (i >= c) v = (((i == c) ? 1.0 : 0.0) - part) / diag[i];
                V[i][c] = v;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        Dinv[(size_t)r * ldd + c] = V[r][c];
    }
}

// zero the strict upper triangle (scipy.linalg.cholesky(lower=True) returns a clean factor)
__global__ void zero_upper_kernel(double* A, int np) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < np && j < np && j > i) A[(size_t)i * np + j] = 0.0;
}

// out[j][i] = in[i][j]  (np multiple of 32)
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int np) {
    __shared__ double t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        t[r][threadIdx.x] = in[(size_t)(by + r) * np + bx + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        out[(size_t)(bx + r) * np + by + threadIdx.x] = t[threadIdx.x][r];
}

// y[i] = sum_j A[i][j] * x[j], j in [jbeg(i), jend(i)); one warp per row, fixed-order reduction.
// tri: 0 full [0,ncols); 1 lower (j <= i); 2 upper (j >= i)
__global__ void gemv_rows_kernel(const double* __restrict__ A, int ld, const double* __restrict__ x,
                                 double* __restrict__ y, int nrows, int ncols, int tri) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    int jb = 0, je = ncols;
    if (tri == 1) je = min(ncols, row + 1);
    if (tri == 2) jb = row & ~31;
    double s = 0.0;
    for (int j = jb + lane; j < je; j += 32) {
        if (tri == 2 && j < row) continue;
        s = fma(A[(size_t)row * ld + j], x[j], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
}

__global__ void diag_kernel(const double* __restrict__ A, int ld, double* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = A[(size_t)i * ld + i];
}

// r = y - r   (elementwise, n entries)
__global__ void residual_kernel(const double* __restrict__ y, double* __restrict__ r, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = y[i] - r[i];
}
// a += b
__global__ void axpy1_kernel(double* __restrict__ a, const double* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}

// ---------------------------------------------------------------------------------------
// LML gradient:  grad_p = 0.5 * sum_ij (alpha_i alpha_j - Kinv_ij) * dK_ij/dtheta_p
// (SK/gaussian_process/_gpr.py:629-651, kernels.py:1745-1786 / :1561-1573).
// theta order: [log const (if has_const)], log length_scale (1 or d).
// Each CTA reduces a 16x16 patch of (i,j) pairs in a fixed order and writes its partial
// sums to part[block][p]; the host adds the partials in index order (deterministic).
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// predict(return_cov=True) helpers (SK/gaussian_process/_gpr.py:464-475)
// ---------------------------------------------------------------------------------------
// Xcs = transform(Xc)/length_scale, rows >= m zero
__global__ void scale_xc_kernel(const double* __restrict__ Xc, const double* __restrict__ ls,
                                const int* __restrict__ xform, double* __restrict__ Xcs, int m, int mp, int d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)mp * d) return;
    const int i = (int)(idx / d), j = (int)(idx % d);
    double v = 0.0;
    if (i < m) {
        v = Xc[idx];
        if (xform && xform[j] == B200BO_XFORM_ROUND) v = rint(v);
        v = v / ls[j];
    }
    Xcs[idx] = v;
}
// Kst[k][c] = const * cov(Xs[k], Xcs[c])  (np x mp, zero for padded rows/columns)
__global__ void kcross_kernel(const double* __restrict__ Xs, const double* __restrict__ Xcs,
                              double* __restrict__ Kst, int n, int np, int m, int mp, int d, int family, int nu,
                              double constv) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= np || c >= mp) return;
    double v = 0.0;
    if (k < n && c < m) {
        const double* a = Xcs + (size_t)c * d;
        const double* b = Xs + (size_t)k * d;
        double r2 = 0.0;
        for (int t = 0; t < d; ++t) {
            const double df = a[t] - b[t];
            r2 = fma(df, df, r2);
        }
        v = constv * cov_from_r2(r2, family, nu);
    }
    Kst[(size_t)k * mp + c] = v;
}
// mu[c] = y_std * sum_k alpha[k] Kst[k][c] + y_mean
__global__ void cross_mean_kernel(const double* __restrict__ Kst, const double* __restrict__ alphav,
                                  double* __restrict__ mu, int np, int m, int mp, double y_mean, double y_std) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    double s = 0.0;
    for (int k = 0; k < np; ++k) s = fma(alphav[k], Kst[(size_t)k * mp + c], s);
    mu[c] = y_std * s + y_mean;
}
// cov[i][j] = (k(x_i, x_j) - (V^T V)[i][j]) * y_std^2   (m x m, contiguous)
__global__ void cov_finish_kernel(const double* __restrict__ Xcs, const double* __restrict__ VtV, int ldv,
                                  double* __restrict__ cov, int m, int d, int family, int nu, double constv,
                                  double y_std) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= m || j >= m) return;
    double kv;
    if (i == j) {
        kv = constv;
    } else {
        const double* a = Xcs + (size_t)min(i, j) * d;
        const double* b = Xcs + (size_t)max(i, j) * d;
        double r2 = 0.0;
        for (int t = 0; t < d; ++t) {
            const double df = a[t] - b[t];
            r2 += df * df;
        }
        kv = constv * cov_from_r2(r2, family, nu);
    }
    cov[(size_t)i * m + j] = (kv - VtV[(size_t)i * ldv + j]) * (y_std * y_std);
}

// ---------------------------------------------------------------------------------------
// Incremental factor update at fixed hyper-parameters: one training point
// is appended in O(N^2): new row of K, of L (l = L^-1 k, pivot sqrt(c+alpha-|l|^2)) and of L^-1.
// ---------------------------------------------------------------------------------------
// Xs[n] = transform(x)/ls ; kvec[i] = const*cov(Xs[i], Xs[n]) for i < n
__global__ void append_krow_kernel(const double* __restrict__ x_new, const double* __restrict__ ls,
                                   const int* __restrict__ xform, double* __restrict__ Xs, double* __restrict__ X,
                                   double* __restrict__ kvec, int n, int d, int family, int nu, double constv) {
    __shared__ double xs[B200BO_MAX_DIM];
    if (threadIdx.x < d) {
        double v = x_new[threadIdx.x];
        if (blockIdx.x == 0) X[(size_t)n * d + threadIdx.x] = v;
        if (xform && xform[threadIdx.x] == B200BO_XFORM_ROUND) v = rint(v);
        v = v / ls[threadIdx.x];
        xs[threadIdx.x] = v;
        if (blockIdx.x == 0) Xs[(size_t)n * d + threadIdx.x] = v;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* a = Xs + (size_t)i * d;
    double r2 = 0.0;
    for (int t = 0; t < d; ++t) {
        const double df = a[t] - xs[t];
        r2 += df * df;
    }
    kvec[i] = constv * cov_from_r2(r2, family, nu);
}
// write row/column n of K, row n of L; pivot check.  lvec = L^-1[0:n,0:n] kvec.
__global__ void append_rows_kernel(double* __restrict__ K, double* __restrict__ L, const double* __restrict__ kvec,
                                   const double* __restrict__ lvec, int n, int np, double diag, int* info,
                                   double* __restrict__ pivot_out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double l = lvec[i], k = kvec[i];
        s = fma(l, l, s);
        K[(size_t)n * np + i] = k;
        K[(size_t)i * np + n] = k;
        L[(size_t)n * np + i] = l;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int t = 128; t > 0; t >>= 1) {
        if (threadIdx.x < t) red[threadIdx.x] += red[threadIdx.x + t];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double piv = diag - red[0];
        if (!(piv > 0.0)) {
            *info = n + 1;
            piv = 1.0;
        }
        const double lnn = sqrt(piv);
        K[(size_t)n * np + n] = diag;
        L[(size_t)n * np + n] = lnn;
        *pivot_out = lnn;
    }
}
// row n of W = L^-1 and column n of WT:  W[n][j] = -(1/l_nn) * t[j], t = W[0:n,0:n]^T lvec ; W[n][n] = 1/l_nn
__global__ void append_winv_kernel(double* __restrict__ W, double* __restrict__ WT, const double* __restrict__ tvec,
                                   const double* __restrict__ pivot, int n, int np) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const double inv = 1.0 / *pivot;
    if (j < n) {
        const double w = -inv * tvec[j];
        W[(size_t)n * np + j] = w;
        WT[(size_t)j * np + n] = w;
    } else if (j == n) {
        W[(size_t)n * np + n] = inv;
        WT[(size_t)n * np + n] = inv;
    }
}

constexpr int kMaxTheta = B200BO_MAX_DIM + 1;

__global__ void __launch_bounds__(256)
lml_grad_kernel(const double* __restrict__ Xs, const double* __restrict__ Kinv, int ldk,
                const double* __restrict__ alphav, int n, int d, int family, int nu, double constv,
                int has_const, int aniso, double* __restrict__ part, int ntheta) {
    const int j = blockIdx.x * 16 + (threadIdx.x & 15);
    const int i = blockIdx.y * 16 + (threadIdx.x >> 4);
    __shared__ double red[256];
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    // K^-1 and dK/dtheta are symmetric: only pairs i >= j are visited (K^-1's upper tiles are never
    // computed) and off-diagonal pairs count twice
    const bool valid = (i < n && j < n && j <= i);
    double w = 0.0, r2 = 0.0, kval = 0.0, gcommon = 0.0;
    const double* a = Xs + (size_t)(valid ? i : 0) * d;
    const double* b = Xs + (size_t)(valid ? j : 0) * d;
    if (valid) {
        w = alphav[i] * alphav[j] - Kinv[(size_t)i * ldk + j];
        if (i != j) w *= 2.0;
        for (int t = 0; t < d; ++t) {
            const double df = a[t] - b[t];
            r2 += df * df;
        }
        kval = (i == j) ? 1.0 : cov_from_r2(r2, family, nu);
        // gradient factor such that dk/dlog(l_t) = gcommon * D_t  (D_t = scaled squared diff)
        if (family == B200BO_KERNEL_RBF || nu == B200BO_NU_INF) {
            gcommon = kval;  // K_gradient = D * K
        } else if (nu == B200BO_NU_25) {
            const double tmp = sqrt(5.0 * r2);
            gcommon = 5.0 / 3.0 * (tmp + 1.0) * exp(-tmp);
        } else if (nu == B200BO_NU_15) {
            gcommon = 3.0 * exp(-sqrt(3.0 * r2));
        } else {  // nu = 0.5: K * D / sqrt(sum D), 0 where the distance is 0
            const double den = sqrt(r2);
            gcommon = (den != 0.0) ? kval / den : 0.0;
        }
    }
    for (int p = 0; p < ntheta; ++p) {
        double g = 0.0;
        if (valid) {
            if (has_const && p == 0) {
                g = constv * kval;
            } else if (!aniso) {
                g = constv * gcommon * r2;
            } else {
                const int t = p - has_const;
                const double df = a[t] - b[t];
                g = constv * gcommon * (df * df);
            }
        }
        red[threadIdx.x] = w * g;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) part[(size_t)bid * ntheta + p] = 0.5 * red[0];
        __syncthreads();
    }
}

}  // namespace b200bo
