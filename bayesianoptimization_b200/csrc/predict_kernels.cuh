// predict_kernels.cuh - the fused posterior-predict + acquisition kernel (fp64) and the
// (value,index) selection kernel.
//
// Replaces, for a batch of M candidates, the whole chain
//   K* = kernel_(X*, X)            SK/gaussian_process/_gpr.py:446, kernels.py:1720-1729
//   mu = s_y * K* alpha_ + y_mean  :447-450
//   V  = L^-1 K*^T                 :460-462   (the N^2 M term)
//   var = diag - sum_i V_i^2, clamp, sd = sqrt(var * s_y^2)   :480-500
//   -base_acq(mu, sd) [* prod_j p_j]   R/bayes_opt/acquisition.py:199-217, :485/:660/:847,
//                                      R/bayes_opt/constraint.py:200-221
// in ONE launch.  V is formed as the triangular GEMM  V = Linv * K*^T  against the cached
// explicit inverse of the Cholesky factor (computed once at fit time), so the per-candidate
// work has no dependency chain.
//
// Decomposition: a persistent grid (one CTA per SM); each CTA owns tiles of BN = 128
// candidates.  Per tile and per GP:
//   phase A  build K*^T (np x 128) once into a CTA-private HBM/L2 scratch + accumulate K* alpha_
//   phase B  for each 128-row block of Linv: acc(128x128) = sum_{k<=rows} LinvT[k][rows]^T K*[k][:],
//            8x8 register tiles, 3-stage cp.async pipeline; then colsq += sum_rows acc^2
//   phase C  mu, sd, acquisition / constraint probability per candidate
// All reductions are fixed-order (no floating-point atomics): results are bit-reproducible and
// independent of the grid size.
#pragma once
#include "common.cuh"

namespace b200bo {

struct GpDev {
    const double* Xs;      // [np][d]   transform(X)/length_scale, zero padded
    const double* linvT;   // [np][np]  (L^-1)^T row-major: linvT[k][i] = Linv[i][k]
    const double* alphav;  // [np]      alpha_, zero padded
    const double* ls;      // [d]       length scales (replicated when isotropic)
    const int* xform;      // [d] or nullptr
    int n, np, family, nu;
    double constv, y_mean, y_std, lb, ub;
};

struct PredictParams {
    GpDev gp[B200BO_MAX_GPS];
    int n_gps, d, acq_kind, pad0;
    double kappa, xi, y_max;
    const double* Xc;  // [m][d]
    long long m;
    double* acq_out;   // [m] or nullptr
    double* mu_out;    // [m] or nullptr (target GP)
    double* sd_out;    // [m] or nullptr (target GP)
    double* scratch;   // gridDim.x * scratch_stride doubles
    long long scratch_stride;
    unsigned long long* clamp_count;  // nullable
};

constexpr int PBM = 128, PBN = 128, PBK = 16, PSTAGES = 3, PNT = 256;
constexpr int kPredictSmemBytes = PSTAGES * PBK * (PBM + PBN) * 8;  // 98304

__global__ void __launch_bounds__(PNT, 1) predict_acq_kernel(const PredictParams P) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double mu_s[2][PBN];
    __shared__ double base_s[PBN];
    __shared__ double prod_s[PBN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int d = P.d;
    double* Ks = P.scratch + (long long)blockIdx.x * P.scratch_stride;
    const long long ntiles = (P.m + PBN - 1) / PBN;

    double* As = smem;                           // [PSTAGES][PBK][PBM]
    double* Bs = smem + PSTAGES * PBK * PBM;     // [PSTAGES][PBK][PBN]
    double* xc_s = smem;                         // phase A: [d][PBN]
    double* red = smem;                          // reduction: [16][PBN]

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long c0 = tile * PBN;
        for (int g = 0; g < P.n_gps; ++g) {
            const GpDev& G = P.gp[g];
            const int np = G.np;
            // ---------------- phase A: K*^T tile + K* alpha_ -------------------------------
            for (int idx = tid; idx < PBN * d; idx += PNT) {
                const int c = idx / d, j = idx - c * d;
                const long long gi = c0 + c;
                double v = 0.0;
                if (gi < P.m) {
                    v = P.Xc[gi * d + j];
                    if (G.xform && G.xform[j] == B200BO_XFORM_ROUND) v = rint(v);
                    v = v / G.ls[j];
                }
                xc_s[j * PBN + c] = v;
            }
            __syncthreads();
            {
                const int c = tid & (PBN - 1), half = tid >> 7;
                double mu_acc = 0.0;
                for (int n0 = half * 4; n0 < np; n0 += 8) {
                    double r2[4] = {0.0, 0.0, 0.0, 0.0};
                    const double* x0 = G.Xs + (size_t)n0 * d;
                    for (int j = 0; j < d; ++j) {
                        const double xv = xc_s[j * PBN + c];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double df = xv - __ldg(x0 + q * d + j);
                            r2[q] = fma(df, df, r2[q]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int n = n0 + q;
                        double kv = 0.0;
                        if (n < G.n) kv = G.constv * cov_from_r2(r2[q], G.family, G.nu);
                        Ks[(size_t)n * PBN + c] = kv;
                        mu_acc = fma(__ldg(G.alphav + n), kv, mu_acc);
                    }
                }
                mu_s[half][c] = mu_acc;
            }
            __threadfence_block();
            __syncthreads();

            // ---------------- phase B: V = Linv K*^T, column sums of V^2 --------------------
            double csq[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) csq[j] = 0.0;
            const int nb = np / PBM;
            for (int ib = 0; ib < nb; ++ib) {
                double acc[8][8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
                const int nks = (ib + 1) * (PBM / PBK);
                const double* Abase = G.linvT + (size_t)ib * PBM;

                auto load_stage = [&](int stage, int ks) {
                    const double* Ag = Abase + (size_t)(ks * PBK) * np;
                    const double* Bg = Ks + (size_t)(ks * PBK) * PBN;
                    double* as = As + stage * PBK * PBM;
                    double* bs = Bs + stage * PBK * PBN;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int q = tid + t * PNT;
                        const int kk = q >> 6, m2 = (q & 63) * 2;
                        cp_async16_cg(as + kk * PBM + m2, Ag + (size_t)kk * np + m2);
                        cp_async16_cg(bs + kk * PBN + m2, Bg + kk * PBN + m2);
                    }
                };

#pragma unroll
                for (int s = 0; s < PSTAGES - 1; ++s) {
                    if (s < nks) load_stage(s, s);
                    cp_async_commit();
                }
                for (int ks = 0; ks < nks; ++ks) {
                    cp_async_wait<PSTAGES - 2>();
                    __syncthreads();
                    const int nxt = ks + PSTAGES - 1;
                    if (nxt < nks) load_stage(nxt % PSTAGES, nxt);
                    cp_async_commit();
                    const double* as = As + (ks % PSTAGES) * PBK * PBM;
                    const double* bs = Bs + (ks % PSTAGES) * PBK * PBN;
#pragma unroll
                    for (int kk = 0; kk < PBK; ++kk) {
                        double a[8], b[8];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const double2 t =
                                *reinterpret_cast<const double2*>(as + kk * PBM + p * 32 + ty * 2);
                            a[2 * p] = t.x;
                            a[2 * p + 1] = t.y;
                        }
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const double2 t =
                                *reinterpret_cast<const double2*>(bs + kk * PBN + p * 32 + tx * 2);
                            b[2 * p] = t.x;
                            b[2 * p + 1] = t.y;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
                    }
                }
                cp_async_wait<0>();
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) s = fma(acc[i][j], acc[i][j], s);
                    csq[j] += s;
                }
            }
            // reduce csq over the 16 row-threads sharing each column (fixed order)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                red[ty * PBN + p * 32 + tx * 2] = csq[2 * p];
                red[ty * PBN + p * 32 + tx * 2 + 1] = csq[2 * p + 1];
            }
            __syncthreads();

            // ---------------- phase C: per-candidate epilogue --------------------------------
            if (tid < PBN) {
                const int c = tid;
                const long long gi = c0 + c;
                double colsq = 0.0;
#pragma unroll
                for (int r = 0; r < 16; ++r) colsq += red[r * PBN + c];
                const double mu_n = mu_s[0][c] + mu_s[1][c];
                const double mean = G.y_std * mu_n + G.y_mean;
                double var = G.constv - colsq;
                if (var < 0.0) {
                    var = 0.0;
                    if (P.clamp_count && gi < P.m) atomicAdd(P.clamp_count, 1ull);
                }
                const double sd = sqrt(var * (G.y_std * G.y_std));
                if (g == 0) {
                    double base = 0.0;
                    if (P.acq_kind == B200BO_ACQ_UCB) {
                        base = mean + P.kappa * sd;
                    } else if (P.acq_kind == B200BO_ACQ_EI) {
                        const double a = mean - P.y_max - P.xi;
                        const double z = a / sd;
                        base = a * ndtr(z) + sd * norm_pdf(z);
                    } else if (P.acq_kind == B200BO_ACQ_POI) {
                        const double z = (mean - P.y_max - P.xi) / sd;
                        base = ndtr(z);
                    }
                    base_s[c] = -1.0 * base;
                    prod_s[c] = 1.0;
                    if (gi < P.m) {
                        if (P.mu_out) P.mu_out[gi] = mean;
                        if (P.sd_out) P.sd_out[gi] = sd;
                    }
                } else {
                    const double p_lo =
                        (G.lb == -CUDART_INF) ? 0.0 : norm_cdf_loc_scale(G.lb, mean, sd);
                    const double p_hi =
                        (G.ub == CUDART_INF) ? 1.0 : norm_cdf_loc_scale(G.ub, mean, sd);
                    // constraint.py:208 (J=1: result = p_hi - p_lo) / :219 (result *= ...)
                    prod_s[c] = (g == 1) ? (p_hi - p_lo) : prod_s[c] * (p_hi - p_lo);
                }
                if (g == P.n_gps - 1 && P.acq_out && gi < P.m) {
                    P.acq_out[gi] = (P.n_gps > 1) ? base_s[c] * prod_s[c] : base_s[c];
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------
// Selection: record 0 = np.argmin(vals) (first NaN wins, ties -> lowest index);
// records 1..k = the k smallest by (value, index) with NaN last (np.argsort order for
// distinct values).  Single CTA of 1024 threads; k+1 fixed-order passes over vals (L2).
// (R/bayes_opt/acquisition.py:313-317)
// ---------------------------------------------------------------------------------------
struct SelRecord {
    double value;
    long long index;
};

__global__ void __launch_bounds__(1024)
select_kernel(const double* __restrict__ vals, long long m, int k, SelRecord* __restrict__ out,
              long long index_base) {
    __shared__ unsigned long long skey[1024];
    __shared__ long long sidx[1024];
    __shared__ unsigned long long prev_key;
    __shared__ long long prev_idx;
    const int tid = threadIdx.x;
    for (int round = 0; round <= k; ++round) {
        const bool argmin_round = (round == 0);
        const bool bounded = (round >= 2);
        const unsigned long long pk = bounded ? prev_key : 0ull;
        const long long pi = bounded ? prev_idx : -1;
        unsigned long long bk = 0xFFFFFFFFFFFFFFFFull;
        long long bi = -1;
        for (long long i = tid; i < m; i += 1024) {
            const double v = vals[i];
            const unsigned long long key = argmin_round ? key_nan_first(v) : key_nan_last(v);
            if (bounded && (key < pk || (key == pk && i <= pi))) continue;
            if (bi < 0 || key < bk) {  // i increases, so ties keep the lowest index
                bk = key;
                bi = i;
            }
        }
        skey[tid] = bk;
        sidx[tid] = bi;
        __syncthreads();
        for (int s = 512; s > 0; s >>= 1) {
            if (tid < s) {
                const unsigned long long ok = skey[tid + s];
                const long long oi = sidx[tid + s];
                const long long mi = sidx[tid];
                const bool take = (oi >= 0) && (mi < 0 || ok < skey[tid] || (ok == skey[tid] && oi < mi));
                if (take) {
                    skey[tid] = ok;
                    sidx[tid] = oi;
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            const long long w = sidx[0];
            out[round].index = (w >= 0) ? w + index_base : -1;
            out[round].value = (w >= 0) ? vals[w] : CUDART_NAN;
            prev_key = skey[0];
            prev_idx = (w >= 0) ? w : (long long)0x7FFFFFFFFFFFFFFFll;
        }
        __syncthreads();
    }
}

}  // namespace b200bo
