// predict_tc3.cuh - fp32 mode with N = 256 per tcgen05.mma (candidate tiles of 256).
//
// predict_acq_tc2_kernel is shared-memory-bandwidth bound: an SS-mode 128x128x8 tf32 MMA reads 8 KiB of operands
// (A 4 + B 4) for 131 k multiply-adds, and the row-block PAIRING of that kernel issues two such MMAs per B tile.
// With the 256 candidates of a tile as ONE N = 256 operand, an MMA reads 12 KiB (A 4 + B 8) for 262 k multiply-adds:
// the same products for 25 % fewer operand bytes (36 instead of 48 KiB per k-step of 8).  Everything else is the
// structure of tc2 (see predict_kernels.cuh): 512 threads = producer (1-D bulk copies) / tcgen05.mma issuer /
// 4 epilogue warps (tcgen05.ld, v^2 transpose-reduce, acquisition epilogue, fused selection) / 8 builder warps that
// write the K* operand images of job j+1 while job j is on the tensor cores; 3xTF32 (a_hi*b_hi + a_hi*b_lo +
// a_lo*b_hi), fp32 accumulators in TMEM: one 128 x 256 accumulator per row block, double-buffered (512 columns).
// A stage = 16 k of [A hi | A lo | B hi | B lo] = 8 + 8 + 16 + 16 KiB; four stages.
// Selected with B200BO_TC_VARIANT=3 (d <= 16); measured against tc2 in DESIGN.md section 6.
#pragma once
#include "predict_kernels.cuh"

namespace b200bo {

constexpr int T3N = 256;                                   // candidates per tile
constexpr int T3_STAGES = 4;
constexpr int T3_AHALF = tc::kTcImgBytes / 2;              // 8192: 16 k of a 128-row image
constexpr int T3_BIMG = T3N * tc::kTcK * 4;                // 32768: 256-row x 32-k fp32 image
constexpr int T3_BHALF = T3_BIMG / 2;                      // 16384
constexpr int T3_STAGE_BYTES = 2 * T3_AHALF + 2 * T3_BHALF;  // 49152
constexpr int T3_BLBO = (T3N / 8) * 128;                   // 4096: K-adjacent core matrices of the B image
constexpr int T3_TMEM_COLS = 512;                          // 2 accumulators x 256 columns
constexpr int kPredictSmemBytesTc3 = T3_STAGES * T3_STAGE_BYTES + 2 * PA_CHUNK * kPredictMaxDimRegs * 8;  // 212992

__host__ __device__ constexpr int t3_bimg_offset(int r, int k) {
    return ((k >> 2) * (T3N / 8) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 3) * 4;
}

// builder: one thread per candidate column, all rows of every staged chunk
template <int COV>
__device__ __forceinline__ void tc3_build_job(const PredictParams& P, const GpDev& G, long long c0, int btid,
                                              uint8_t* __restrict__ Bimg, double* xs_s, double* mu_out) {
    const int d = P.d, np = G.np;
    const int c = btid;  // 0..255
    double xc[kPredictMaxDimRegs];
    {
        const long long gi = c0 + c;
#pragma unroll
        for (int j = 0; j < kPredictMaxDimRegs; ++j) {
            double v = 0.0;
            if (j < d && gi < P.m) {
                v = candidate_coord(P, gi, j);
                if (G.xform && G.xform[j] == B200BO_XFORM_ROUND) v = rint(v);
                v = v / G.ls[j];
            }
            xc[j] = v;
        }
    }
    const int chunk_pieces = PA_CHUNK * d / 2;
    auto load_chunk = [&](int buf, int ch) {
        const double* src = G.Xs + (size_t)ch * PA_CHUNK * d;
        double* dst = xs_s + (size_t)buf * PA_CHUNK * kPredictMaxDimRegs;
        for (int q = btid; q < chunk_pieces; q += TC2_NB) cp_async16_cg(dst + 2 * q, src + 2 * q);
    };
    const int nch = np / PA_CHUNK;
    load_chunk(0, 0);
    cp_async_commit();
    double mu_acc = 0.0;
    constexpr int R = 8;
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) load_chunk((ch + 1) & 1, ch + 1);
        cp_async_commit();
        cp_async_wait<1>();
        tc::named_bar_sync(2, TC2_NB);
        const double* xs = xs_s + (size_t)(ch & 1) * PA_CHUNK * kPredictMaxDimRegs;
        for (int r0 = 0; r0 < PA_CHUNK; r0 += R) {
            double r2[R];
#pragma unroll
            for (int q = 0; q < R; ++q) r2[q] = 0.0;
            if ((d & 1) == 0) {
#pragma unroll
                for (int j = 0; j < kPredictMaxDimRegs; j += 2) {
                    if (j < d) {
#pragma unroll
                        for (int q = 0; q < R; ++q) {
                            const double2 xv = *reinterpret_cast<const double2*>(xs + (r0 + q) * d + j);
                            const double d0 = xc[j] - xv.x, d1 = xc[j + 1] - xv.y;
                            r2[q] = fma(d0, d0, r2[q]);
                            r2[q] = fma(d1, d1, r2[q]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < kPredictMaxDimRegs; ++j) {
                    if (j < d) {
#pragma unroll
                        for (int q = 0; q < R; ++q) {
                            const double df = xc[j] - xs[(r0 + q) * d + j];
                            r2[q] = fma(df, df, r2[q]);
                        }
                    }
                }
            }
            float hi[R], lo[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int n = ch * PA_CHUNK + r0 + q;
                double kv = G.constv * cov_eval<COV>(r2[q]);
                if (n >= G.n) kv = 0.0;
                hi[q] = tc::to_tf32((float)kv);
                lo[q] = tc::to_tf32((float)(kv - (double)hi[q]));
                mu_acc = fma(__ldg(G.alphav + n), kv, mu_acc);
            }
            const int n0 = ch * PA_CHUNK + r0;
            uint8_t* img = Bimg + (size_t)(n0 >> 5) * (2 * T3_BIMG);
#pragma unroll
            for (int h4 = 0; h4 < 2; ++h4) {
                const int off = t3_bimg_offset(c, (n0 & 31) + 4 * h4);
                *reinterpret_cast<float4*>(img + off) =
                    make_float4(hi[4 * h4], hi[4 * h4 + 1], hi[4 * h4 + 2], hi[4 * h4 + 3]);
                *reinterpret_cast<float4*>(img + T3_BIMG + off) =
                    make_float4(lo[4 * h4], lo[4 * h4 + 1], lo[4 * h4 + 2], lo[4 * h4 + 3]);
            }
        }
        tc::named_bar_sync(2, TC2_NB);
    }
    cp_async_wait<0>();
    *mu_out = mu_acc;
}

__global__ void __launch_bounds__(TC2_NT, 1) predict_acq_tc3_kernel(const PredictParams P) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double mu_s[2][T3N];  // [job parity][candidate]
    __shared__ float red_s[4][T3N];
    __shared__ uint64_t full_bar[T3_STAGES], empty_bar[T3_STAGES], accfull_bar[2], accempty_bar[2];
    __shared__ uint64_t bready_bar[2], jobdone_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ SelShared sel_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* stage_mem = reinterpret_cast<uint8_t*>(smem);
    double* xs_s = reinterpret_cast<double*>(stage_mem + T3_STAGES * T3_STAGE_BYTES);
    uint8_t* scratch = reinterpret_cast<uint8_t*>(P.scratch + (long long)blockIdx.x * P.scratch_stride);
    const size_t buf_bytes = (size_t)P.scratch_stride * 4;  // two buffers of scratch_stride*4 bytes each
    const long long ntiles = (P.m + T3N - 1) / T3N;
    const long long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long njobs = my_tiles * P.n_gps;
    constexpr int KT_PER_BLOCK = PBM / tc::kTcK;  // 4 k-tiles per 128 rows
    constexpr uint32_t AIMG2 = 2 * tc::kTcImgBytes, BIMG2 = 2 * T3_BIMG;

    if (tid == 0) {
        for (int s = 0; s < T3_STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&accfull_bar[b], 1);
            tc::mbar_init(&accempty_bar[b], 4);
            tc::mbar_init(&bready_bar[b], TC2_NB / 32);
            tc::mbar_init(&jobdone_bar[b], 4);
        }
        tc::mbar_fence_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base_s, T3_TMEM_COLS);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ------------------------------ producer -------------------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            for (long long j = 0; j < njobs; ++j) {
                const GpDev& G = P.gp[j % P.n_gps];
                const int nb = G.np / PBM, nkt_row = G.np / tc::kTcK;
                const uint8_t* Bimg = scratch + (size_t)(j & 1) * buf_bytes;
                tc::mbar_wait(&bready_bar[j & 1], (uint32_t)((j >> 1) & 1));
                for (int ib = 0; ib < nb; ++ib) {
                    const int nkt = (ib + 1) * KT_PER_BLOCK;
                    const uint8_t* A = G.linv_tc + (size_t)ib * nkt_row * AIMG2;
                    for (int ht = 0; ht < 2 * nkt; ++ht, ++it) {
                        const int kt = ht >> 1;
                        const int s = it % T3_STAGES;
                        tc::mbar_wait(&empty_bar[s], ((it / T3_STAGES) & 1) ^ 1);
                        tc::mbar_arrive_expect_tx(&full_bar[s], T3_STAGE_BYTES);
                        uint8_t* dst = stage_mem + (size_t)s * T3_STAGE_BYTES;
                        const uint8_t* ap = A + (size_t)kt * AIMG2 + (size_t)(ht & 1) * T3_AHALF;
                        const uint8_t* bp = Bimg + (size_t)kt * BIMG2 + (size_t)(ht & 1) * T3_BHALF;
                        tc::bulk_g2s(dst, ap, T3_AHALF, &full_bar[s]);
                        tc::bulk_g2s(dst + T3_AHALF, ap + tc::kTcImgBytes, T3_AHALF, &full_bar[s]);
                        tc::bulk_g2s(dst + 2 * T3_AHALF, bp, T3_BHALF, &full_bar[s]);
                        tc::bulk_g2s(dst + 2 * T3_AHALF + T3_BHALF, bp + T3_BIMG, T3_BHALF, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ tcgen05.mma issuer ---------------------------------
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_tf32(128, T3N);
            uint32_t it = 0, ai = 0;
            for (long long j = 0; j < njobs; ++j) {
                const GpDev& G = P.gp[j % P.n_gps];
                const int nb = G.np / PBM;
                for (int ib = 0; ib < nb; ++ib, ++ai) {
                    const int nkt = (ib + 1) * KT_PER_BLOCK;
                    const uint32_t buf = ai & 1;
                    tc::mbar_wait(&accempty_bar[buf], ((ai >> 1) & 1) ^ 1);
                    tc::tc_fence_after_sync();
                    const uint32_t d0 = tmem_base + buf * T3N;
                    for (int ht = 0; ht < 2 * nkt; ++ht, ++it) {
                        const int s = it % T3_STAGES;
                        tc::mbar_wait(&full_bar[s], (it / T3_STAGES) & 1);
                        tc::tc_fence_after_sync();
                        const uint32_t base = tc::smem_u32(stage_mem + (size_t)s * T3_STAGE_BYTES);
#pragma unroll
                        for (int k8 = 0; k8 < 2; ++k8) {
                            const uint32_t koa = k8 * 2 * tc::kTcLBO, kob = k8 * 2 * T3_BLBO;
                            const uint64_t a_hi = tc::umma_desc_kmajor_noswz(base + koa, tc::kTcLBO, tc::kTcSBO);
                            const uint64_t a_lo = tc::umma_desc_kmajor_noswz(base + T3_AHALF + koa, tc::kTcLBO, tc::kTcSBO);
                            const uint64_t b_hi = tc::umma_desc_kmajor_noswz(base + 2 * T3_AHALF + kob, T3_BLBO, tc::kTcSBO);
                            const uint64_t b_lo =
                                tc::umma_desc_kmajor_noswz(base + 2 * T3_AHALF + T3_BHALF + kob, T3_BLBO, tc::kTcSBO);
                            tc::umma_tf32(d0, a_hi, b_hi, idesc, (ht | k8) ? 1u : 0u);
                            tc::umma_tf32(d0, a_hi, b_lo, idesc, 1u);
                            tc::umma_tf32(d0, a_lo, b_hi, idesc, 1u);
                        }
                        tc::umma_commit(&empty_bar[s]);
                    }
                    tc::umma_commit(&accfull_bar[buf]);
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------ epilogue --------------------------------------------
        const int q = warp & 3, etid = tid - 128;
        uint32_t ai = 0;
        double base_neg[2] = {0.0, 0.0}, prod[2] = {1.0, 1.0};
        if (P.sel_cta) {
            runsel_begin(sel_s, P.sel_cta + blockIdx.x, P.sel_resume, etid);
            tc::named_bar_sync(1, 128);
        }
        for (long long j = 0; j < njobs; ++j) {
            const int g = (int)(j % P.n_gps);
            const GpDev& G = P.gp[g];
            const long long tile = blockIdx.x + (j / P.n_gps) * gridDim.x;
            const int nb = G.np / PBM;
            float csum[8];  // columns lane + 32 * c
#pragma unroll
            for (int c = 0; c < 8; ++c) csum[c] = 0.f;
            for (int ib = 0; ib < nb; ++ib, ++ai) {
                const uint32_t buf = ai & 1;
                tc::mbar_wait(&accfull_bar[buf], (ai >> 1) & 1);
                tc::tc_fence_after_sync();
                const uint32_t taddr = tmem_base + buf * T3N + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    uint32_t r[32];
                    tc::tmem_ld_32x32(taddr + cc * 32, r);
                    tc::tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float x = __uint_as_float(r[i]);
                        v[i] = x * x;
                    }
                    // warp transpose-reduce: afterwards lane L holds the sum over the 32 rows of column L
#pragma unroll
                    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
                        for (int i = 0; i < s; ++i) {
                            const bool up = (lane & s) != 0;
                            const float send = up ? v[i] : v[i + s];
                            const float recv = __shfl_xor_sync(0xffffffffu, send, s);
                            v[i] = (up ? v[i + s] : v[i]) + recv;
                        }
                    }
                    csum[cc] += v[0];
                }
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&accempty_bar[buf]);
            }
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) red_s[q][cc * 32 + lane] = csum[cc];
            tc::named_bar_sync(1, 128);
            tc::mbar_wait(&bready_bar[j & 1], (uint32_t)((j >> 1) & 1));  // acquire the builders' mean
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = etid + 128 * h;
                const double colsq =
                    (((double)red_s[0][c] + (double)red_s[1][c]) + (double)red_s[2][c]) + (double)red_s[3][c];
                double val = 0.0;
                candidate_epilogue(P, G, g, mu_s[j & 1][c], colsq, tile * T3N + c, base_neg[h], prod[h], &val);
                if (P.sel_cta && g == P.n_gps - 1)
                    runsel_update<1>(sel_s, P.sel_k, etid, val, tile * T3N + c + P.index_base, tile * T3N + c < P.m);
            }
            tc::named_bar_sync(1, 128);
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&jobdone_bar[j & 1]);
        }
        if (P.sel_cta) runsel_store(sel_s, P.sel_cta + blockIdx.x, etid);
    } else if (warp >= 8) {
        // ------------------------------ builders ---------------------------------------------
        const int btid = tid - 256;
        for (long long j = 0; j < njobs; ++j) {
            const int g = (int)(j % P.n_gps);
            const GpDev& G = P.gp[g];
            const long long tile = blockIdx.x + (j / P.n_gps) * gridDim.x;
            tc::mbar_wait(&jobdone_bar[j & 1], (uint32_t)(((j >> 1) & 1) ^ 1));  // buffer j&1 free again
            uint8_t* Bimg = scratch + (size_t)(j & 1) * buf_bytes;
            double mu = 0.0;
            switch (cov_code(G.family, G.nu)) {
                case 0: tc3_build_job<0>(P, G, tile * T3N, btid, Bimg, xs_s, &mu); break;
                case 1: tc3_build_job<1>(P, G, tile * T3N, btid, Bimg, xs_s, &mu); break;
                case 2: tc3_build_job<2>(P, G, tile * T3N, btid, Bimg, xs_s, &mu); break;
                default: tc3_build_job<3>(P, G, tile * T3N, btid, Bimg, xs_s, &mu); break;
            }
            mu_s[j & 1][btid] = mu;
            tc::fence_proxy_async_global();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bready_bar[j & 1]);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, T3_TMEM_COLS);
}


// ---------------------------------------------------------------------------------------
// predict_acq_tc4_kernel = tc3 with the row blocks processed in PAIRS against each K* stage again (as tc2 does):
// ncu on tc3 showed 573 GB of DRAM reads per launch (67 % of the HBM peak, L2 hit rate 20 %) - un-paired, every row
// block re-streams the K* images.  Two 128 x 256 accumulators fill all 512 TMEM columns, so the accumulator is single-
// buffered (the issuer waits for the epilogue between pairs: ~3 us per pair of 50-200 us); a stage =
// [A0 hi|lo][A1 hi|lo][B hi|lo] = 64 KiB, three stages.  B200BO_TC_VARIANT=4.
// ---------------------------------------------------------------------------------------
constexpr int T4_STAGES = 3;
constexpr int T4_STAGE_BYTES = 4 * T3_AHALF + 2 * T3_BHALF;  // 65536
constexpr int kPredictSmemBytesTc4 = T4_STAGES * T4_STAGE_BYTES + 2 * PA_CHUNK * kPredictMaxDimRegs * 8;  // 212992
__global__ void __launch_bounds__(TC2_NT, 1) predict_acq_tc4_kernel(const PredictParams P) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double mu_s[2][T3N];  // [job parity][candidate]
    __shared__ float red_s[4][T3N];
    __shared__ uint64_t full_bar[T4_STAGES], empty_bar[T4_STAGES], accfull_bar[1], accempty_bar[1];
    __shared__ uint64_t bready_bar[2], jobdone_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ SelShared sel_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* stage_mem = reinterpret_cast<uint8_t*>(smem);
    double* xs_s = reinterpret_cast<double*>(stage_mem + T4_STAGES * T4_STAGE_BYTES);
    uint8_t* scratch = reinterpret_cast<uint8_t*>(P.scratch + (long long)blockIdx.x * P.scratch_stride);
    const size_t buf_bytes = (size_t)P.scratch_stride * 4;  // two buffers of scratch_stride*4 bytes each
    const long long ntiles = (P.m + T3N - 1) / T3N;
    const long long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long njobs = my_tiles * P.n_gps;
    constexpr int KT_PER_BLOCK = PBM / tc::kTcK;  // 4 k-tiles per 128 rows
    constexpr uint32_t AIMG2 = 2 * tc::kTcImgBytes, BIMG2 = 2 * T3_BIMG;

    if (tid == 0) {
        for (int s = 0; s < T4_STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(&accfull_bar[0], 1);
        tc::mbar_init(&accempty_bar[0], 4);
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&bready_bar[b], TC2_NB / 32);
            tc::mbar_init(&jobdone_bar[b], 4);
        }
        tc::mbar_fence_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base_s, T3_TMEM_COLS);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ------------------------------ producer -------------------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            for (long long j = 0; j < njobs; ++j) {
                const GpDev& G = P.gp[j % P.n_gps];
                const int nb = G.np / PBM, nkt_row = G.np / tc::kTcK;
                const uint8_t* Bimg = scratch + (size_t)(j & 1) * buf_bytes;
                tc::mbar_wait(&bready_bar[j & 1], (uint32_t)((j >> 1) & 1));
                for (int ib0 = 0; ib0 < nb; ib0 += 2) {
                    const bool two = ib0 + 1 < nb;
                    const int nkt0 = (ib0 + 1) * KT_PER_BLOCK, nkt = two ? nkt0 + KT_PER_BLOCK : nkt0;
                    const uint8_t* A0 = G.linv_tc + (size_t)ib0 * nkt_row * AIMG2;
                    const uint8_t* A1 = A0 + (size_t)nkt_row * AIMG2;
                    for (int ht = 0; ht < 2 * nkt; ++ht, ++it) {
                        const int kt = ht >> 1;
                        const int s = it % T4_STAGES;
                        const bool a0 = kt < nkt0;
                        tc::mbar_wait(&empty_bar[s], ((it / T4_STAGES) & 1) ^ 1);
                        tc::mbar_arrive_expect_tx(&full_bar[s], ((a0 ? 2u : 0u) + (two ? 2u : 0u)) * T3_AHALF + 2u * T3_BHALF);
                        uint8_t* dst = stage_mem + (size_t)s * T4_STAGE_BYTES;
                        const size_t ah = (size_t)kt * AIMG2 + (size_t)(ht & 1) * T3_AHALF;
                        const uint8_t* bp = Bimg + (size_t)kt * BIMG2 + (size_t)(ht & 1) * T3_BHALF;
                        if (a0) {
                            tc::bulk_g2s(dst, A0 + ah, T3_AHALF, &full_bar[s]);
                            tc::bulk_g2s(dst + T3_AHALF, A0 + ah + tc::kTcImgBytes, T3_AHALF, &full_bar[s]);
                        }
                        if (two) {
                            tc::bulk_g2s(dst + 2 * T3_AHALF, A1 + ah, T3_AHALF, &full_bar[s]);
                            tc::bulk_g2s(dst + 3 * T3_AHALF, A1 + ah + tc::kTcImgBytes, T3_AHALF, &full_bar[s]);
                        }
                        tc::bulk_g2s(dst + 4 * T3_AHALF, bp, T3_BHALF, &full_bar[s]);
                        tc::bulk_g2s(dst + 4 * T3_AHALF + T3_BHALF, bp + T3_BIMG, T3_BHALF, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ tcgen05.mma issuer ---------------------------------
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_tf32(128, T3N);
            uint32_t it = 0, ai = 0;
            for (long long j = 0; j < njobs; ++j) {
                const GpDev& G = P.gp[j % P.n_gps];
                const int nb = G.np / PBM;
                for (int ib0 = 0; ib0 < nb; ib0 += 2, ++ai) {
                    const bool two = ib0 + 1 < nb;
                    const int nkt0 = (ib0 + 1) * KT_PER_BLOCK, nkt = two ? nkt0 + KT_PER_BLOCK : nkt0;
                    tc::mbar_wait(&accempty_bar[0], (ai & 1) ^ 1);  // ONE pair of accumulators: the epilogue drains it first
                    tc::tc_fence_after_sync();
                    const uint32_t d0 = tmem_base, d1 = tmem_base + T3N;
                    for (int ht = 0; ht < 2 * nkt; ++ht, ++it) {
                        const int kt = ht >> 1;
                        const int s = it % T4_STAGES;
                        const bool a0 = kt < nkt0;
                        tc::mbar_wait(&full_bar[s], (it / T4_STAGES) & 1);
                        tc::tc_fence_after_sync();
                        const uint32_t base = tc::smem_u32(stage_mem + (size_t)s * T4_STAGE_BYTES);
#pragma unroll
                        for (int k8 = 0; k8 < 2; ++k8) {
                            const uint32_t koa = k8 * 2 * tc::kTcLBO, kob = k8 * 2 * T3_BLBO;
                            const uint64_t b_hi = tc::umma_desc_kmajor_noswz(base + 4 * T3_AHALF + kob, T3_BLBO, tc::kTcSBO);
                            const uint64_t b_lo =
                                tc::umma_desc_kmajor_noswz(base + 4 * T3_AHALF + T3_BHALF + kob, T3_BLBO, tc::kTcSBO);
                            if (a0) {
                                const uint64_t a_hi = tc::umma_desc_kmajor_noswz(base + koa, tc::kTcLBO, tc::kTcSBO);
                                const uint64_t a_lo = tc::umma_desc_kmajor_noswz(base + T3_AHALF + koa, tc::kTcLBO, tc::kTcSBO);
                                tc::umma_tf32(d0, a_hi, b_hi, idesc, (ht | k8) ? 1u : 0u);
                                tc::umma_tf32(d0, a_hi, b_lo, idesc, 1u);
                                tc::umma_tf32(d0, a_lo, b_hi, idesc, 1u);
                            }
                            if (two) {
                                const uint64_t a_hi = tc::umma_desc_kmajor_noswz(base + 2 * T3_AHALF + koa, tc::kTcLBO, tc::kTcSBO);
                                const uint64_t a_lo = tc::umma_desc_kmajor_noswz(base + 3 * T3_AHALF + koa, tc::kTcLBO, tc::kTcSBO);
                                tc::umma_tf32(d1, a_hi, b_hi, idesc, (ht | k8) ? 1u : 0u);
                                tc::umma_tf32(d1, a_hi, b_lo, idesc, 1u);
                                tc::umma_tf32(d1, a_lo, b_hi, idesc, 1u);
                            }
                        }
                        tc::umma_commit(&empty_bar[s]);
                    }
                    tc::umma_commit(&accfull_bar[0]);
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------ epilogue --------------------------------------------
        const int q = warp & 3, etid = tid - 128;
        uint32_t ai = 0;
        double base_neg[2] = {0.0, 0.0}, prod[2] = {1.0, 1.0};
        if (P.sel_cta) {
            runsel_begin(sel_s, P.sel_cta + blockIdx.x, P.sel_resume, etid);
            tc::named_bar_sync(1, 128);
        }
        for (long long j = 0; j < njobs; ++j) {
            const int g = (int)(j % P.n_gps);
            const GpDev& G = P.gp[g];
            const long long tile = blockIdx.x + (j / P.n_gps) * gridDim.x;
            const int nb = G.np / PBM;
            float csum[8];  // columns lane + 32 * c
#pragma unroll
            for (int c = 0; c < 8; ++c) csum[c] = 0.f;
            for (int ib0 = 0; ib0 < nb; ib0 += 2, ++ai) {
                const bool two = ib0 + 1 < nb;
                tc::mbar_wait(&accfull_bar[0], ai & 1);
                tc::tc_fence_after_sync();
                for (int rb = 0; rb < (two ? 2 : 1); ++rb) {
                const uint32_t taddr = tmem_base + rb * T3N + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    uint32_t r[32];
                    tc::tmem_ld_32x32(taddr + cc * 32, r);
                    tc::tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float x = __uint_as_float(r[i]);
                        v[i] = x * x;
                    }
                    // warp transpose-reduce: afterwards lane L holds the sum over the 32 rows of column L
#pragma unroll
                    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
                        for (int i = 0; i < s; ++i) {
                            const bool up = (lane & s) != 0;
                            const float send = up ? v[i] : v[i + s];
                            const float recv = __shfl_xor_sync(0xffffffffu, send, s);
                            v[i] = (up ? v[i + s] : v[i]) + recv;
                        }
                    }
                    csum[cc] += v[0];
                }
                }
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&accempty_bar[0]);
            }
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) red_s[q][cc * 32 + lane] = csum[cc];
            tc::named_bar_sync(1, 128);
            tc::mbar_wait(&bready_bar[j & 1], (uint32_t)((j >> 1) & 1));  // acquire the builders' mean
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = etid + 128 * h;
                const double colsq =
                    (((double)red_s[0][c] + (double)red_s[1][c]) + (double)red_s[2][c]) + (double)red_s[3][c];
                double val = 0.0;
                candidate_epilogue(P, G, g, mu_s[j & 1][c], colsq, tile * T3N + c, base_neg[h], prod[h], &val);
                if (P.sel_cta && g == P.n_gps - 1)
                    runsel_update<1>(sel_s, P.sel_k, etid, val, tile * T3N + c + P.index_base, tile * T3N + c < P.m);
            }
            tc::named_bar_sync(1, 128);
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&jobdone_bar[j & 1]);
        }
        if (P.sel_cta) runsel_store(sel_s, P.sel_cta + blockIdx.x, etid);
    } else if (warp >= 8) {
        // ------------------------------ builders ---------------------------------------------
        const int btid = tid - 256;
        for (long long j = 0; j < njobs; ++j) {
            const int g = (int)(j % P.n_gps);
            const GpDev& G = P.gp[g];
            const long long tile = blockIdx.x + (j / P.n_gps) * gridDim.x;
            tc::mbar_wait(&jobdone_bar[j & 1], (uint32_t)(((j >> 1) & 1) ^ 1));  // buffer j&1 free again
            uint8_t* Bimg = scratch + (size_t)(j & 1) * buf_bytes;
            double mu = 0.0;
            switch (cov_code(G.family, G.nu)) {
                case 0: tc3_build_job<0>(P, G, tile * T3N, btid, Bimg, xs_s, &mu); break;
                case 1: tc3_build_job<1>(P, G, tile * T3N, btid, Bimg, xs_s, &mu); break;
                case 2: tc3_build_job<2>(P, G, tile * T3N, btid, Bimg, xs_s, &mu); break;
                default: tc3_build_job<3>(P, G, tile * T3N, btid, Bimg, xs_s, &mu); break;
            }
            mu_s[j & 1][btid] = mu;
            tc::fence_proxy_async_global();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bready_bar[j & 1]);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, T3_TMEM_COLS);
}


}  // namespace b200bo
