// tc_test.cu - standalone check of the tcgen05 building blocks used by the fp32-mode kernel:
// bulk async copies + mbarriers, no-swizzle K-major UMMA descriptors, kind::tf32 MMA into TMEM,
// tcgen05.ld read-back.  D[128x128] = A[128xK] * B[128xK]^T with the 3xTF32 split
// (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi), compared with an fp64 host reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tc_test tools/tc_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../bayesianoptimization_b200/csrc/tc_common.cuh"

using namespace b200bo::tc;

constexpr int STAGES = 2;
constexpr int IMG = kTcImgBytes;  // 16 KiB per operand image (128 rows x 32 k fp32)

// global layout: for each k-tile kt: [A_hi][A_lo][B_hi][B_lo] images, 64 KiB per k-tile
__global__ void __launch_bounds__(192) tc_gemm_kernel(const uint8_t* __restrict__ imgs, int nkt, int nsplit,
                                                      float* __restrict__ D) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt % STAGES;
                const uint32_t use = kt / STAGES;
                mbar_wait(&empty_bar[s], (use & 1) ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], 4 * IMG);
                bulk_g2s(smem + (size_t)s * 4 * IMG, imgs + (size_t)kt * 4 * IMG, 4 * IMG, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 128);
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt % STAGES;
                const uint32_t use = kt / STAGES;
                mbar_wait(&full_bar[s], use & 1);
                tc_fence_after_sync();
                const uint32_t base = smem_u32(smem + (size_t)s * 4 * IMG);
                for (int j = 0; j < kTcK / 8; ++j) {
                    const uint32_t koff = j * 2 * kTcLBO;  // 8 k = two core matrices along K
                    const uint64_t a_hi = umma_desc_kmajor_noswz(base + 0 * IMG + koff, kTcLBO, kTcSBO);
                    const uint64_t a_lo = umma_desc_kmajor_noswz(base + 1 * IMG + koff, kTcLBO, kTcSBO);
                    const uint64_t b_hi = umma_desc_kmajor_noswz(base + 2 * IMG + koff, kTcLBO, kTcSBO);
                    const uint64_t b_lo = umma_desc_kmajor_noswz(base + 3 * IMG + koff, kTcLBO, kTcSBO);
                    umma_tf32(tmem_base, a_hi, b_hi, idesc, (kt | j) ? 1u : 0u);
                    if (nsplit == 3) {
                        umma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
                        umma_tf32(tmem_base, a_lo, b_hi, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&acc_bar);
        }
    } else {
        // epilogue warps 2..5: TMEM quadrant = warp % 4
        const int q = warp & 3;
        mbar_wait(&acc_bar, 0);
        tc_fence_after_sync();
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
            tmem_ld_wait();
            const int row = q * 32 + lane;
            for (int j = 0; j < 32; ++j) D[row * 128 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
}

static float tf32_round(float x) {  // round-to-nearest (ties away) to 10 mantissa bits
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

int main() {
    const int K = 512, nkt = K / kTcK;
    std::vector<float> A(128 * K), B(128 * K);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = (float)rand() / RAND_MAX;
    std::vector<uint8_t> imgs((size_t)nkt * 4 * IMG);
    for (int kt = 0; kt < nkt; ++kt)
        for (int r = 0; r < 128; ++r)
            for (int k = 0; k < kTcK; ++k) {
                const float a = A[r * K + kt * kTcK + k], b = B[r * K + kt * kTcK + k];
                const float ah = tf32_round(a), al = tf32_round(a - ah);
                const float bh = tf32_round(b), bl = tf32_round(b - bh);
                uint8_t* base = imgs.data() + (size_t)kt * 4 * IMG;
                const int off = tc_img_offset(r, k);
                memcpy(base + 0 * IMG + off, &ah, 4);
                memcpy(base + 1 * IMG + off, &al, 4);
                memcpy(base + 2 * IMG + off, &bh, 4);
                memcpy(base + 3 * IMG + off, &bl, 4);
            }
    uint8_t* d_imgs;
    float* d_D;
    cudaMalloc(&d_imgs, imgs.size());
    cudaMalloc(&d_D, 128 * 128 * 4);
    cudaMemcpy(d_imgs, imgs.data(), imgs.size(), cudaMemcpyHostToDevice);
    const int smem_bytes = STAGES * 4 * IMG;
    cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    printf("{");
    for (int nsplit = 1; nsplit <= 3; nsplit += 2) {
        cudaMemset(d_D, 0, 128 * 128 * 4);
        tc_gemm_kernel<<<1, 192, smem_bytes>>>(d_imgs, nkt, nsplit, d_D);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> D(128 * 128);
        cudaMemcpy(D.data(), d_D, D.size() * 4, cudaMemcpyDeviceToHost);
        double max_err = 0, max_ref = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
                max_err = fmax(max_err, fabs(ref - D[m * 128 + n]));
                max_ref = fmax(max_ref, fabs(ref));
            }
        printf("\"split%d\": {\"cuda\": \"%s\", \"max_abs_err\": %.3e, \"max_ref\": %.3e, \"d00\": %.6f}%s", nsplit,
               cudaGetErrorString(e), max_err, max_ref, D[0], nsplit == 1 ? ", " : "");
    }
    printf("}\n");
    return 0;
}
