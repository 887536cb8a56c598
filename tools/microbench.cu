// microbench.cu - measured fp64 ceilings on this B200 (roofline denominators that
// MEASURED_PEAKS.json does not carry): DFMA issue rate and DMMA (mma.sync f64) rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16816_kernel(double* out, int iters, double a, double b) {
    double c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(b), "d"(b), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half the warps issue DMMA, the other half DFMA: do the two fp64 paths share execution units?
__global__ void mixed_kernel(double* out, int iters, double a, double b) {
    const int warp = threadIdx.x >> 5;
    double acc = 0;
    if (warp & 1) {
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += x[i];
    } else {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += c[i][0] + c[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"clock_khz\": %d", p.name, sms, p.l2CacheSize, p.clockRate);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32, blocks = sms * (warps <= 8 ? 2 : 1);
        float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double tf = 2.0 * 16 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
        printf(", \"dfma_tflops_w%d\": %.2f", warps * (warps <= 8 ? 2 : 1), tf);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        int threads = warps * 32, blocks = sms;
        float ms = time_ms([&] { dmma884_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double tf = 2.0 * 8 * 256 * iters * (double)warps * blocks / (ms * 1e-3) / 1e12;
        printf(", \"dmma884_tflops_w%d\": %.2f", warps, tf);
        ms = time_ms([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
        tf = 2.0 * 8 * 2048 * (iters / 4) * (double)warps * blocks / (ms * 1e-3) / 1e12;
        printf(", \"dmma16816_tflops_w%d\": %.2f", warps, tf);
    }
    {
        // 16 warps/SM: 8 DMMA warps (8 mma x 256 FMA per iter) + 8 DFMA warps (16 x 32 FMA per iter)
        int threads = 512, blocks = sms;
        float ms = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double tf_dmma = 2.0 * 8 * 256 * iters * 8.0 * blocks / (ms * 1e-3) / 1e12;
        double tf_dfma = 2.0 * 16 * 32 * iters * 8.0 * blocks / (ms * 1e-3) / 1e12;
        printf(", \"mixed_ms\": %.3f, \"mixed_dmma_tflops\": %.2f, \"mixed_dfma_tflops\": %.2f", ms, tf_dmma, tf_dfma);
    }
    printf("}\n");
    return 0;
}
