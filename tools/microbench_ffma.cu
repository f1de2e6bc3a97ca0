// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a, plus SHFL+FMNMX rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_ffma tools/microbench_ffma.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_shfl(float* out) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 0.37f + i;
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float p = __shfl_xor_sync(0xffffffffu, v[i], 1 + (it & 15));
            v[i] = (threadIdx.x & 1) ? fminf(v[i], p) : fmaxf(v[i], p) + 1.0f;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fmnmx(float* out, float a) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fminf(fmaxf(acc[i], a), acc[(i + 1) & 15] + 1.0f);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    const double n = double(blocks) * threads * ITERS * 16;
    float t1 = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t3 = time_ms([&] { k_shfl<<<blocks, threads>>>(out); });
    float t4 = time_ms([&] { k_fmnmx<<<blocks, threads>>>(out, 0.5f); });
    printf("{\"fmnmx_fmnmx_fadd_lane_triplets_per_clk_per_sm\": %.2f}\n", n / (t4 * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
    printf("{\"device\": \"%s\", \"sms\": %d, \"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"ffma_ms\": %.4f, \"ffma2_ms\": %.4f, "
           "\"shfl_minmax_gops\": %.1f}\n",
           p.name, p.multiProcessorCount, 2 * n / t1 / 1e9, 2 * n / t2 / 1e9, t1, t2,
           double(blocks) * threads * (ITERS / 4) * 8 / t3 / 1e6);
    return 0;
}
