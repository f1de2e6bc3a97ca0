// Microbenchmark: issue rate of FMNMX (2-input fp32 max), FMNMX3 (3-input, sm_100+), VIMNMX3 (3-input s32 max) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_alu tools/microbench_alu.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__device__ __forceinline__ float max3f(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int MODE>
__global__ void k_max(float* out) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i + blockIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) acc[i] = fmaxf(acc[i], acc[(i + 5) & 15]);
            else if (MODE == 1) acc[i] = max3f(acc[i], acc[(i + 5) & 15], acc[(i + 9) & 15]);
            else {
                int a = __float_as_int(acc[i]), b = __float_as_int(acc[(i + 5) & 15]), c = __float_as_int(acc[(i + 9) & 15]);
                acc[i] = __int_as_float(__vimax3_s32(a, b, c));
            }
        }
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
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    const double n = double(blocks) * threads * ITERS * 16;
    float t0 = time_ms([&] { k_max<0><<<blocks, threads>>>(out); });
    float t1 = time_ms([&] { k_max<1><<<blocks, threads>>>(out); });
    float t2 = time_ms([&] { k_max<2><<<blocks, threads>>>(out); });
    auto rate = [&](float t) { return n / (t * 1e-3) / p.multiProcessorCount / (clk_khz * 1e3); };
    printf("{\"device\": \"%s\", \"nominal_clock_mhz\": %d, \"fmnmx_lanes_per_clk_per_sm\": %.1f, \"fmnmx3_lanes_per_clk_per_sm\": %.1f, "
           "\"vimnmx3_lanes_per_clk_per_sm\": %.1f}\n", p.name, clk_khz / 1000, rate(t0), rate(t1), rate(t2));
    return 0;
}
