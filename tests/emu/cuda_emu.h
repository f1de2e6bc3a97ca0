// TEST INFRASTRUCTURE ONLY — a minimal CUDA execution-model emulator for the host compiler.
//
// Purpose: this repo is developed in a container without a GPU and with a small budget of B200 minutes.  To debug the
// kernels (sg_pr_b200/csrc/*.cuh) before spending GPU time, tests/emu/build_emu.py compiles the SAME kernel and host
// source with g++ -DSGPR_EMU against this header into tests/emu/libsgpr_emu.so; every CUDA thread of a block becomes a
// std::thread, __syncthreads / __syncwarp / warp shuffles become barriers.  Blocks run one after the other.  Nothing in
// sg_pr_b200/ ever loads that library by itself (sg_pr_b200/_lib.py binds libsgpr_b200.so only and fails without it): it
// exists so that tests/test_train_emu.py and tests/test_eval_emu.py can check the kernels' logic against the reference's
// golden vectors on the CPU.
//
// Supported subset: 1-D blocks, 1-D / 2-D grids, full-mask warp primitives, static and dynamic shared memory, atomicAdd on
// int / float / double, the cudaMalloc / cudaMemcpy / cudaMemset family on host memory, streams as no-ops.  mbarriers
// and bulk (TMA) copies are emulated in csrc/common.cuh under SGPR_EMU.
#pragma once
#include <atomic>
#include <barrier>
#include <math.h>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)
#define __shared__ static          // blocks run sequentially, so one static instance per kernel IS the block's copy

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() = default; dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace emu {
struct Block {
    unsigned nthreads;
    std::barrier<> all;
    std::vector<std::unique_ptr<std::barrier<>>> warp;
    std::vector<std::array<uint64_t, 32>> xch;
    unsigned char* dyn = nullptr;
    explicit Block(unsigned n) : nthreads(n), all(n) {
        for (unsigned w = 0; w < (n + 31) / 32; ++w) {
            const unsigned lanes = (w + 1) * 32 <= n ? 32 : n - w * 32;
            warp.emplace_back(new std::barrier<>(lanes));
            xch.emplace_back();
        }
    }
};
inline thread_local Block* blk = nullptr;
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::blk->all.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::blk->warp[threadIdx.x >> 5]->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <typename T>
inline T emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    auto& x = emu::blk->xch[threadIdx.x >> 5];
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    x[threadIdx.x & 31] = raw;
    __syncwarp();
    raw = x[src_lane & 31];
    __syncwarp();
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int lane) { return emu_exchange(v, lane); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int d) { return emu_exchange(v, (threadIdx.x & 31) ^ d); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) {
    const int l = threadIdx.x & 31;
    return emu_exchange(v, l + d < 32 ? l + d : l);
}
inline int __reduce_max_sync(unsigned, int v) {
    for (int d = 16; d >= 1; d >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
    return v;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    for (int d = 16; d >= 1; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
inline unsigned __ballot_sync(unsigned, bool pred) {
    unsigned v = pred ? (1u << (threadIdx.x & 31)) : 0u;
    return __reduce_or_sync(0xffffffffu, v);
}
inline int __reduce_add_sync(unsigned, int v) {
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

template <typename T> inline T atomicAdd(T* p, T v) { return std::atomic_ref<T>(*p).fetch_add(v); }
inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { return std::atomic_ref<unsigned long long>(*p).fetch_or(v); }

template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::fmaxf; using std::fminf; using std::fmaf;   // expf/tanhf/logf/sqrtf come from <math.h> in the global namespace
using std::max; using std::min;

// ---- runtime subset ------------------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 4; size_t sharedMemPerBlockOptin = 232448; };
inline const char* cudaGetErrorString(cudaError_t) { return "emulator"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* c) { *c = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
struct cudaFuncAttributes { size_t sharedSizeBytes = 0; };
template <typename F> inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { *a = cudaFuncAttributes(); return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type = cudaMemoryTypeDevice; void* devicePointer = nullptr; };
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {      // every buffer is "device" memory here
    a->type = cudaMemoryTypeDevice; a->devicePointer = const_cast<void*>(p); return cudaSuccess;
}
enum { cudaStreamNonBlocking = 1 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {
// Run `body` once per CUDA thread, block after block.  One pool of `block` host threads serves every block of the
// launch; each block gets fresh barrier objects (a CUDA thread that returns early drops out of them).
inline void launch(dim3 grid3, unsigned block, size_t smem_bytes, const std::function<void()>& body) {
    const unsigned grid = grid3.x * grid3.y;
    std::vector<unsigned char> dyn(smem_bytes + 256);
    unsigned char* dyn_aligned = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 127) & ~uintptr_t(127));
    std::barrier<> outer(block);
    Block* current = nullptr;
    std::vector<std::thread> threads;
    threads.reserve(block);
    for (unsigned t = 0; t < block; ++t) {
        threads.emplace_back([&, t] {
            for (unsigned b = 0; b < grid; ++b) {
                if (t == 0) { current = new Block(block); current->dyn = dyn_aligned; }
                outer.arrive_and_wait();
                blk = current;
                threadIdx = dim3(t); blockIdx = dim3(b % grid3.x, b / grid3.x); blockDim = dim3(block); gridDim = grid3;
                body();
                // a thread that returns early must not block its peers' later barriers
                current->warp[t >> 5]->arrive_and_drop();
                current->all.arrive_and_drop();
                outer.arrive_and_wait();
                if (t == 0) delete current;
            }
        });
    }
    for (auto& th : threads) th.join();
}
inline unsigned char* dyn_smem() { return blk->dyn; }
}  // namespace emu
