// Shared definitions for the sm_100a SG_PR kernels.
#pragma once
#ifdef SGPR_EMU                      // tests/emu: the same source built by the host compiler (test infrastructure only)
#include "../../tests/emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace sgpr {

#ifndef SGPR_THREADS
#define SGPR_THREADS 256
#endif
constexpr int kThreads = SGPR_THREADS; // one CTA (8 warps by default) works on one graph
constexpr int kWarps   = kThreads / 32;

constexpr int kLabels  = 12;           // sg_net.py:200-202
constexpr int kInCh    = 3 + kLabels;  // [15][N] channel-major input block, sg_net.py:296-299
constexpr int kF1 = 64, kF2 = 64, kF3 = 32;      // config.yml:11-13
constexpr int kT  = 16;                // tensor_neurons, config.yml:14
constexpr int kBn = 16;                // bottle_neck_neurons, config.yml:15

constexpr int XS = 68;    // row stride (floats) of node-major feature tiles: 64 + 4 keeps LDS.128 conflict-free
constexpr int YS = 132;   // row stride of the per-node GEMM output tile A|B: 128 + 4

constexpr float kSlope = 0.2f;         // LeakyReLU(negative_slope=0.2), sg_net.py:53

// Packed eval-mode parameters in device memory (built by pack.cpp, one allocation).
// EdgeConv l with C_in -> C_out: `w` is the [C_in][2*C_out] matrix
//   columns [0, C_out)        = sign[c] * W[c][ci]          (the (x_j - x_i) half of the 1x1 conv, dgcnn.py:47)
//   columns [C_out, 2*C_out)  = sign[c] * W[c][C_in + ci]   (the x_i half)
// stored in the channel-PAIR layout of pack.hpp::pair_index (row p = input channels 2p, 2p+1 interleaved per output)
// so that one FFMA2 advances the even/odd-channel partial sums of an output.
// `alpha`/`beta` are the BN-eval scale/shift with alpha made non-negative by folding its sign into `w`
// (max over neighbours then commutes exactly with BN+LeakyReLU, SURVEY §7 hard part 5).
struct PackedWeights {
    const float* s1;        // xyz layer 1: [64][8] = {wa0,wa1,wa2, wb0,wb1,wb2, alpha, beta} per output channel
    const float* w_s2;      // 64 in x 128 out (pair layout)
    const float* w_s3;      // 64 x 64
    const float* w_f1;      // 12 x 128
    const float* w_f2;      // 64 x 128
    const float* w_f3;      // 64 x 64
    const float* w_end;     // 64 x 32    transpose of dgcnn_conv_end.0.weight (pair layout)
    const float* ab_s2;     // alpha[64] beta[64]
    const float* ab_s3;     // alpha[32] beta[32]
    const float* ab_f1;     // alpha[64] beta[64]
    const float* ab_f2;     // alpha[64] beta[64]
    const float* ab_f3;     // alpha[32] beta[32]
    const float* ab_end;    // alpha[32] beta[32]  (sign NOT folded: no max follows)
    const float* att_w;     // [32][32]   attention.weight_matrix (row = input feature)
    const float* ntn_w;     // [32][512]  tensor_network.weight_matrix.view(32, 32*16): col = b*16 + t
    const float* ntn_v;     // [16][64]
    const float* ntn_b;     // [16]
    // tensor-core operand planes of the four 64-channel EdgeConv layers: [big, small][128][64] (pack.hpp::pack_edgeconv_tc)
    const float* wtc_s2;
    const float* wtc_s3;
    const float* wtc_f2;
    const float* wtc_f3;
};

// FC head, small enough to travel as a kernel parameter (constant bank, uniform access).
struct HeadParams {
    float fc1_w[kBn * kT];  // [16][16] row = output
    float fc1_b[kBn];
    float fc2_w[kBn];
    float fc2_b;
};

// Kernel launch / dynamic shared memory spelled once for nvcc and for the host-compiler emulator build of tests/emu
#ifdef SGPR_EMU
#define SGPR_LAUNCH(kern, grid, block, smem, st, ...) emu::launch((grid), (block), (smem), [=] { kern(__VA_ARGS__); })
#define SGPR_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#else
#define SGPR_LAUNCH(kern, grid, block, smem, st, ...) kern<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__)
#define SGPR_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

#ifdef SGPR_EMU
// tests/emu: an mbarrier is one 64-bit word (bit 0 = phase parity, the rest = bytes still expected); a bulk copy is a
// memcpy by the issuing thread that completes the transaction count.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t) { std::atomic_ref<uint64_t>(*bar).store(0); }
__device__ __forceinline__ void fence_mbar_init() { std::atomic_thread_fence(std::memory_order_seq_cst); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    std::atomic_ref<uint64_t>(*bar).fetch_add(static_cast<uint64_t>(bytes) << 1);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    std::memcpy(dst_smem, src_gmem, bytes);
    std::atomic_ref<uint64_t> b(*bar);
    const uint64_t left = b.fetch_sub(static_cast<uint64_t>(bytes) << 1) - (static_cast<uint64_t>(bytes) << 1);
    if ((left >> 1) == 0) b.fetch_xor(1);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    std::atomic_ref<uint64_t> b(*bar);
    while ((b.load() & 1u) == parity) std::this_thread::yield();
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) -------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

#endif  // !SGPR_EMU

__device__ __forceinline__ float lrelu(float x) { return x > 0.0f ? x : x * kSlope; }

}  // namespace sgpr
