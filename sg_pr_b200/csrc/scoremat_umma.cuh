// All-pairs pair head on the 5th-generation tensor cores (tcgen05 / TMEM): the score matrix of BASELINE config 4.
//
// Reference math, per ordered pair (i = row graph = side 1, j = column graph = side 2):
//     s[t]  = sum_b (e_i^T W)[b][t] * e_j[b]                      TenorNetworkModule, /root/reference/layers_batch.py:78-79
//     z[t]  = relu(s[t] + (V [e_i ; e_j])[t] + bias[t])           layers_batch.py:80-82
//     score = sigmoid(w2 . relu(W1 z + b1) + b2)                  /root/reference/sg_net.py:131-136
// Over all pairs the first line is a dense contraction (the one genuine GEMM of the path):
//     D[j][i*16 + t] = sum_b E[j][b] * P[i*16 + t][b],     E = pooled column vectors [M][32],  P = e_i^T W  [R*16][32]
// i.e. per tile an M=128 (j) x N=256 (16 row graphs x 16 neurons) x K=32 UMMA with both operands K-major.  fp32 faithfulness
// comes from the 3xTF32 split  x = big + small  (big = x rounded to TF32, small = x - big):
//     D = E_small.P_big + E_big.P_small + E_big.P_big      (12 tcgen05.mma kind::tf32 per tile, fp32 accumulate in TMEM)
// which leaves ~2^-21 relative error per product — the oracle's 1e-5 on the sigmoid is met with two orders to spare.
// With lanes = j and columns = (i, t), a thread's tcgen05.ld returns the 16 neurons of ONE pair: the whole epilogue
// (V-block, bias, relu, FC1, relu, FC2, sigmoid) is thread-local and a warp's store covers 32 consecutive j of one row.
//
// The operand planes (big / small, 128-byte-swizzled K-major, zero-padded to whole tiles) are written to global memory by
// the preparation kernel (sgpr_ntn_split_kernel below) exactly as the tensor core wants them in shared memory, so a
// stage is filled by plain 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx): no thread touches operand data.
//
// CTA = kEpiWarps + 2 warps, persistent over a contiguous range of (column block, row group) tiles:
//     warps 0..kEpiWarps-1  epilogue: TMEM -> registers -> scores (warp w owns TMEM lane quadrant w%4 and a slice of
//                           the tile's 16 row graphs)
//     warp  kEpiWarps       one thread issues the bulk copies of the next stage
//     warp  kEpiWarps+1     one thread issues the MMAs and commits to the mbarriers
// Two shared-memory stages (full/empty mbarriers) and two 256-column TMEM accumulators (tfull/tempty) overlap the three.
#pragma once
#include "common.cuh"

namespace sgpr {
namespace umma {

constexpr int kTileJ = 128;                 // UMMA M: column graphs per tile = TMEM lanes
constexpr int kTileI = 16;                  // row graphs per tile
constexpr int kUmmaN = kTileI * kT;         // 256 accumulator columns
#ifndef SGPR_UMMA_EPI_WARPS
#define SGPR_UMMA_EPI_WARPS 16
#endif
constexpr int kEpiWarps = SGPR_UMMA_EPI_WARPS;                       // 4, 8 or 16
constexpr int kRowsPerWarp = kTileI / (kEpiWarps / 4);               // row graphs of a tile handled by one epilogue warp
constexpr int kThreadsUmma = (kEpiWarps + 2) * 32;
constexpr int kABytes = kTileJ * 128;       // one TF32 plane of the A tile: 128 rows x 32 floats
constexpr int kBBytes = kUmmaN * 128;       // one plane of the B tile: 256 rows x 32 floats
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;               // 96 KB
constexpr int kSmemUmma = 2 * kStageBytes + 1024 /* alignment slack */ + 128 /* barriers + TMEM pointer */;

struct ScoreMatArgs {
    const float* cols_big;  // [ceil128(M)][32]      pooled vectors of the column graphs (side 2), TF32 big plane, swizzled
    const float* cols_small;
    const float* proj_big;  // [ceil16(R)*16][32]    row i*16+t, column b: sum_a e_i[a] W[a][b][t], big plane, swizzled
    const float* proj_small;
    const float* rowblk;    // [R][16]      first half of V [e1;e2]  + tensor_network.bias
    const float* colblk;    // [M][16]      second half
    float* scores;          // [R][ld]
    long long ld;
    int R, M;
    int n_ib;               // row groups = ceil(R / 16)
    int n_tiles;            // column blocks * row groups
};

#ifndef SGPR_EMU
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait with a watchdog: a protocol bug must fail loudly (trap -> launch error), never hang the GPU
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
    uint32_t done, polls = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++polls > (1u << 26)) __trap();
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand descriptor, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), descriptor
// version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  The tile base is 1024-byte aligned; a K step of 8 TF32 values
// advances the start address by 32 bytes inside the swizzle atom.
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1) << 16) |
           (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, N = 256, M = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kUmmaN >> 3) << 17) |
                                (static_cast<uint32_t>(kTileJ >> 4) << 24);

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Preparation: graph g of `count` (pad graphs up to `padded` produce zeros) -> operand planes in the shared-memory image.
//   role 0 (rows, side 1): 16 operand rows g*16+t with  P[t][b] = sum_a e_g[a] W[a][b][t]  (layers_batch.py:78), and
//                          blk[g][t] = sum_c V[t][c] e_g[c] + bias[t]                       (layers_batch.py:80-82)
//   role 1 (cols, side 2): 1 operand row g = e_g, and blk[g][t] = sum_c V[t][32+c] e_g[c]
// Element b of operand row r lives at float offset r*32 + (((b>>2) ^ (r&7)) << 2) + (b&3): the 128-byte swizzle, which
// depends only on r mod 8 and therefore survives the copy of whole 8-row-aligned tiles.
__global__ void __launch_bounds__(kThreads)
sgpr_ntn_split_kernel(const float* __restrict__ pooled, int count, int padded, int role, float* __restrict__ big,
                      float* __restrict__ small, float* __restrict__ blk, const PackedWeights W) {
    __shared__ float e[kF3];
    const int tid = threadIdx.x;
    for (int g = blockIdx.x; g < padded; g += gridDim.x) {
        const bool live = g < count;
        if (tid < kF3) e[tid] = live ? __ldg(pooled + static_cast<size_t>(g) * kF3 + tid) : 0.0f;
        __syncthreads();
        const int nval = role == 0 ? 512 : 32;
        for (int c = tid; c < nval; c += kThreads) {
            const int t = c >> 5, b = c & 31;                 // role 1: t == 0
            float x;
            if (role == 0) {
                x = 0.0f;
#pragma unroll 8
                for (int a = 0; a < kF3; ++a) x = fmaf(e[a], __ldg(W.ntn_w + a * 512 + b * kT + t), x);
            } else {
                x = e[b];
            }
            const size_t r = role == 0 ? static_cast<size_t>(g) * kT + t : static_cast<size_t>(g);
            const size_t off = r * 32 + ((((b >> 2) ^ static_cast<int>(r & 7)) << 2) + (b & 3));
            const float hi = tf32_rna(x);
            big[off] = hi;
            small[off] = x - hi;
        }
        if (live && tid < kT) {
            float acc = 0.0f;
            for (int c = 0; c < kF3; ++c) acc = fmaf(__ldg(W.ntn_v + tid * 64 + role * 32 + c), e[c], acc);
            blk[static_cast<size_t>(g) * kT + tid] = role == 0 ? __fadd_rn(acc, __ldg(W.ntn_b + tid)) : acc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreadsUmma, 1)
sgpr_score_matrix_umma_kernel(const ScoreMatArgs A, const HeadParams H) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * kStageBytes);
    uint64_t* full = bars;            // [2] bulk copies -> MMA   (one arrive.expect_tx + the copies' byte count)
    uint64_t* empty = bars + 2;       // [2] MMA done  -> producers
    uint64_t* tfull = bars + 4;       // [2] MMA done  -> epilogue
    uint64_t* tempty = bars + 6;      // [2] epilogue  -> MMA     (count = epilogue threads)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * A.n_tiles / gridDim.x);
    const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * A.n_tiles / gridDim.x);

    if (tid == kThreadsUmma - 32) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
            mbar_init(tfull + s, 1);
            mbar_init(tempty + s, kEpiWarps * 32);
        }
        fence_mbar_init();
    }
    if (warp == 0) {       // one warp allocates all 512 TMEM columns (two accumulators) for the CTA
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kEpiWarps) {
        // ===================== epilogue =====================
        const int q = warp & 3, slice = warp >> 2;
        float cb[kT];
#pragma unroll
        for (int t = 0; t < kT; ++t) cb[t] = 0.0f;
        int cur_jb = -1;
        for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
            const int a = it & 1, aph = (it >> 1) & 1;
            const int jb = t / A.n_ib, ib = t - jb * A.n_ib;
            const int j = jb * kTileJ + q * 32 + lane;
            if (jb != cur_jb) {
                cur_jb = jb;
                const int js = min(j, A.M - 1);
#pragma unroll
                for (int t4 = 0; t4 < kT / 4; ++t4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(A.colblk + static_cast<size_t>(js) * kT) + t4);
                    cb[4 * t4] = v.x; cb[4 * t4 + 1] = v.y; cb[4 * t4 + 2] = v.z; cb[4 * t4 + 3] = v.w;
                }
            }
            mbar_wait_wd(tfull + a, aph);
            tc_fence_after();
            const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                  static_cast<uint32_t>(a * kUmmaN + slice * kRowsPerWarp * kT);
            float cur[kT], nxt[kT];
            tc_ld16(tcol, cur);
            tc_wait_ld();
#pragma unroll 1
            for (int u8 = 0; u8 < kRowsPerWarp; ++u8) {
                const int i = ib * kTileI + slice * kRowsPerWarp + u8;
                if (u8 < kRowsPerWarp - 1) tc_ld16(tcol + (u8 + 1) * kT, nxt);         // next row graph's 16 neurons in flight
                if (i < A.R) {
                    float z[kT];
#pragma unroll
                    for (int t4 = 0; t4 < kT / 4; ++t4) {
                        const float4 rb = __ldg(reinterpret_cast<const float4*>(A.rowblk + static_cast<size_t>(i) * kT) + t4);
                        z[4 * t4 + 0] = fmaxf(__fadd_rn(cur[4 * t4 + 0], __fadd_rn(rb.x, cb[4 * t4 + 0])), 0.0f);
                        z[4 * t4 + 1] = fmaxf(__fadd_rn(cur[4 * t4 + 1], __fadd_rn(rb.y, cb[4 * t4 + 1])), 0.0f);
                        z[4 * t4 + 2] = fmaxf(__fadd_rn(cur[4 * t4 + 2], __fadd_rn(rb.z, cb[4 * t4 + 2])), 0.0f);
                        z[4 * t4 + 3] = fmaxf(__fadd_rn(cur[4 * t4 + 3], __fadd_rn(rb.w, cb[4 * t4 + 3])), 0.0f);
                    }
                    float y = 0.0f;
#pragma unroll
                    for (int u = 0; u < kBn; ++u) {
                        float h = 0.0f;
#pragma unroll
                        for (int t2 = 0; t2 < kT; ++t2) h = fmaf(z[t2], H.fc1_w[u * kT + t2], h);
                        h = fmaxf(__fadd_rn(h, H.fc1_b[u]), 0.0f);
                        y = fmaf(h, H.fc2_w[u], y);
                    }
                    if (j < A.M) A.scores[static_cast<size_t>(i) * A.ld + j] = 1.0f / (1.0f + expf(-__fadd_rn(y, H.fc2_b)));
                }
                if (u8 < kRowsPerWarp - 1) {
                    tc_wait_ld();
#pragma unroll
                    for (int t2 = 0; t2 < kT; ++t2) cur[t2] = nxt[t2];
                }
            }
            tc_fence_before();
            mbar_arrive(tempty + a);
        }
    } else if (warp == kEpiWarps) {
        // ===================== operand loads: one thread, 1-D bulk TMA =====================
        if (lane == 0) {
            int held_jb[2] = {-1, -1};
            for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
                const int s = it & 1, ph = (it >> 1) & 1;
                const int jb = t / A.n_ib, ib = t - jb * A.n_ib;
                unsigned char* st = sm + s * kStageBytes;
                mbar_wait_wd(empty + s, ph ^ 1);             // the MMAs that read this stage two tiles ago have completed
                const bool need_a = held_jb[s] != jb;
                held_jb[s] = jb;
                mbar_expect_tx(full + s, 2 * kBBytes + (need_a ? 2 * kABytes : 0));
                if (need_a) {
                    const size_t o = static_cast<size_t>(jb) * kTileJ * 32;
                    bulk_g2s(st, A.cols_big + o, kABytes, full + s);
                    bulk_g2s(st + kABytes, A.cols_small + o, kABytes, full + s);
                }
                const size_t o = static_cast<size_t>(ib) * kUmmaN * 32;
                bulk_g2s(st + 2 * kABytes, A.proj_big + o, kBBytes, full + s);
                bulk_g2s(st + 2 * kABytes + kBBytes, A.proj_small + o, kBBytes, full + s);
            }
        }
    } else if (lane == 0) {
        // ===================== MMA issuer (one thread) =====================
        for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
            const int s = it & 1, ph = (it >> 1) & 1;
            mbar_wait_wd(tempty + s, ph ^ 1);                 // the epilogue has drained this accumulator
            mbar_wait_wd(full + s, ph);
            tc_fence_after();
            const uint32_t st = smem_u32(sm + s * kStageBytes);
            const uint64_t a_big = kmajor_sw128_desc(st), a_small = kmajor_sw128_desc(st + kABytes);
            const uint64_t b_big = kmajor_sw128_desc(st + 2 * kABytes), b_small = kmajor_sw128_desc(st + 2 * kABytes + kBBytes);
            const uint32_t d = tmem_base + static_cast<uint32_t>(s * kUmmaN);
#pragma unroll
            for (int k = 0; k < 4; ++k) {                  // K = 32 = 4 steps of 8 TF32 values (32 bytes -> +2 in the address field)
                tc_mma_tf32(d, a_small + 2 * k, b_big + 2 * k, kIdescTf32, k > 0 ? 1u : 0u);
                tc_mma_tf32(d, a_big + 2 * k, b_small + 2 * k, kIdescTf32, 1u);
                tc_mma_tf32(d, a_big + 2 * k, b_big + 2 * k, kIdescTf32, 1u);
            }
            tc_commit(empty + s);                          // arrives once every MMA above has completed
            tc_commit(tfull + s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}
#endif  // !SGPR_EMU

}  // namespace umma
}  // namespace sgpr
