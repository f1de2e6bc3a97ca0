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
constexpr int kEpiWarps = SGPR_UMMA_EPI_WARPS;                       // 4, 8 or 16 (kRowsPerWarp stays even)
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
    const float* rowblk;    // [R][16]      first half of V [e1;e2] + tensor_network.bias (the second half rides in proj)
    float* scores;          // [R][ld]
    long long ld;
    int R, M;
    int n_ib;               // row groups = ceil(R / 16)
    int n_tiles;            // column blocks * row groups
    // Fused exchange (multi-GPU scan): every score is stored into n_out destination matrices — this GPU's and its peers'
    // [M][M] results, mapped over NVLink (CUDA IPC + peer access) — instead of one local store followed by an all-gather.
    // n_out == 0: store to `scores` only.
    int n_out;
    float* outs[8];
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
//   role 0 (rows, side 1): 16 operand rows g*16+t with  P[t][b] = sum_a e_g[a] W[a][b][t] + V[t][32+b]  (layers_batch.py:78
//                          and the column half of :80-81), and blk[g][t] = sum_c V[t][c] e_g[c] + bias[t]  (:80-82)
//   role 1 (cols, side 2): 1 operand row g = e_g
// Element b of operand row r lives at float offset r*32 + (((b>>2) ^ (r&7)) << 2) + (b&3): the 128-byte swizzle, which
// depends only on r mod 8 and therefore survives the copy of whole 8-row-aligned tiles.
constexpr int kSplitSmem = ((kF3 + 1) * 512 + 8 * kF3) * 4;      // transposed NTN tensor + V2 + 8 pooled vectors

__global__ void __launch_bounds__(kThreads)
sgpr_ntn_split_kernel(const float* __restrict__ pooled_r, int count_r, int padded_r, float* __restrict__ big_r,
                      float* __restrict__ small_r, float* __restrict__ blk_r, const float* __restrict__ pooled_c, int count_c,
                      int padded_c, float* __restrict__ big_c, float* __restrict__ small_c, float* __restrict__ blk_c,
                      int row_ctas, const PackedWeights W) {
    extern __shared__ float split_smem[];
    float* sWt = split_smem;                 // role 0: [a][t*32 + b] = W[a][b*16 + t]  (conflict-free reads, coalesced writes)
    float* sE = split_smem + (kF3 + 1) * 512;  // [8][32] pooled vectors of the graphs in flight
    const int tid = threadIdx.x;
    const int role = static_cast<int>(blockIdx.x) >= row_ctas ? 1 : 0;
    const int cta = role ? static_cast<int>(blockIdx.x) - row_ctas : static_cast<int>(blockIdx.x);
    const int ctas = role ? static_cast<int>(gridDim.x) - row_ctas : row_ctas;
    const float* pooled = role ? pooled_c : pooled_r;
    const int count = role ? count_c : count_r, padded = role ? padded_c : padded_r;
    float* big = role ? big_c : big_r;
    float* small = role ? small_c : small_r;
    float* blk = role ? blk_c : blk_r;
    constexpr int GB = 8;                    // graphs per step
    if (role == 0) {
        // + the column half of the V-block: sum_b (W_a[b][t] e_i[a] + V[t][32+b]) e_j[b] = bilinear + V2 e_j, so the
        // per-column term of layers_batch.py:80-81 rides through the MMA (row a = 32 of the table, weight 1)
        for (int e = tid; e < kF3 * 512; e += kThreads) {
            const int a = e >> 9, c = e & 511;                   // c = t*32 + b as used; stored as b*16 + t
            sWt[e] = __ldg(W.ntn_w + a * 512 + (c & 31) * kT + (c >> 5));
        }
        for (int c = tid; c < 512; c += kThreads) sWt[kF3 * 512 + c] = __ldg(W.ntn_v + (c >> 5) * 64 + 32 + (c & 31));
    }
    for (int g0 = cta * GB; g0 < padded; g0 += ctas * GB) {
        __syncthreads();
        {
            const int g = g0 + (tid >> 5);
            sE[tid] = (g < count) ? __ldg(pooled + static_cast<size_t>(g) * kF3 + (tid & 31)) : 0.0f;
        }
        __syncthreads();
        if (role == 0) {
            for (int c = tid; c < 512; c += kThreads) {           // c = t*32 + b
                float x[GB];
                const float v2 = sWt[kF3 * 512 + c];
#pragma unroll
                for (int u = 0; u < GB; ++u) x[u] = (g0 + u < count) ? v2 : 0.0f;
#pragma unroll 8
                for (int a = 0; a < kF3; ++a) {
                    const float w = sWt[a * 512 + c];
#pragma unroll
                    for (int u = 0; u < GB; ++u) x[u] = fmaf(sE[u * kF3 + a], w, x[u]);
                }
                const int t = c >> 5, b = c & 31;
#pragma unroll
                for (int u = 0; u < GB; ++u) {
                    if (g0 + u >= padded) break;
                    const size_t r = static_cast<size_t>(g0 + u) * kT + t;
                    const size_t off = r * 32 + ((((b >> 2) ^ static_cast<int>(r & 7)) << 2) + (b & 3));
                    const float hi = tf32_rna(x[u]);
                    big[off] = hi;
                    small[off] = x[u] - hi;
                }
            }
        } else {
            const int u = tid >> 5, b = tid & 31;
            if (g0 + u < padded) {
                const size_t r = static_cast<size_t>(g0 + u);
                const size_t off = r * 32 + ((((b >> 2) ^ static_cast<int>(r & 7)) << 2) + (b & 3));
                const float x = sE[tid], hi = tf32_rna(x);
                big[off] = hi;
                small[off] = x - hi;
            }
        }
        if (role == 0 && tid < GB * kT) {                          // row half of the V-block + bias
            const int u = tid >> 4, t = tid & 15;
            if (g0 + u < count) {
                float acc = 0.0f;
                for (int c = 0; c < kF3; ++c) acc = fmaf(__ldg(W.ntn_v + t * 64 + c), sE[u * kF3 + c], acc);
                blk[static_cast<size_t>(g0 + u) * kT + t] = __fadd_rn(acc, __ldg(W.ntn_b + t));
            }
        }
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
        for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
            const int a = it & 1, aph = (it >> 1) & 1;
            const int jb = t / A.n_ib, ib = t - jb * A.n_ib;
            const int j = jb * kTileJ + q * 32 + lane;
            mbar_wait_wd(tfull + a, aph);
            tc_fence_after();
            const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                  static_cast<uint32_t>(a * kUmmaN + slice * kRowsPerWarp * kT);
            // two row graphs per step: their 2 x 16 neurons ride in the two halves of packed fp32x2 registers, so every FC1
            // weight (a uniform operand) feeds one FFMA2 = two pairs
#pragma unroll 1
            for (int u8 = 0; u8 < kRowsPerWarp; u8 += 2) {
                const int i = ib * kTileI + slice * kRowsPerWarp + u8;
                float c0[kT], c1[kT];
                tc_ld16(tcol + u8 * kT, c0);
                tc_ld16(tcol + (u8 + 1) * kT, c1);
                tc_wait_ld();                                        // (the other 15 warps cover this latency)
                if (i < A.R) {
                    const int i1 = min(i + 1, A.R - 1);
                    float2 z[kT];
#pragma unroll
                    for (int t4 = 0; t4 < kT / 4; ++t4) {
                        const float4 ra = __ldg(reinterpret_cast<const float4*>(A.rowblk + static_cast<size_t>(i) * kT) + t4);
                        const float4 rb = __ldg(reinterpret_cast<const float4*>(A.rowblk + static_cast<size_t>(i1) * kT) + t4);
                        const float a4[4] = {ra.x, ra.y, ra.z, ra.w}, b4[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int tt = 4 * t4 + c;
                            z[tt].x = fmaxf(__fadd_rn(c0[tt], a4[c]), 0.0f);
                            z[tt].y = fmaxf(__fadd_rn(c1[tt], b4[c]), 0.0f);
                        }
                    }
                    float2 y = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int u = 0; u < kBn; ++u) {
                        float2 h = make_float2(0.0f, 0.0f);
#pragma unroll
                        for (int t2 = 0; t2 < kT; ++t2) {
                            const float w = H.fc1_w[u * kT + t2];
                            h = __ffma2_rn(z[t2], make_float2(w, w), h);
                        }
                        h.x = fmaxf(__fadd_rn(h.x, H.fc1_b[u]), 0.0f);
                        h.y = fmaxf(__fadd_rn(h.y, H.fc1_b[u]), 0.0f);
                        y = __ffma2_rn(h, make_float2(H.fc2_w[u], H.fc2_w[u]), y);
                    }
                    if (j < A.M) {
                        const float s0 = 1.0f / (1.0f + expf(-__fadd_rn(y.x, H.fc2_b)));
                        const float s1 = 1.0f / (1.0f + expf(-__fadd_rn(y.y, H.fc2_b)));
                        const size_t o = static_cast<size_t>(i) * A.ld + j;
                        if (A.n_out == 0) {
                            A.scores[o] = s0;
                            if (i + 1 < A.R) A.scores[o + A.ld] = s1;
                        } else {
                            for (int p = 0; p < A.n_out; ++p) {        // posted stores: local HBM and NVLink peers alike
                                A.outs[p][o] = s0;
                                if (i + 1 < A.R) A.outs[p][o + A.ld] = s1;
                            }
                        }
                    }
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

// =====================================================================================================================
// Version 2: FC1 on the tensor cores as well.
//
// After the bilinear MMA the epilogue of version 1 still spends 256 fp32 FMAs per pair on  h = W1 z  (sg_net.py:134) — the
// FMA-pipe floor of the whole kernel.  Here the epilogue thread only forms z (16 values), splits it into TF32 big / small
// and writes both back into TENSOR MEMORY (tcgen05.st); a second, small UMMA per row graph takes its A operand straight
// from TMEM (the ".ts" form: D2[j][u] = sum_t Z[j][t] W1[u][t], M = 128, N = 16, K = 16, 3xTF32 = 6 instructions) and
// leaves h in the accumulator columns the bilinear result occupied.  The thread then reads 16 h values back and finishes
// with bias, relu, FC2 and the sigmoid: ~160 instructions per pair instead of ~400.
//
// TMEM map (512 columns): [0,128) and [128,256) the two D1 accumulators of an 8-row-graph tile (N = 128), [256,512) the Z
// operands of the tile in flight: row graph il owns columns 256 + 32 il .. +31 = [ Z_big(16) | Z_small(16) ].
//   MMA2 pass 1: A = [Z_big | Z_small] (K = 32), B = rows [W1_big | W1_big]   ->  Z_big.W1_big + Z_small.W1_big
//   MMA2 pass 2: A = Z_big (K = 16),            B = rows [W1_small | 0]      ->  + Z_big.W1_small
// Warps 0-7 only PRODUCE Z (per slice of four row graphs the four quadrant warps arrive on zfull[slice], 128 arrivals),
// warps 8-15 only CONSUME h (hfull[buffer][slice], tcgen05.commit of the slice's FC1 MMAs) — neither half ever idles through the MMA round trip of
// its own tile; the MMA thread issues the next tile's bilinear MMAs before it serves the current tile's FC1 MMAs.
// =====================================================================================================================
constexpr int kTileI2 = 8;                       // row graphs per tile
constexpr int kN1 = kTileI2 * kT;                // 128 accumulator columns per D1 buffer
constexpr int kEpiWarps2 = 16;
constexpr int kThreadsUmma2 = (kEpiWarps2 + 2) * 32;
constexpr int kB2Bytes = kN1 * 128;              // one plane of the B tile (128 rows x 32 floats)
constexpr int kStage2Bytes = 2 * kABytes + 2 * kB2Bytes;          // 64 KB
constexpr int kWPlaneBytes = kT * 128;           // 16 rows x 32 floats
constexpr int kSmemUmma2 = 2 * kStage2Bytes + 2 * kWPlaneBytes + 1024 + 256;
constexpr uint32_t kIdesc1 = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kN1 >> 3) << 17) |
                             (static_cast<uint32_t>(kTileJ >> 4) << 24);
constexpr uint32_t kIdesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kT >> 3) << 17) |
                             (static_cast<uint32_t>(kTileJ >> 4) << 24);

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float (&a)[16], const float (&b)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(a[4])), "r"(__float_as_uint(a[5])), "r"(__float_as_uint(a[6])), "r"(__float_as_uint(a[7])),
                   "r"(__float_as_uint(a[8])), "r"(__float_as_uint(a[9])), "r"(__float_as_uint(a[10])), "r"(__float_as_uint(a[11])),
                   "r"(__float_as_uint(a[12])), "r"(__float_as_uint(a[13])), "r"(__float_as_uint(a[14])), "r"(__float_as_uint(a[15])),
                   "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])), "r"(__float_as_uint(b[2])), "r"(__float_as_uint(b[3])),
                   "r"(__float_as_uint(b[4])), "r"(__float_as_uint(b[5])), "r"(__float_as_uint(b[6])), "r"(__float_as_uint(b[7])),
                   "r"(__float_as_uint(b[8])), "r"(__float_as_uint(b[9])), "r"(__float_as_uint(b[10])), "r"(__float_as_uint(b[11])),
                   "r"(__float_as_uint(b[12])), "r"(__float_as_uint(b[13])), "r"(__float_as_uint(b[14])), "r"(__float_as_uint(b[15]))
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct ScoreMat2Args {
    const float* cols_big;     // as ScoreMatArgs; proj planes padded to whole 8-row-graph tiles
    const float* cols_small;
    const float* proj_big;
    const float* proj_small;
    const float* rowblk;       // [R][16]  V-block first half + tensor_network.bias
    const float* fc1_planes;   // [2][16][32] swizzled: rows [W1_big | W1_big], rows [W1_small | 0]
    float* scores;
    long long ld;
    int R, M, n_ib, n_tiles;
};

__global__ void __launch_bounds__(kThreadsUmma2, 1)
sgpr_score_matrix_umma2_kernel(const ScoreMat2Args A, const HeadParams H) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* sWp = sm + 2 * kStage2Bytes;                       // W1 planes (1024-aligned: 128 KB offset)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sWp + 2 * kWPlaneBytes);
    uint64_t* full = bars;            // [2]
    uint64_t* empty = bars + 2;       // [2]
    uint64_t* tfull = bars + 4;       // [2]
    uint64_t* tempty = bars + 6;      // [2]
    uint64_t* zfull = bars + 8;       // [2]     Z of a slice (four row graphs) in TMEM (128 arrivals: its four quadrant warps)
    uint64_t* hfull = bars + 16;      // [2][2]  FC1 accumulators of a slice of the tile in D1 buffer a complete
    uint64_t* wfull = bars + 32;      // W1 planes landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 33);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * A.n_tiles / gridDim.x);
    const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * A.n_tiles / gridDim.x);

    if (tid == kThreadsUmma2 - 32) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
            mbar_init(tfull + s, 1);
            mbar_init(tempty + s, (kEpiWarps2 / 2) * 32);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(zfull + i, 128); mbar_init(hfull + i, 1); mbar_init(hfull + 2 + i, 1); }
        mbar_init(wfull, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kEpiWarps2 / 2) {
        // ===================== epilogue, first half: s -> z = relu(s + V-block + bias) -> TF32 big/small -> TMEM =========
        // (these warps never wait for the FC1 MMA of the tile they just fed: they move on to the next tile's s)
        const int q = warp & 3, slice = warp >> 2;                 // slice 0..1 -> row graphs 4*slice .. 4*slice + 3 of the tile
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
            const int a = it & 1, aph = (it >> 1) & 1;
            const int jb = t / A.n_ib, ib = t - jb * A.n_ib;
            mbar_wait_wd(tfull + a, aph);
            tc_fence_after();
            const uint32_t d1 = tmem_base + lane_base + static_cast<uint32_t>(a * kN1);
#pragma unroll 1
            for (int u = 0; u < kTileI2 / 2; ++u) {
                const int il = (kTileI2 / 2) * slice + u;
                const int i = min(ib * kTileI2 + il, A.R - 1);      // rows beyond R: computed, never stored
                float sv[kT], zb[kT], zs[kT];
                tc_ld16(d1 + il * kT, sv);
                float4 rb[kT / 4];
#pragma unroll
                for (int t4 = 0; t4 < kT / 4; ++t4) rb[t4] = __ldg(reinterpret_cast<const float4*>(A.rowblk + static_cast<size_t>(i) * kT) + t4);
                tc_wait_ld();
#pragma unroll
                for (int t4 = 0; t4 < kT / 4; ++t4) {
                    const float r4[4] = {rb[t4].x, rb[t4].y, rb[t4].z, rb[t4].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float z = fmaxf(__fadd_rn(sv[4 * t4 + c], r4[c]), 0.0f);
                        zb[4 * t4 + c] = tf32_rna(z);
                        zs[4 * t4 + c] = __fsub_rn(z, zb[4 * t4 + c]);
                    }
                }
                // Z slots of this slice: free once the FC1 MMAs of the previous tile have read them
                if (u == 0 && it > 0) { mbar_wait_wd(hfull + ((it - 1) & 1) * 2 + slice, ((it - 1) >> 1) & 1); tc_fence_after(); }
                tc_st32(tmem_base + lane_base + static_cast<uint32_t>(2 * kN1 + il * 32), zb, zs);
            }
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(zfull + slice);                              // one hand-over per slice of four row graphs
        }
    } else if (warp < kEpiWarps2) {
        // ===================== epilogue, second half: h <- TMEM, bias, relu, FC2, sigmoid, store ========================
        const int w2 = warp - kEpiWarps2 / 2;
        const int q = w2 & 3, slice = w2 >> 2;                      // (warp & 3 == w2 & 3: the TMEM lane quadrant of this warp)
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
            const int a = it & 1, aph = (it >> 1) & 1;
            const int jb = t / A.n_ib, ib = t - jb * A.n_ib;
            const int j = jb * kTileJ + q * 32 + lane;
            const uint32_t d1 = tmem_base + lane_base + static_cast<uint32_t>(a * kN1);
#pragma unroll 1
            for (int u = 0; u < kTileI2 / 2; ++u) {
                const int il = (kTileI2 / 2) * slice + u;
                const int i = ib * kTileI2 + il;
                float h[kT];
                if (u == 0) { mbar_wait_wd(hfull + a * 2 + slice, aph); tc_fence_after(); }
                tc_ld16(d1 + il * kT, h);
                tc_wait_ld();
                float y = 0.0f;
#pragma unroll
                for (int c = 0; c < kBn; ++c) y = fmaf(fmaxf(__fadd_rn(h[c], H.fc1_b[c]), 0.0f), H.fc2_w[c], y);
                if (i < A.R && j < A.M) A.scores[static_cast<size_t>(i) * A.ld + j] = 1.0f / (1.0f + expf(-__fadd_rn(y, H.fc2_b)));
            }
            tc_fence_before();
            mbar_arrive(tempty + a);
        }
    } else if (warp == kEpiWarps2) {
        // ===================== operand loads: one thread, 1-D bulk TMA =====================
        if (lane == 0) {
            mbar_expect_tx(wfull, 2 * kWPlaneBytes);
            bulk_g2s(sWp, A.fc1_planes, 2 * kWPlaneBytes, wfull);
            int held_jb[2] = {-1, -1};
            for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
                const int s = it & 1, ph = (it >> 1) & 1;
                const int jb = t / A.n_ib, ib = t - jb * A.n_ib;
                unsigned char* st = sm + s * kStage2Bytes;
                mbar_wait_wd(empty + s, ph ^ 1);
                const bool need_a = held_jb[s] != jb;
                held_jb[s] = jb;
                mbar_expect_tx(full + s, 2 * kB2Bytes + (need_a ? 2 * kABytes : 0));
                if (need_a) {
                    const size_t o = static_cast<size_t>(jb) * kTileJ * 32;
                    bulk_g2s(st, A.cols_big + o, kABytes, full + s);
                    bulk_g2s(st + kABytes, A.cols_small + o, kABytes, full + s);
                }
                const size_t o = static_cast<size_t>(ib) * kN1 * 32;
                bulk_g2s(st + 2 * kABytes, A.proj_big + o, kB2Bytes, full + s);
                bulk_g2s(st + 2 * kABytes + kB2Bytes, A.proj_small + o, kB2Bytes, full + s);
            }
        }
    } else if (lane == 0) {
        // ===================== MMA issuer (one thread) =====================
        auto issue_bilinear = [&](int it) {
            const int s = it & 1, ph = (it >> 1) & 1;
            mbar_wait_wd(tempty + s, ph ^ 1);                       // epilogue finished the tile that used this accumulator
            mbar_wait_wd(full + s, ph);
            tc_fence_after();
            const uint32_t st = smem_u32(sm + s * kStage2Bytes);
            const uint64_t a_big = kmajor_sw128_desc(st), a_small = kmajor_sw128_desc(st + kABytes);
            const uint64_t b_big = kmajor_sw128_desc(st + 2 * kABytes), b_small = kmajor_sw128_desc(st + 2 * kABytes + kB2Bytes);
            const uint32_t d = tmem_base + static_cast<uint32_t>(s * kN1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tc_mma_tf32(d, a_small + 2 * k, b_big + 2 * k, kIdesc1, k > 0 ? 1u : 0u);
                tc_mma_tf32(d, a_big + 2 * k, b_small + 2 * k, kIdesc1, 1u);
                tc_mma_tf32(d, a_big + 2 * k, b_big + 2 * k, kIdesc1, 1u);
            }
            tc_commit(empty + s);
            tc_commit(tfull + s);
        };
        const int n_it = t_end - t_begin;
        mbar_wait_wd(wfull, 0);
        const uint64_t wb = kmajor_sw128_desc(smem_u32(sWp)), ws = kmajor_sw128_desc(smem_u32(sWp + kWPlaneBytes));
        if (n_it > 0) issue_bilinear(0);
        for (int it = 0; it < n_it; ++it) {
            if (it + 1 < n_it) issue_bilinear(it + 1);              // next tile's bilinear MMAs run under this tile's epilogue
            const uint32_t d1 = tmem_base + static_cast<uint32_t>((it & 1) * kN1);
            for (int sl = 0; sl < 2; ++sl) {
                mbar_wait_wd(zfull + sl, it & 1);
                tc_fence_after();
                for (int il = sl * (kTileI2 / 2); il < (sl + 1) * (kTileI2 / 2); ++il) {
                    const uint32_t d2 = d1 + il * kT, az = tmem_base + static_cast<uint32_t>(2 * kN1 + il * 32);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_tf32_ts(d2, az + 8 * k, wb + 2 * k, kIdesc2, k > 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 2; ++k) tc_mma_tf32_ts(d2, az + 8 * k, ws + 2 * k, kIdesc2, 1u);
                }
                tc_commit(hfull + (it & 1) * 2 + sl);
            }
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
