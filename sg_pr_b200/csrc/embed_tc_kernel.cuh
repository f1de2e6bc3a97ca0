// Fused per-graph kernel with the 64-channel EdgeConv contractions on the 5th-generation tensor cores (N <= 64).
//
// Same path, same results contract as embed_kernel.cuh (its header names the reference code); what changes is where the
// two dense contractions of the four 64-channel layers (xyz 2-3, sem 2-3: sg_net.py:87-92, 97-102) run:
//     Gram   D_g[i][j]  = sum_c x_i[c] x_j[c]            (dgcnn.py:15: the distance rows are (2 D_g - xx_j) - xx_i)
//     GEMM   D_w[m][j]  = sum_c W[m][c] x_j[c]           (the per-node A|B = X [W_a | W_b] of the EdgeConv refactor)
// become tcgen05.mma kind::tf32 with the 3-term split x = big + small (tf32x3: no more k-NN divergences than fp32,
// profiles/r02_tf32x3_knn_experiment.json), fp32 accumulation in TMEM:
//   * the layer's INPUT lives in shared memory only as two operand planes (big, small; K-major, 128-byte swizzle,
//     tc_ops.cuh) written by the previous layer's gather epilogue — they are the A and the B operand of the Gram and the
//     B operand of the GEMM; there is no fp32 feature tile and no weight tile in shared memory any more;
//   * the layer's WEIGHTS are the TMEM-resident A operand of the GEMM (128 lanes = output channels A|B, 64 + 64 columns
//     big | small), stored with tcgen05.st from pre-split global arrays (pack.hpp::pack_edgeconv_tc) one layer ahead;
//   * accumulators: D_w in TMEM columns [128,192) (lane = channel, column = node), D_g in [192,256) (lane = node);
//     256 columns per CTA, so two CTAs still share an SM (shared memory ~100 KB per CTA as before).
// One thread issues the 48 MMAs of a layer right after the barrier that completes the planes; the Gram result comes back
// first (tcgen05.commit -> mbarrier), warps 0,1,4,5 (TMEM lane quadrants 0,1 = the 64 nodes) turn it into distance rows in
// shared memory, every warp runs the unchanged selection network on its rows while the GEMM MMAs finish, then all warps
// copy A|B out of TMEM (lane = channel -> conflict-free row stores) and the unchanged gather-max runs.
// xyz layer 1 (3 channels, direct form) and sem layer 1 (12 one-hot channels) keep their FFMA code.
//
// Barriers per tensor-core layer: A (planes complete) | C1 (distance rows complete) | C2 (selection done, rows may be
// overwritten) | B (A|B complete) — two more than the FFMA kernel, because the warp that can read a TMEM lane is not the
// warp that owns the row.
#pragma once
#include "embed_kernel.cuh"
#include "tc_ops.cuh"

namespace sgpr {

constexpr int kTcCols = 256;                 // TMEM columns per CTA
constexpr int kTcColW = 0, kTcColDw = 128, kTcColDg = 192;

struct TcSmem { int xb, xs, y, cat, in, ws, xx, red, bar, idx, cnt, total; };

__host__ __device__ inline TcSmem make_tc_layout(int ks) {
    TcSmem L;
    int o = 0;
    L.xb = o;  o += 2 * 64 * 128;               // big plane: two atoms of [64 rows][32 floats]
    L.xs = o;  o += 2 * 64 * 128;               // small plane
    L.y = o;   o += 64 * YS * 4;                // distance rows, then A|B   (also the overrun area of the M = 128 operand reads)
    L.cat = o; o += 64 * XS * 4;                // layer-0 coordinates, one-hot staging, cat(xyz3, sem3), node embeddings
    L.in = o;  o += kInCh * 64 * 4;
    L.ws = o;  o += 64 * 32 * 4;                // sem layer 1 matrix (6 KB), later conv_end (8 KB)
    L.xx = o;  o += 2 * 64 * 4;
    L.red = o; o += (kWarps * 32 + 64) * 4;
    L.bar = o; o += 64;
    L.idx = o; o += ((64 * ks * 2 + 15) / 16) * 16;
    L.cnt = o; o += 64;
    L.total = o + 1024;                         // + alignment slack: the planes must sit on a 1024-byte boundary
    return L;
}

// This thread's share of a layer's weight planes (lane = channel 32*(warp&3) + lane, plane = warp >> 2: 64 values): loaded
// from L2 into registers EARLY (the loads fly while the thread waits for the GEMM and copies A|B out), stored into TMEM
// once the GEMM that reads the current weights has completed.
struct TcWeights { float4 v[16]; };
__device__ __forceinline__ void tc_load_weights(TcWeights& w, const float* __restrict__ wtc, int warp, int lane) {
    const float4* src = reinterpret_cast<const float4*>(wtc + (warp >> 2) * 128 * 64 + (32 * (warp & 3) + lane) * 64);
#pragma unroll
    for (int q = 0; q < 16; ++q) w.v[q] = __ldg(src + q);
}
__device__ __forceinline__ void tc_store_weights(const TcWeights& w, uint32_t tmem_base, int warp) {
    const uint32_t dst = tmem_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + kTcColW + (warp >> 2) * 64;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[4 * q] = w.v[4 * c + q].x; v[4 * q + 1] = w.v[4 * c + q].y; v[4 * q + 2] = w.v[4 * c + q].z; v[4 * q + 3] = w.v[4 * c + q].w;
        }
        tc::st16(dst + 16 * c, v);
    }
    tc::wait_st();
}

// The MMAs of one layer, fully unrolled with precomputed operand handles (an MMA is issued by ONE thread and every
// instruction of a lone thread costs a pipeline latency: the issue loop has to be as short as the hardware allows).
// Two threads issue concurrently, one per accumulator: Gram (its result is needed first) and GEMM.
__device__ __forceinline__ void tc_issue_gram(tc::Operand xb, tc::Operand xs, uint32_t tmem_base, uint64_t* barG) {
    tc::fence_after();
    const uint32_t dg = tmem_base + kTcColDg;
#pragma unroll
    for (int t = 0; t < 3; ++t) {                    // small.big, big.small, big.big
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
            tc::mma_ss(dg, tc::advance((t == 0) ? xs : xb, ks), tc::advance((t == 1) ? xs : xb, ks), 64, (t | ks) != 0);
    }
    tc::commit(barG);
}
__device__ __forceinline__ void tc_issue_gemm(tc::Operand xb, tc::Operand xs, uint32_t tmem_base, uint64_t* barM) {
    tc::fence_after();
    const uint32_t dw = tmem_base + kTcColDw, aw = tmem_base + kTcColW;
#pragma unroll
    for (int t = 0; t < 3; ++t) {                    // W_small.X_big, W_big.X_small, W_big.X_big
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
            tc::mma_ts(dw, aw + ((t == 0) ? 64 : 0) + 8 * ks, tc::advance((t == 1) ? xs : xb, ks), 64, (t | ks) != 0);
    }
    tc::commit(barM);
}

// Gram accumulator -> distance rows in shared memory (gram_rows' output contract: columns R <= c < N carry the pad class's
// value pd[i][R-1], columns >= N carry -inf).  Thread = node 32*(warp&3) + lane (warps of quadrants 0, 1), half = warp >> 2.
__device__ __forceinline__ void tc_copy_pd(uint32_t tmem_base, const float* __restrict__ sXX, float* __restrict__ sY, int R, int N,
                                           int warp, int lane) {
    const int row = 32 * (warp & 3) + lane, half = warp >> 2;
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + kTcColDg;
    float d[32], pc[16];
    tc::ld16(t0 + 32 * half, *reinterpret_cast<float(*)[16]>(d));
    tc::ld16(t0 + 32 * half + 16, *reinterpret_cast<float(*)[16]>(d + 16));
    tc::ld16(t0 + (((R - 1) >> 4) << 4), pc);                        // the 16-column chunk that holds column R-1
    tc::wait_ld();
    if (row >= R) return;
    const float xxr = sXX[row];
    float dpad = pc[0];
#pragma unroll
    for (int e = 1; e < 16; ++e) dpad = (((R - 1) & 15) == e) ? pc[e] : dpad;
    const float padv = __fsub_rn(__fsub_rn(__fmul_rn(2.0f, dpad), sXX[R - 1]), xxr);
    float* dst = sY + row * YS + 32 * half;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 32 * half + 4 * c4 + e;
            const float real = __fsub_rn(__fsub_rn(__fmul_rn(2.0f, d[4 * c4 + e]), sXX[min(c, 63)]), xxr);
            o[e] = (c < R) ? real : ((c < N) ? padv : -INFINITY);
        }
        *reinterpret_cast<float4*>(dst + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// GEMM accumulator (lane = channel, column = node) -> A|B rows in shared memory.  Thread = channel 32*(warp&3) + lane.
__device__ __forceinline__ void tc_copy_ab(uint32_t tmem_base, float* __restrict__ sY, int R, int cout2, int warp, int lane) {
    const int ch = 32 * (warp & 3) + lane, half = warp >> 2;
    if (32 * (warp & 3) >= cout2) return;                             // 32-channel layers: quadrants 2, 3 hold nothing
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + kTcColDw + 32 * half;
    float v[32];
    tc::ld16(t0, *reinterpret_cast<float(*)[16]>(v));
    tc::ld16(t0 + 16, *reinterpret_cast<float(*)[16]>(v + 16));
    tc::wait_ld();
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        const int n = 32 * half + e;
        if (n < R) sY[n * YS + ch] = v[e];
    }
}

__global__ void __launch_bounds__(kThreads, 2)
sgpr_embed_tc_kernel(const EmbedArgs A, const PackedWeights W, const HeadParams H) {
    SGPR_DYN_SMEM(smem_raw);
#ifdef SGPR_EMU
    unsigned char* smem = smem_raw + ((1024u - (reinterpret_cast<uintptr_t>(smem_raw) & 1023u)) & 1023u);
#else
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#endif
    const TcSmem L = make_tc_layout(A.KS);
    float* sXb = reinterpret_cast<float*>(smem + L.xb);
    float* sXs = reinterpret_cast<float*>(smem + L.xs);
    float* sY = reinterpret_cast<float*>(smem + L.y);
    float* sCat = reinterpret_cast<float*>(smem + L.cat);
    float* sIn = reinterpret_cast<float*>(smem + L.in);
    float* sWs = reinterpret_cast<float*>(smem + L.ws);
    float* sXX = reinterpret_cast<float*>(smem + L.xx);
    float* sXX0 = sXX + 64;
    float* sRed = reinterpret_cast<float*>(smem + L.red);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar);
    uint16_t* sIdx = reinterpret_cast<uint16_t*>(smem + L.idx);
    uint8_t* sCnt = smem + L.cnt;
    __shared__ int sFlag;
    __shared__ int sLast[kWarps];
    __shared__ int sSlot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = A.N, k = A.k, KS = A.KS;
    uint64_t* barIn = bars;          // input record landed
    uint64_t* barWs = bars + 1;      // small weight tile landed
    uint64_t* barG = bars + 2;       // Gram MMAs complete
    uint64_t* barM = bars + 3;       // GEMM MMAs complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    uint32_t phIn = 0, phWs = 0, phG = 0, phM = 0;
    const PlaneOut planes{sXb, sXs, sXX};

    if (tid == 0) { mbar_init(barIn, 1); mbar_init(barWs, 1); mbar_init(barG, 1); mbar_init(barM, 1); fence_mbar_init(); }
    if (warp == 0) tc::alloc(tmem_slot, kTcCols);
    for (int e = tid; e < 64 * XS; e += kThreads) sCat[e] = 0.0f;
    for (int e = tid; e < 2 * 64; e += kThreads) sXX[e] = 0.0f;
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = *tmem_slot;
    {   // weights of the first tensor-core layer
        TcWeights w;
        tc_load_weights(w, W.wtc_s2, warp, lane);
        tc_store_weights(w, tmem_base, warp);
    }
    tc::fence_before();
    const tc::Operand opXb = tc::operand(sXb), opXs = tc::operand(sXs);

    const uint32_t inBytes = A.compact ? static_cast<uint32_t>(A.compact) : static_cast<uint32_t>(kInCh * N * 4);

#pragma unroll 1
    for (int it = blockIdx.x;; it += gridDim.x) {
        int slot = it;
        if (A.work_ctr) {
            if (tid == 0) sSlot = atomicAdd(A.work_ctr, 1);
            __syncthreads();
            slot = sSlot;
            __syncthreads();
        }
        if (slot >= A.G) break;
        const int g = A.order ? __ldg(A.order + slot) : slot;
        const float* gin = reinterpret_cast<const float*>(
            reinterpret_cast<const unsigned char*>((A.pairs && (g & 1)) ? A.g1 : A.g0) + static_cast<size_t>(A.pairs ? (g >> 1) : g) * inBytes);
        const bool bulk_ok = ((inBytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(gin) & 15u) == 0);

        // ---- stage the input record and the sem layer 1 matrix (bulk TMA) ----
        if (tid == 0) {
            if (bulk_ok) { mbar_expect_tx(barIn, inBytes); bulk_g2s(sIn, gin, inBytes, barIn); }
            mbar_expect_tx(barWs, 12 * 128 * 4);
            bulk_g2s(sWs, W.w_f1, 12 * 128 * 4, barWs);
        }
        if (bulk_ok) { mbar_wait(barIn, phIn); phIn ^= 1; }
        else { for (int e = tid; e < static_cast<int>(inBytes / 4); e += kThreads) sIn[e] = __ldg(gin + e); __syncthreads(); }
        if (A.compact) {
            const int lab = (tid < N) ? reinterpret_cast<const uint8_t*>(sIn)[12 * N + tid] : 255;
            __syncthreads();
            if (tid < N) {
#pragma unroll
                for (int c = 0; c < kLabels; ++c) sIn[(3 + c) * N + tid] = (lab == c) ? 1.0f : 0.0f;
            }
            __syncthreads();
        }

        // ---- layer-0 tile (x, y, z, 0) + squared norms, and the last non-zero node ----
        int last = -1;
        for (int n = tid; n < N; n += kThreads) {
            const float x = sIn[n], y = sIn[N + n], z = sIn[2 * N + n];
            *reinterpret_cast<float4*>(sCat + n * XS) = make_float4(x, y, z, 0.0f);
            sXX0[n] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
            uint32_t bits = 0;
#pragma unroll
            for (int c = 0; c < kInCh; ++c) bits |= __float_as_uint(sIn[c * N + n]);
            if ((bits << 1) != 0u) last = n;
        }
        last = __reduce_max_sync(0xffffffffu, last);
        if (lane == 0) sLast[warp] = last;
        __syncthreads();
        int R = N;
        if (A.dedup) {
            int m = -1;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) m = max(m, sLast[w]);
            R = min(m + 2, N);
        }
        const int rpw = (R + kWarps - 1) / kWarps;
        const int w0 = min(R, warp * rpw), w1 = min(R, w0 + rpw);

        // ================= the six EdgeConv layers: xyz 1,2,3 then sem 1,2,3 (sg_net.py:84-102) =================
#pragma unroll 1
        for (int l = 0; l < 6; ++l) {
            const bool tcl = (l != 0 && l != 3);
            const float* ab = (l == 1) ? W.ab_s2 : (l == 2) ? W.ab_s3 : (l == 3) ? W.ab_f1 : (l == 4) ? W.ab_f2 : W.ab_f3;
            const int cout = (l == 2 || l == 5) ? 32 : 64;
            if (tcl) {
                // ---- contractions on the tensor cores ----
                if (tid == 0) tc_issue_gram(opXb, opXs, tmem_base, barG);
                if (tid == 32) tc_issue_gemm(opXb, opXs, tmem_base, barM);
                mbar_wait(barG, phG); phG ^= 1;
                tc::fence_after();
                if ((warp & 3) < 2) tc_copy_pd(tmem_base, sXX, sY, R, N, warp, lane);
                tc::fence_before();
                __syncthreads();                                   // barrier C1: every distance row is in shared memory
                for (int r0 = w0; r0 < w1; r0 += 8)
                    select_rows<2, 0>(sY, sIdx, sCnt, nullptr, R, N, k, KS, YS, r0, min(8, w1 - r0), lane);
                __syncthreads();                                   // barrier C2: the distance rows are dead
                // weights of the next tensor-core layer (after the last one: of the first, for this CTA's next graph)
                TcWeights wnext;
                tc_load_weights(wnext, (l == 1) ? W.wtc_s3 : (l == 2) ? W.wtc_f2 : (l == 4) ? W.wtc_f3 : W.wtc_s2, warp, lane);
                mbar_wait(barM, phM); phM ^= 1;
                tc::fence_after();
                tc_copy_ab(tmem_base, sY, R, 2 * cout, warp, lane);
                tc_store_weights(wnext, tmem_base, warp);          // the GEMM has completed: its weight columns are free
                tc::fence_before();
            } else {
                // ---- xyz layer 1 / sem layer 1: FFMA front (distance rows -> selection -> GEMM rows), own rows ----
                const float* sXt = (l == 0) ? sCat : sCat + 32;
                const float* sXXl = (l == 0) ? sXX0 : sXX;
                const int c4n = (l == 0) ? 1 : 3;
                if (l == 3) { mbar_wait(barWs, phWs); phWs ^= 1; }
                for (int r0 = w0; r0 < w1; r0 += 8) {
                    const int nr = min(8, w1 - r0);
                    if (R <= 32) { SGPR_NR_SWITCH(nr, (gram_rows<2, 1, NR>(sXt, sXXl, sY, c4n, R, N, r0, lane))) }
                    else         { SGPR_NR_SWITCH(nr, (gram_rows<2, 2, NR>(sXt, sXXl, sY, c4n, R, N, r0, lane))) }
                    __syncwarp();
                    select_rows<2, 0>(sY, sIdx, sCnt, nullptr, R, N, k, KS, (l == 0) ? XS : YS, r0, nr, lane);
                    __syncwarp();
                    if (l == 3) { SGPR_NR_SWITCH(nr, (gemm_rows<NR, 4, 0>(sXt, sWs, sY, YS, nullptr, c4n, r0, lane))) }
                }
            }
            if (l == 0) {
                xyz_rows<1>(sCat, sIdx, sCnt, KS, W.s1, nullptr, nullptr, w0, w1, lane, &planes);
            } else {
                __syncthreads();                                   // barrier B: every A|B row is in place
                if (l == 3 && tid == 0) { mbar_expect_tx(barWs, 64 * 32 * 4); bulk_g2s(sWs, W.w_end, 64 * 32 * 4, barWs); }
                if (cout == 64) {
                    gather_rows<64, 1>(sY, sIdx, sCnt, KS, ab, nullptr, nullptr, w0, w1, lane, &planes);
                } else {
                    gather_rows<32>(sY, sIdx, sCnt, KS, ab, (l == 2) ? sCat : sCat + 32, nullptr, w0, w1, lane);
                    if (l == 2) {   // stage the semantic branch input: one-hot rows at sCat[n][32..47] (sg_net.py:82,94)
                        for (int i = w0; i < w1; ++i)
                            if (lane < 16) sCat[i * XS + 32 + lane] = (lane < kLabels) ? sIn[(3 + lane) * N + i] : 0.0f;
                        __syncwarp();
                        norms_rows(sCat + 32, sXX, 3, w0, w1, lane);
                    } else {        // l == 5: conv_end on own rows, in place (sg_net.py:104-109): [.,64] -> [.,32]
                        __syncwarp();
                        mbar_wait(barWs, phWs); phWs ^= 1;
                        for (int r0 = w0; r0 < w1; r0 += 8)
                            conv_end_dispatch(sCat, sWs, sCat, W.ab_end, r0, min(8, w1 - r0), lane);
                    }
                }
            }
            tc::fence_proxy_async();                               // operand planes: generic-proxy stores -> tensor core
            __syncthreads();                                       // barrier A: the next layer's input is complete
        }

        finish_graph(A, W, H, sCat, sRed, sY, &sFlag, g, N, R, tid, warp, lane);
        __syncthreads();
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) { tc::fence_after(); tc::dealloc(tmem_base, kTcCols); }
}

}  // namespace sgpr
