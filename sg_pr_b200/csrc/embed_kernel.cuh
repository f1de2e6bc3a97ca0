// Fused per-graph kernel: 6 dynamic EdgeConv layers -> conv_end -> attention pooling (-> pair head).
//
// Reference path replaced (all fp32):
//   SG.dgcnn_conv_pass   /root/reference/sg_net.py:79-110
//   dgcnn.knn            /root/reference/dgcnn.py:14-20      (pd = -xx - inner - xx^T ; topk)
//   get_graph_feature    /root/reference/dgcnn.py:23-49      (gather, cat(nbr-ctr, ctr))
//   AttentionModule      /root/reference/layers_batch.py:28-39
//   TenorNetworkModule   /root/reference/layers_batch.py:70-83   + FC/sigmoid sg_net.py:131-136
//
// One CTA (8 warps) owns one graph; everything after the 60*N-byte input block stays in shared memory.
//
// Arithmetic form (exact refactor of the reference's 1x1 conv over the materialised [2C, N, k] edge tensor):
//     W [x_j - x_i ; x_i] = (A_j - A_i) + B_i,   A = Wa x (per node),  B = Wb x (per node)
//     max_j LReLU(BN(.))  = LReLU(alpha * ((max_j A_j - A_i) + B_i) + beta)   with alpha >= 0 (sign folded into W)
// so each layer is: k-NN (distance rows -> sorting network -> threshold select) -> per-node GEMM [R x C]x[C x 2C']
// -> gather-max over the neighbour rows of A.  Layer 1 of the xyz branch keeps the reference's direct form
// W_a (x_j - x_i): metre-scale coordinates would lose ~5 bits to cancellation in A_j - A_i.
//
// Zero padding (sg_net.py:258-262, 276-278: graphs are padded to node_num with all-zero nodes): trailing all-zero
// nodes have identical inputs, hence identical features, distances and neighbour sets in every layer.  The kernel
// works on R = (#nodes up to the last non-zero one) + 1 rows — the first pad stands for its whole class, with its
// multiplicity kept in the k-NN selection — and replicates that row before the attention stage.  Results are
// identical to processing every pad; graphs without trailing zero nodes simply have R = N.
//
// Schedule.  The kernel is issue/latency-bound (~1.6 kflop per input byte, 16 warps per SM), so the design
// minimises barriers and instruction count rather than bytes:
//   * every warp OWNS ceil(R/8) consecutive rows for the whole graph.  Per layer it runs, without any CTA barrier,
//     distance rows -> selection -> GEMM rows for its rows ("front"), then after ONE barrier the gather-max for
//     its rows ("back"), then ONE barrier before the next layer: 2 barriers per layer.
//   * the distance row of node i lives in the same shared-memory row that later receives A_i|B_i, so fronts of
//     different warps never touch each other's memory.
//   * packed fp32 FMA (fma.rn.f32x2, SASS FFMA2) for every dot product — the two halves carry the even- and
//     odd-channel partial sums; an in-register Batcher network + two cross-lane bitonic merges (4 lanes per row)
//     for the k-th-largest threshold instead of a shuffle-heavy warp sort.
//   * layer matrices are prefetched into shared memory with 1-D bulk TMA (cp.async.bulk + mbarrier) one layer ahead.
#pragma once
#include "common.cuh"
#include "topk_nth.cuh"

// tuning switches (defaults are the measured-best variants; see profiles/r01_variants.txt)
#ifndef SGPR_GRAM_V
#define SGPR_GRAM_V 0      // 0: explicit register double-buffering of the operand loads, 1: plain loop, unroll 4
#endif
#ifndef SGPR_GEMM_V
#define SGPR_GEMM_V 0      // 0: weights of the next channel pair prefetched, 1: plain loop, unroll 2
#endif
#ifndef SGPR_GATHER_W
#define SGPR_GATHER_W 3    // packed index words (4 neighbours each) fetched per gather step
#endif

namespace sgpr {

// Optional per-phase clock stamps of CTA 0 (debug builds only: -DSGPR_TIMELINE); see tools/timeline.py.
#ifdef SGPR_TIMELINE
__device__ long long g_timeline[kWarps * 128];
__device__ int g_smid[1024];
__device__ long long g_cta_t[2048];
__device__ int g_cta_g[1024];      // graph (and its active rows << 16) each CTA processed last
#define SGPR_TL(slot) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_timeline[(threadIdx.x >> 5) * 128 + (slot)] = clock64(); } while (0)
#else
#define SGPR_TL(slot) do { } while (0)
#endif

struct EmbedArgs {
    const float* g0;        // graphs of side 0 (or all graphs when !pairs)   [*, 15, N]
    const float* g1;        // graphs of side 1 (pairs mode)
    int G;                  // number of graphs to embed (2*B in pairs mode: g = 2*b + side)
    int N, k, KS;           // KS = k rounded up to 4 (neighbour-list row stride in 16-bit entries)
    int pairs;
    int dedup;              // 1: collapse trailing all-zero nodes into one row (exact); 0: process every node
    int compact;            // 0: graphs are [15][N] fp32 blocks; else: compact records of `compact` bytes each —
                            //    xyz [3][N] fp32 followed by N uint8 labels (0..11, anything else = no label), see sgpr_b200.h
    float* pooled;          // [G][32]
    float* att0;            // pairs: [B][N] side 0 ; else [G][N]   (may be null)
    float* att1;            // pairs: [B][N] side 1                 (may be null)
    float* emb;             // [G][N][32] or null
    float* score;           // [B]  (pairs)
    int* counters;          // [B]  arrival counters, zero on entry, zero on exit
    uint8_t* trace_knn;     // [G][6][N][k] or null
    float* trace_layers;    // [G][6][N][64] or null
    const int* order;       // [G] slot -> graph (heavy graphs placed so that co-resident CTAs balance), or null = identity
    int* work_ctr;          // non-null: CTAs pop slots from this counter (persistent launches, heaviest graph first)
    int split;              // 1: a work unit is one BRANCH of a graph (unit u = 2*slot + branch: xyz layers 1-3 | semantic layers
                            //    1-3, independent until conv_end, sg_net.py:84-104); the unit that finishes second merges.  Halves
                            //    the critical path of a graph when the launch has fewer graphs than CTA slots.  Same results.
    float* halves;          // split: [G][2][N][32] branch outputs handed to the merging unit
    int* gctr;              // split: [G] arrival counters, zero on entry, zero on exit
    int sm_count;           // SMs of the device (launch-side choice between the compilations of the kernel)
};

struct SmemLayout {
    int w, in, x, y, cat, xx, red, bar, idx, cnt, total;   // byte offsets
};

__host__ __device__ inline SmemLayout make_layout(int nmax, int ks) {
    SmemLayout L;
    int o = 0;
    L.w = o;   o += 64 * 128 * 4;              // largest packed layer matrix (64 in x 128 out)
    L.in = o;  o += ((kInCh * nmax * 4 + 15) / 16) * 16;
    L.x = o;   o += nmax * XS * 4;             // layer input / output, node-major
    L.y = o;   o += nmax * YS * 4;             // row i: distance row of node i, then A_i | B_i
    L.cat = o; o += nmax * XS * 4;             // layer-0 coordinates tile, then cat(xyz3, sem3)
    L.xx = o;  o += 2 * nmax * 4;              // squared norms of the layer input (+ layer-0 copy); later attention scores
    L.red = o; o += (kWarps * 32 + 64) * 4;
    L.bar = o; o += 16;
    L.idx = o; o += ((nmax * ks * 2 + 15) / 16) * 16;   // neighbour lists: ks 16-bit entries per node
    L.cnt = o; o += ((nmax + 15) / 16) * 16;
    L.total = o;
    return L;
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ void cmpx(float& a, float& b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}
// max of three in ONE instruction (sm_100 FMNMX3): the neighbour-max of the gather is bound by the half-rate min/max
// pipe, and max is exact and order-independent, so regrouping changes no bit
__device__ __forceinline__ float max3(float a, float b, float c) {
#ifdef SGPR_EMU
    return fmaxf(fmaxf(a, b), c);
#else
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
#endif
}
__device__ __forceinline__ float warp_sum(float s) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, d));
    return s;
}

// in-register ascending sort of EPL values (Batcher odd-even mergesort networks, generated headers)
template <int EPL>
__device__ __forceinline__ void sort_regs(float (&v)[EPL]) {
#define SGPR_CX(i, j) cmpx(v[i], v[j]);
    if constexpr (EPL == 8) {
#include "sortnet8.inc"
    } else if constexpr (EPL == 16) {
#include "sortnet16.inc"
    } else {
#include "sortnet32.inc"
    }
#undef SGPR_CX
}

// v[jt] for a warp-uniform jt, as a select tree (keeps v[] in registers)
template <int EPL>
__device__ __forceinline__ float pick_uniform(const float (&v)[EPL], int jt) {
    float cur[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) cur[i] = v[i];
    int bit = 1;
#pragma unroll
    for (int w = EPL / 2; w >= 1; w >>= 1) {
#pragma unroll
        for (int i = 0; i < w; ++i) cur[i] = (jt & bit) ? cur[2 * i + 1] : cur[2 * i];
        bit <<= 1;
    }
    return cur[0];
}

// ------------------------------------------------------------------------------------------------------------
// Distance rows of NR own nodes against all columns (dgcnn.py:15-17):
//     pd[i][c] = (2*dot(x_i,x_c) - xx_c) - xx_i      == -xx - inner - xx^T with inner = -2*dot, same rounding order
// lane <-> column (c = lane + 32q).  Row i is stored in its own A|B row, sY[i*YS + c], for EVERY c < 32*NPL so that the
// selection can read it with unconditional vector loads: columns R <= c < N (the collapsed zero pads) carry the pad
// class's value pd[i][R-1], columns >= N carry -inf.  Only the NQ = ceil(R/32) column blocks that hold active nodes
// are computed.
// ------------------------------------------------------------------------------------------------------------
// experiment switch (profiles/experiments/r02_embed_upper_bound.txt): -DSGPR_UB_EXPERIMENT=1 cuts the K loops of the
// 64-channel Gram / GEMM tiles to one step — wrong results, the time of a kernel whose contractions cost nothing
#ifndef SGPR_UB_EXPERIMENT
#define SGPR_UB_EXPERIMENT 0
#endif

template <int NPL, int NQ, int NR>
__device__ __forceinline__ void gram_rows(const float* __restrict__ sXt, const float* __restrict__ sXX,
                                          float* __restrict__ sY, int c4n, int R, int N, int r0, int lane) {
    if (SGPR_UB_EXPERIMENT && c4n == 16) c4n = 1;
    float2 acc[NR][NQ];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[r][q] = make_float2(0.0f, 0.0f);
    const float* pa = sXt + r0 * XS;
    const float* pb = sXt + lane * XS;
#if SGPR_GRAM_V == 1
#pragma unroll 4
    for (int c = 0; c < c4n; ++c) {
        float4 a[NR], b[NQ];
#pragma unroll
        for (int r = 0; r < NR; ++r) a[r] = *reinterpret_cast<const float4*>(pa + r * XS + 4 * c);
#pragma unroll
        for (int q = 0; q < NQ; ++q) b[q] = *reinterpret_cast<const float4*>(pb + 32 * q * XS + 4 * c);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                acc[r][q] = ffma2(make_float2(a[r].x, a[r].y), make_float2(b[q].x, b[q].y), acc[r][q]);
                acc[r][q] = ffma2(make_float2(a[r].z, a[r].w), make_float2(b[q].z, b[q].w), acc[r][q]);
            }
        }
    }
#else
    // software pipeline: the loads of channel group c+1 are in flight while the FFMA2s of group c issue
    float4 a[NR], b[NQ];
#pragma unroll
    for (int r = 0; r < NR; ++r) a[r] = *reinterpret_cast<const float4*>(pa + r * XS);
#pragma unroll
    for (int q = 0; q < NQ; ++q) b[q] = *reinterpret_cast<const float4*>(pb + 32 * q * XS);
#pragma unroll 2
    for (int c = 0; c < c4n; ++c) {
        float4 an[NR], bn[NQ];
        const int cn = min(c + 1, c4n - 1);
#pragma unroll
        for (int r = 0; r < NR; ++r) an[r] = *reinterpret_cast<const float4*>(pa + r * XS + 4 * cn);
#pragma unroll
        for (int q = 0; q < NQ; ++q) bn[q] = *reinterpret_cast<const float4*>(pb + 32 * q * XS + 4 * cn);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                acc[r][q] = ffma2(make_float2(a[r].x, a[r].y), make_float2(b[q].x, b[q].y), acc[r][q]);
                acc[r][q] = ffma2(make_float2(a[r].z, a[r].w), make_float2(b[q].z, b[q].w), acc[r][q]);
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) a[r] = an[r];
#pragma unroll
        for (int q = 0; q < NQ; ++q) b[q] = bn[q];
    }
#endif
    float pd[NR][NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const float xxc = sXX[lane + 32 * q];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const float dot = __fadd_rn(acc[r][q].x, acc[r][q].y);
            const float t = __fsub_rn(__fmul_rn(2.0f, dot), xxc);
            pd[r][q] = __fsub_rn(t, sXX[r0 + r]);
        }
    }
    const int qp = (R - 1) >> 5, lp = (R - 1) & 31;             // where column R-1 lives
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        float own = pd[r][0];
#pragma unroll
        for (int q = 1; q < NQ; ++q) own = (qp == q) ? pd[r][q] : own;
        const float padv = __shfl_sync(0xffffffffu, own, lp);
        float* dst = sY + (r0 + r) * YS + lane;
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int c = lane + 32 * q;
            float val = (c < N) ? padv : -INFINITY;
            if (q < NQ) val = (c < R) ? pd[r][q] : val;
            dst[32 * q] = val;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k-NN selection (dgcnn.py:19 `topk`) for up to 8 own rows r0 .. r0+nr-1: pick the k largest of the row's N
// distances, ties at the k-th value to the lowest column index.  The row holds 32*NPL values (gram_rows): columns
// R <= c < N all carry the pad-class value pd[i][R-1], columns >= N are -inf.  Only the SET matters downstream (max
// over neighbours), so the list is emitted in ascending column order with the pad class, if selected, represented
// once by column R-1, then padded to a multiple of 4 entries.  A list entry is the 16-bit word offset `column * scale`
// of the neighbour's row in the tile the gather reads (scale = row stride of that tile).
//
// 4 lanes cooperate on a row (lane = sub*8 + row), each holding EPL = NMAX/4 consecutive columns in registers:
// in-register sort, two cross-lane bitonic merges — the last one pruned to the single output position that is the
// threshold (element NMAX-k of the sorted row) — then count and emit passes over the unsorted values.
// ------------------------------------------------------------------------------------------------------------
#ifdef SGPR_EMU
__device__ __forceinline__ uint32_t fgt_mask(float a, float b) { return a > b ? 0xffffffffu : 0u; }
__device__ __forceinline__ uint32_t feq_mask(float a, float b) { return a == b ? 0xffffffffu : 0u; }
#else
__device__ __forceinline__ uint32_t fgt_mask(float a, float b) {          // all-ones if a > b (one FSET)
    uint32_t m;
    asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(a), "f"(b));
    return m;
}
__device__ __forceinline__ uint32_t feq_mask(float a, float b) {
    uint32_t m;
    asm("set.eq.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(a), "f"(b));
    return m;
}
#endif

// Selection with the reference's CPU tie behaviour (`knn_ties = cpu`, topk_nth.cuh): one lane runs ATen's
// nth_element over the row's N (value, column) pairs in place — the collapsed pad columns R..N-1 are present in the row
// with the pad class's value, exactly as an uncollapsed run would see them — and lists the first k columns, the pad
// class (columns >= R-1 when row R-1 is a pad) once.  A parity mode: ~8 lanes of a warp work, serially.
template <int NPL>
__device__ __forceinline__ void select_rows_cpu_rule(float* __restrict__ sY, uint16_t* __restrict__ sIdx,
                                                     uint8_t* __restrict__ sCnt, uint8_t* __restrict__ trace, int R, int N,
                                                     int k, int KS, int scale, int r0, int nr, int lane) {
    constexpr int NMAX = 32 * NPL;
    if (lane < nr) {
        const int row = r0 + lane;
        float* prow = sY + row * YS;
        uint8_t ix[NMAX];
        for (int c = 0; c < N; ++c) ix[c] = static_cast<uint8_t>(c);
        nth::topk_cpu_rule<uint8_t>(prow, ix, N, k);
        uint16_t* list = sIdx + row * KS;
        int pos = 0;
        bool pad_listed = false;
        uint16_t last = 0;
        for (int e = 0; e < k; ++e) {
            int c = ix[e];
            if (trace) trace[row * k + e] = static_cast<uint8_t>(c);
            if (c >= R - 1) {
                if (pad_listed) continue;
                pad_listed = true;
                c = R - 1;
            }
            last = static_cast<uint16_t>(c * scale);
            list[pos++] = last;
        }
        for (int e = pos; e < ((pos + 3) & ~3); ++e) list[e] = last;
        sCnt[row] = static_cast<uint8_t>((pos + 3) & ~3);
    }
}

template <int NPL, int TIES = 0>
__device__ __forceinline__ void select_rows(float* __restrict__ sY, uint16_t* __restrict__ sIdx,
                                            uint8_t* __restrict__ sCnt, uint8_t* __restrict__ trace, int R, int N, int k,
                                            int KS, int scale, int r0, int nr, int lane) {
    if constexpr (TIES == 1) {
        select_rows_cpu_rule<NPL>(sY, sIdx, sCnt, trace, R, N, k, KS, scale, r0, nr, lane);
        return;
    }
    constexpr int NMAX = 32 * NPL;
    constexpr int EPL = 8 * NPL;
    const int sub = lane >> 3, rl = lane & 7;
    const int row = r0 + min(rl, nr - 1);             // lanes beyond nr shadow the last row (they never store)
    const bool live = rl < nr;
    const float* prow = sY + row * YS;
    const int c0 = sub * EPL;
    float o[EPL], v[EPL];
#pragma unroll
    for (int j4 = 0; j4 < EPL / 4; ++j4) {
        const float4 x = *reinterpret_cast<const float4*>(prow + c0 + 4 * j4);
        o[4 * j4] = x.x; o[4 * j4 + 1] = x.y; o[4 * j4 + 2] = x.z; o[4 * j4 + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < EPL; ++j) v[j] = o[j];
    sort_regs<EPL>(v);
    // ---- merge across the 4 lanes of the row (bitonic "flip" merges; every exchange ascending) ----
    {   // lanes (0,1) and (2,3): mirror exchange, element j meets the partner's element EPL-1-j, then sort in place
        const bool keep_min = (sub & 1) == 0;
#pragma unroll
        for (int j = 0; j < EPL / 2; ++j) {
            const float pa = __shfl_xor_sync(0xffffffffu, v[EPL - 1 - j], 8);
            const float pb = __shfl_xor_sync(0xffffffffu, v[j], 8);
            v[j] = keep_min ? fminf(v[j], pa) : fmaxf(v[j], pa);
            v[EPL - 1 - j] = keep_min ? fminf(v[EPL - 1 - j], pb) : fmaxf(v[EPL - 1 - j], pb);
        }
#pragma unroll
        for (int d = EPL / 2; d >= 1; d >>= 1)
#pragma unroll
            for (int j = 0; j < EPL; ++j)
                if ((j & d) == 0) cmpx(v[j], v[j | d]);
    }
    {   // pairs (0,1) with (2,3): mirror exchange with lane sub^3, half-cleaner with lane sub^1
        const bool keep_lo = (sub & 2) == 0;
#pragma unroll
        for (int j = 0; j < EPL / 2; ++j) {
            const float pa = __shfl_xor_sync(0xffffffffu, v[EPL - 1 - j], 24);
            const float pb = __shfl_xor_sync(0xffffffffu, v[j], 24);
            v[j] = keep_lo ? fminf(v[j], pa) : fmaxf(v[j], pa);
            v[EPL - 1 - j] = keep_lo ? fminf(v[EPL - 1 - j], pb) : fmaxf(v[EPL - 1 - j], pb);
        }
        const bool keep_min = (sub & 1) == 0;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const float p = __shfl_xor_sync(0xffffffffu, v[j], 8);
            v[j] = keep_min ? fminf(v[j], p) : fmaxf(v[j], p);
        }
    }
    // ---- threshold: k-th largest = sorted position P = NMAX - k, i.e. element P % EPL of lane sub = P / EPL.  Each
    // lane now holds a bitonic-cleaned block whose in-place sort would be log2(EPL) half-cleaner stages; only the one
    // output position is needed, so every stage keeps just the half that feeds it (min or max by that bit of P).
    const int P = NMAX - k;
    const int pe = P % EPL;
    float tl;
    {
        float cur[EPL / 2];
        {
            const bool up = (pe & (EPL / 2)) != 0;
#pragma unroll
            for (int j = 0; j < EPL / 2; ++j) cur[j] = up ? fmaxf(v[j], v[j + EPL / 2]) : fminf(v[j], v[j + EPL / 2]);
        }
#pragma unroll
        for (int d = EPL / 4; d >= 1; d >>= 1) {
            const bool up = (pe & d) != 0;
#pragma unroll
            for (int j = 0; j < d; ++j) cur[j] = up ? fmaxf(cur[j], cur[j + d]) : fminf(cur[j], cur[j + d]);
        }
        tl = cur[0];
    }
    const float thr = __shfl_sync(0xffffffffu, tl, rl + (P / EPL) * 8);
    // ---- count pass: per-lane bit masks over its EPL columns ----
    uint32_t mgt = 0, meq = 0;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
        mgt |= fgt_mask(o[j], thr) & (1u << j);
        meq |= feq_mask(o[j], thr) & (1u << j);
    }
    const int ne = min(max(R - c0, 0), EPL);          // emittable columns of this lane (c < R)
    const uint32_t em = (ne >= 32) ? 0xffffffffu : ((1u << ne) - 1u);
    const int gt_all = __popc(mgt), gt_e = __popc(mgt & em), eq_e = __popc(meq & em);
    int gts[4], eqs[4], gt_tot = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        gt_tot += __shfl_sync(0xffffffffu, gt_all, rl + s * 8);
        gts[s] = __shfl_sync(0xffffffffu, gt_e, rl + s * 8);
        eqs[s] = __shfl_sync(0xffffffffu, eq_e, rl + s * 8);
    }
    const int need = k - gt_tot;                      // tied-at-threshold elements to take, lowest columns first (>= 1)
    int base = 0, total = 0, eq_before = 0, eb = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int take = gts[s] + max(0, min(eqs[s], need - eb));
        if (s < sub) base += take;
        if (s == sub) eq_before = eb;
        total += take;
        eb += eqs[s];
    }
    // ---- emit pass ----
    if (live) {
        uint16_t* list = sIdx + row * KS;
        uint32_t ties = meq & em, keep = 0;
        for (int t = min(max(need - eq_before, 0), eq_e); t > 0; --t) {   // lowest t tied columns (t is 0 or 1 but for pads)
            const uint32_t low = ties & (0u - ties);
            keep |= low;
            ties ^= low;
        }
        uint32_t taken = (mgt & em) | keep;
        int pos = base;
        uint16_t last = 0;
        while (taken) {
            const int j = __ffs(taken) - 1;
            last = static_cast<uint16_t>((c0 + j) * scale);
            list[pos++] = last;
            taken &= taken - 1u;
        }
        // the lane that wrote the last entry pads the list to a multiple of 4 (repeats are harmless under max)
        if (pos == total && pos > base) {
            for (int e = total; e < ((total + 3) & ~3); ++e) list[e] = last;
        }
        if (sub == 0) sCnt[row] = static_cast<uint8_t>((total + 3) & ~3);
        // debug tap: the k columns an undeduplicated run lists (pad class expanded, lowest index first)
        if (trace && sub == 0) {
            int ngt = 0;
            for (int c = 0; c < N; ++c) ngt += prow[c] > thr;
            int nd = k - ngt, p = 0;
            for (int c = 0; c < N && p < k; ++c) {
                const float x = prow[c];
                if (x > thr || (x == thr && nd-- > 0)) trace[row * k + p++] = static_cast<uint8_t>(c);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// GEMM rows: out[n][co] = sum_ci X[n][ci] * W[ci][co] for NR own nodes, co in [0, 32*CPL); a lane owns CPL
// consecutive outputs.  W is packed in channel PAIRS (pack.hpp::pair_index): row p holds, for every output,
// (W[2p][co], W[2p+1][co]) so that one FFMA2 advances the even- and odd-channel partial sums of an output.
// EPI: 0 = store raw, 1 = BN(alpha,beta)+LeakyReLU (conv_end, sg_net.py:74-76,105).
// ------------------------------------------------------------------------------------------------------------
template <int NR, int CPL, int EPI>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ sXin, const float* __restrict__ sW,
                                          float* __restrict__ sOut, int outStride, const float* __restrict__ ab,
                                          int cin4, int r0, int lane) {
    if (SGPR_UB_EXPERIMENT && cin4 == 16 && EPI == 0) cin4 = 1;
    constexpr int CO = 32 * CPL;
    constexpr int ROW = 2 * CO;                  // floats per channel-pair row
    float2 acc[NR][CPL];
#pragma unroll
    for (int n = 0; n < NR; ++n)
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[n][j] = make_float2(0.0f, 0.0f);
    const float* px = sXin + r0 * XS;
    auto load_w = [&](float2 (&w)[CPL], int p) {
        const float* wp = sW + p * ROW;
        if constexpr (CPL == 4) {
            const float4 t0 = *reinterpret_cast<const float4*>(wp + lane * 4);
            const float4 t1 = *reinterpret_cast<const float4*>(wp + 128 + lane * 4);
            w[0] = make_float2(t0.x, t0.y); w[1] = make_float2(t0.z, t0.w);
            w[2] = make_float2(t1.x, t1.y); w[3] = make_float2(t1.z, t1.w);
        } else if constexpr (CPL == 2) {
            const float4 t0 = *reinterpret_cast<const float4*>(wp + lane * 4);
            w[0] = make_float2(t0.x, t0.y); w[1] = make_float2(t0.z, t0.w);
        } else {
            w[0] = *reinterpret_cast<const float2*>(wp + lane * 2);
        }
    };
#if SGPR_GEMM_V == 1
#pragma unroll 2
    for (int c4 = 0; c4 < cin4; ++c4) {
        float4 x[NR];
#pragma unroll
        for (int n = 0; n < NR; ++n) x[n] = *reinterpret_cast<const float4*>(px + n * XS + 4 * c4);
        float2 wa[CPL], wb[CPL];
        load_w(wa, 2 * c4);
        load_w(wb, 2 * c4 + 1);
#pragma unroll
        for (int n = 0; n < NR; ++n)
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[n][j] = ffma2(make_float2(x[n].x, x[n].y), wa[j], acc[n][j]);
#pragma unroll
        for (int n = 0; n < NR; ++n)
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[n][j] = ffma2(make_float2(x[n].z, x[n].w), wb[j], acc[n][j]);
    }
#else
    // software pipeline over channel pairs: the weights of pair p+1 load while the FFMA2s of pair p issue
    const int npairs = 2 * cin4;
    float2 w0[CPL], w1[CPL];
    load_w(w0, 0);
#pragma unroll 2
    for (int c4 = 0; c4 < cin4; ++c4) {
        float4 x[NR];
#pragma unroll
        for (int n = 0; n < NR; ++n) x[n] = *reinterpret_cast<const float4*>(px + n * XS + 4 * c4);
        load_w(w1, 2 * c4 + 1);
#pragma unroll
        for (int n = 0; n < NR; ++n)
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[n][j] = ffma2(make_float2(x[n].x, x[n].y), w0[j], acc[n][j]);
        load_w(w0, min(2 * c4 + 2, npairs - 1));
#pragma unroll
        for (int n = 0; n < NR; ++n)
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[n][j] = ffma2(make_float2(x[n].z, x[n].w), w1[j], acc[n][j]);
    }
#endif
    float al[CPL], be[CPL];
    if constexpr (EPI == 1) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) { al[j] = __ldg(ab + lane * CPL + j); be[j] = __ldg(ab + CO + lane * CPL + j); }
    }
    if constexpr (EPI == 1) __syncwarp();          // conv_end may run in place (output rows = input rows): reads first
#pragma unroll
    for (int n = 0; n < NR; ++n) {
        float y[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            y[j] = __fadd_rn(acc[n][j].x, acc[n][j].y);
            if constexpr (EPI == 1) y[j] = lrelu(fmaf(y[j], al[j], be[j]));
        }
        float* op = sOut + (r0 + n) * outStride + lane * CPL;
        if constexpr (CPL == 4) *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
        else if constexpr (CPL == 2) *reinterpret_cast<float2*>(op) = make_float2(y[0], y[1]);
        else *op = y[0];
    }
}

// squared norms (dgcnn.py:16 term of the next layer) of up to 8 own rows: 4 lanes per row, each sums a quarter of the
// channels in order, partials combined pairwise.  Call under __syncwarp after the rows were written by this warp.
__device__ __forceinline__ void norms_rows(const float* __restrict__ sT, float* __restrict__ sXX, int c4n, int r0, int r1,
                                           int lane) {
    const int part = lane >> 3, rl = lane & 7;
    const int per = (c4n + 3) >> 2;
    for (int b0 = r0; b0 < r1; b0 += 8) {
        const int row = min(b0 + rl, r1 - 1);
        float s = 0.0f;
        for (int g = part * per; g < min(c4n, (part + 1) * per); ++g) {
            const float4 x = *reinterpret_cast<const float4*>(sT + row * XS + 4 * g);
            s = __fadd_rn(s, __fmul_rn(x.x, x.x));
            s = __fadd_rn(s, __fmul_rn(x.y, x.y));
            s = __fadd_rn(s, __fmul_rn(x.z, x.z));
            s = __fadd_rn(s, __fmul_rn(x.w, x.w));
        }
        s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));
        s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 16));
        if (part == 0 && b0 + rl < r1) sXX[b0 + rl] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Gather-max + BN + LeakyReLU for own rows (sg_net.py:85-86 etc.): for node i and channel c
//     out = LReLU(alpha_c * ((max_{j in knn(i)} A[j][c] - A[i][c]) + B[i][c]) + beta_c)
// sY row = [A(0..COUT) | B(COUT..2COUT)]; a lane owns COUT/32 channels.  The neighbour list holds 16-bit word offsets
// (j*YS) of the neighbour rows; 4*GW of them are fetched per step (GW 8-byte list words, then 4*GW independent row
// loads) to keep the shared-memory pipe busy.
// ------------------------------------------------------------------------------------------------------------
// Output of a layer as tensor-core operand planes (embed_tc_kernel.cuh): z = big + small, big = z rounded to TF32, both in
// the K-major 128-byte-swizzled image of tc_ops.cuh (two atoms of 32 channels), plus the row's squared norm.
struct PlaneOut {
    float* big;       // [2 atoms][64 rows][32]
    float* small;
    float* xx;        // [64] squared norms of the rows (dgcnn.py:16 term of the next layer)
};
#ifdef SGPR_EMU
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
#else
__device__ __forceinline__ float tf32_round(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
#endif
// a lane's channel pair (2*lane, 2*lane+1) of row i -> planes, and the row norm (all lanes call; lane 0 stores it)
__device__ __forceinline__ void store_planes(const PlaneOut& P, int i, int lane, float z0, float z1) {
    const int c = 2 * lane;
    const uint32_t off = static_cast<uint32_t>(c >> 5) * 2048u + static_cast<uint32_t>(i) * 32u +
                         (((((c & 31) >> 2) ^ (i & 7)) << 2) | (c & 3));
    const float b0 = tf32_round(z0), b1 = tf32_round(z1);
    *reinterpret_cast<float2*>(P.big + off) = make_float2(b0, b1);
    *reinterpret_cast<float2*>(P.small + off) = make_float2(__fsub_rn(z0, b0), __fsub_rn(z1, b1));
    const float s = warp_sum(__fadd_rn(__fmul_rn(z0, z0), __fmul_rn(z1, z1)));
    if (lane == 0) P.xx[i] = s;
}

template <int COUT, int PLANES = 0>
__device__ __forceinline__ void gather_rows(const float* __restrict__ sY, const uint16_t* __restrict__ sIdx,
                                            const uint8_t* __restrict__ sCnt, int KS, const float* __restrict__ ab,
                                            float* __restrict__ sDst, float* __restrict__ trace, int r0, int r1, int lane,
                                            const PlaneOut* planes = nullptr) {
    constexpr int CPL = COUT / 32;
    constexpr int GW = SGPR_GATHER_W;
    float al[CPL], be[CPL];
#pragma unroll
    for (int p = 0; p < CPL; ++p) { al[p] = __ldg(ab + lane * CPL + p); be[p] = __ldg(ab + COUT + lane * CPL + p); }
    const float* base = sY + lane * CPL;
#pragma unroll 1
    for (int i = r0; i < r1; ++i) {
        float m[CPL];
#pragma unroll
        for (int p = 0; p < CPL; ++p) m[p] = -INFINITY;
        const uint2* row = reinterpret_cast<const uint2*>(sIdx + i * KS);
        const int nw = sCnt[i] >> 2;                    // list words (4 neighbours each), >= 1
        float ai[CPL], bi[CPL];
#pragma unroll
        for (int p = 0; p < CPL; ++p) { ai[p] = base[i * YS + p]; bi[p] = base[i * YS + COUT + p]; }
#pragma unroll 1
        for (int t = 0; t < nw; t += GW) {
            uint2 wd[GW];
#pragma unroll
            for (int q = 0; q < GW; ++q) wd[q] = row[min(t + q, nw - 1)];      // tail words repeat (harmless under max)
            float v[4 * GW][CPL];
#pragma unroll
            for (int e = 0; e < 4 * GW; ++e) {
                const uint32_t w = (e & 2) ? wd[e >> 2].y : wd[e >> 2].x;
                const uint32_t off = (e & 1) ? (w >> 16) : (w & 0xffffu);
                if constexpr (CPL == 2) {
                    const float2 a = *reinterpret_cast<const float2*>(base + off);
                    v[e][0] = a.x; v[e][1] = a.y;
                } else {
                    v[e][0] = base[off];
                }
            }
#pragma unroll
            for (int p = 0; p < CPL; ++p) {
#pragma unroll
                for (int q = 0; q < GW; ++q)      // 4 neighbours + the running max: two 3-input max
                    m[p] = max3(m[p], max3(v[4 * q][p], v[4 * q + 1][p], v[4 * q + 2][p]), v[4 * q + 3][p]);
            }
        }
        float zz[CPL];
#pragma unroll
        for (int p = 0; p < CPL; ++p) {
            const float y = __fadd_rn(__fsub_rn(m[p], ai[p]), bi[p]);
            const float z = lrelu(fmaf(y, al[p], be[p]));
            zz[p] = z;
            if constexpr (!PLANES) sDst[i * XS + lane * CPL + p] = z;
            if (trace) trace[i * 64 + lane * CPL + p] = z;
        }
        if constexpr (PLANES) {
            static_assert(!PLANES || CPL == 2, "operand planes are 64 channels wide");
            store_planes(*planes, i, lane, zz[0], zz[CPL - 1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// xyz layer 1 (3 -> 64) in the reference's direct form (sg_net.py:84-86) for own rows: per edge
//     e = wa0*d0 + wa1*d1 + wa2*d2  (d = x_j - x_i, sequential FMA), max over the edges, then the centre
//     terms wb.x_i appended in the same sequential order (monotone in e, so they commute with the max).
// A lane owns output channels 2*lane, 2*lane+1; neighbour coordinates come from the layer-0 node tile (list entries
// are word offsets j*XS into it).
// ------------------------------------------------------------------------------------------------------------
template <int PLANES = 0>
__device__ __forceinline__ void xyz_rows(const float* __restrict__ sT, const uint16_t* __restrict__ sIdx,
                                         const uint8_t* __restrict__ sCnt, int KS, const float* __restrict__ s1,
                                         float* __restrict__ sDst, float* __restrict__ trace, int r0, int r1, int lane,
                                         const PlaneOut* planes = nullptr) {
    // sT: the (x, y, z, 0) node tile of layer 0 (stride XS).  p0 = {wa0,wa1,wa2,wb0}, p1 = {wb1,wb2,alpha,beta} of
    // channel 2*lane; q0/q1 the same for channel 2*lane+1
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane) * 2);
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane) * 2 + 1);
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane + 1) * 2);
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane + 1) * 2 + 1);
#pragma unroll 1
    for (int i = r0; i < r1; ++i) {
        const float4 xi = *reinterpret_cast<const float4*>(sT + i * XS);
        float m0 = -INFINITY, m1 = -INFINITY;
        const uint2* row = reinterpret_cast<const uint2*>(sIdx + i * KS);
        const int nw = sCnt[i] >> 2;
#pragma unroll 1
        for (int t = 0; t < nw; ++t) {
            const uint2 wd = row[t];
            const uint32_t js[4] = {wd.x & 0xffffu, wd.x >> 16, wd.y & 0xffffu, wd.y >> 16};
            float e0[4], e1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 xj = *reinterpret_cast<const float4*>(sT + js[u]);
                const float d0 = __fsub_rn(xj.x, xi.x), d1 = __fsub_rn(xj.y, xi.y), d2 = __fsub_rn(xj.z, xi.z);
                e0[u] = fmaf(p0.z, d2, fmaf(p0.y, d1, __fmul_rn(p0.x, d0)));
                e1[u] = fmaf(q0.z, d2, fmaf(q0.y, d1, __fmul_rn(q0.x, d0)));
            }
            m0 = max3(m0, max3(e0[0], e0[1], e0[2]), e0[3]);
            m1 = max3(m1, max3(e1[0], e1[1], e1[2]), e1[3]);
        }
        float y0 = fmaf(p0.w, xi.x, m0); y0 = fmaf(p1.x, xi.y, y0); y0 = fmaf(p1.y, xi.z, y0);
        float y1 = fmaf(q0.w, xi.x, m1); y1 = fmaf(q1.x, xi.y, y1); y1 = fmaf(q1.y, xi.z, y1);
        const float z0 = lrelu(fmaf(y0, p1.z, p1.w));
        const float z1 = lrelu(fmaf(y1, q1.z, q1.w));
        if constexpr (PLANES) store_planes(*planes, i, lane, z0, z1);
        else *reinterpret_cast<float2*>(sDst + i * XS + 2 * lane) = make_float2(z0, z1);
        if (trace) { trace[i * 64 + 2 * lane] = z0; trace[i * 64 + 2 * lane + 1] = z1; }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pair head: NTN + FC + sigmoid for one ordered pair (layers_batch.py:70-83, sg_net.py:131-136).
// e1/e2: 32 pooled floats each (shared memory). scratch: >= 512+64 floats of shared memory.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_acc(float z) { return 1.0f / (1.0f + expf(-z)); }

__device__ __forceinline__ float group16_sum(float v) {       // sum over the 16 lanes of a half-warp (tree order)
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

__device__ __forceinline__ void pair_head_cta(const float* __restrict__ e1, const float* __restrict__ e2,
                                              const PackedWeights& W, const HeadParams& H, float* __restrict__ scratch,
                                              float* __restrict__ score_out, int tid) {
    float* P = scratch;          // [512]  P[b*16+t] = sum_a e1[a] * W[a][b*16+t]   (layers_batch.py:78)
    float* s = scratch + 512;    // [16]
    // all 64 weight loads of a thread's two outputs are issued before the first FMA: one L2 round trip, not eight
    const bool act = tid < 256;                  // the head is laid out for 256 threads; extra warps only keep the barriers
    if (act) {
        float w0[kF3], w1[kF3];
#pragma unroll
        for (int a = 0; a < kF3; ++a) {
            w0[a] = __ldg(W.ntn_w + a * 512 + tid);
            w1[a] = __ldg(W.ntn_w + a * 512 + 256 + tid);
        }
        float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
        for (int a = 0; a < kF3; ++a) { acc0 = fmaf(e1[a], w0[a], acc0); acc1 = fmaf(e1[a], w1[a], acc1); }
        P[tid] = acc0;
        P[256 + tid] = acc1;
    }
    __syncthreads();
    // thread (t, part): t = tid / 16 is the NTN neuron, part = tid % 16 a slice of the contraction
    const int t = tid >> 4, part = tid & 15;
    if (act) {
        float bil = fmaf(P[(2 * part + 1) * kT + t], e2[2 * part + 1], __fmul_rn(P[(2 * part) * kT + t], e2[2 * part]));   // :79
        float blk = 0.0f;                                                                                                 // :80-81
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = part * 4 + q;
            blk = fmaf(__ldg(W.ntn_v + t * 64 + c), (c < kF3) ? e1[c] : e2[c - kF3], blk);
        }
        bil = group16_sum(bil);
        blk = group16_sum(blk);
        if (part == 0) s[t] = fmaxf(__fadd_rn(__fadd_rn(bil, blk), __ldg(W.ntn_b + t)), 0.0f);                            // :82
    }
    __syncthreads();
    if (act) {
        // h[u] = relu(fc1_w[u] . s + fc1_b[u]) with (u, part) = (t, part); then score = sigmoid(fc2_w . h + fc2_b)
        float h = group16_sum(__fmul_rn(s[part], H.fc1_w[t * kT + part]));                                                // sg_net.py:134
        h = fmaxf(__fadd_rn(h, H.fc1_b[t]), 0.0f);
        // one value per half-warp -> collect the 16 h's in warp 0 via shared memory
        if (part == 0) P[t] = __fmul_rn(h, H.fc2_w[t]);
    }
    __syncthreads();
    if (tid < 32) {
        float z = (tid < kBn) ? P[tid] : 0.0f;                                                                          // sg_net.py:136
        z = group16_sum(z);
        if (tid == 0) *score_out = sigmoidf_acc(__fadd_rn(z, H.fc2_b));
    }
}

// ------------------------------------------------------------------------------------------------------------
// Everything after the EdgeConv stack for one graph (shared by the FFMA kernel below and the tensor-core kernel of
// embed_tc_kernel.cuh): replicate the collapsed pad rows, attention pooling over all N nodes (layers_batch.py:28-39),
// pooled vector out, and — in pairs mode — the pair head in whichever CTA of the pair finishes second.
// sE: node embeddings [n][XS] (first 32 columns), complete for rows < R behind a CTA barrier; sScratch: >= 576 floats.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void finish_graph(const EmbedArgs& A, const PackedWeights& W, const HeadParams& H, float* sE,
                                             float* sRed, float* sScratch, int* sFlag, int g, int N, int R, int tid,
                                             int warp, int lane) {
    // every trailing pad is a copy of row R-1
    for (int e = tid; e < (N - R) * kF3; e += kThreads) sE[(R + (e >> 5)) * XS + (e & 31)] = sE[(R - 1) * XS + (e & 31)];
    __syncthreads();
    if (A.emb) {
        float* eo = A.emb + static_cast<size_t>(g) * N * kF3;
        for (int e = tid; e < N * kF3; e += kThreads) eo[e] = sE[(e >> 5) * XS + (e & 31)];
    }

    // ================= attention pooling over all N nodes (layers_batch.py:28-39) =================
    // warp w handles nodes w, w+8, ...; lane = feature index
    float* sCtx = sRed + kWarps * 32;        // [32]
    float* sPool = sRed + kWarps * 32 + 32;  // [32]
    {   // ctx[b] = tanh(mean_n sum_a E[n][a] Watt[a][b])
        float wcol[kF3];
#pragma unroll
        for (int a = 0; a < kF3; ++a) wcol[a] = __ldg(W.att_w + a * kF3 + lane);
        float colsum = 0.0f;
#pragma unroll 2
        for (int n = warp; n < N; n += kWarps) {
            float t = 0.0f;
#pragma unroll
            for (int a4 = 0; a4 < kF3 / 4; ++a4) {
                const float4 e = *reinterpret_cast<const float4*>(sE + n * XS + 4 * a4);
                t = fmaf(e.x, wcol[4 * a4 + 0], t);
                t = fmaf(e.y, wcol[4 * a4 + 1], t);
                t = fmaf(e.z, wcol[4 * a4 + 2], t);
                t = fmaf(e.w, wcol[4 * a4 + 3], t);
            }
            colsum = __fadd_rn(colsum, t);
        }
        sRed[warp * 32 + lane] = colsum;
    }
    __syncthreads();
    if (tid < kF3) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s = __fadd_rn(s, sRed[w * 32 + tid]);
        sCtx[tid] = tanhf(s / static_cast<float>(N));
    }
    __syncthreads();
    {   // att[n] = sigmoid(E[n] . ctx); pooled[a] = sum_n E[n][a] att[n]   (per-warp partials, then 8-way sum)
        const float cb = sCtx[lane];
        float* ao = A.pairs ? ((g & 1) ? A.att1 : A.att0) : A.att0;
        if (ao) ao += static_cast<size_t>(A.pairs ? (g >> 1) : g) * N;
        float pool = 0.0f;
        for (int n0 = warp; n0 < N; n0 += 8 * kWarps) {          // up to 8 nodes of this warp at a time
            float e[8], d[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int n = n0 + u * kWarps;
                e[u] = (n < N) ? sE[n * XS + lane] : 0.0f;
                d[u] = __fmul_rn(e[u], cb);
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1)
#pragma unroll
                for (int u = 0; u < 8; ++u) d[u] = __fadd_rn(d[u], __shfl_xor_sync(0xffffffffu, d[u], sft));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int n = n0 + u * kWarps;
                if (n < N) {
                    const float a = sigmoidf_acc(d[u]);
                    if (ao && lane == 0) ao[n] = a;
                    pool = fmaf(e[u], a, pool);
                }
            }
        }
        sRed[warp * 32 + lane] = pool;
    }
    __syncthreads();
    if (tid < kF3) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s = __fadd_rn(s, sRed[w * 32 + tid]);
        sPool[tid] = s;
        A.pooled[static_cast<size_t>(g) * kF3 + tid] = s;
    }

    SGPR_TL(60);
    // ================= pair head, run by whichever CTA of the pair finishes last =================
    if (A.pairs) {
        __threadfence();
        __syncthreads();
        if (tid == 0) *sFlag = atomicAdd(A.counters + (g >> 1), 1);
        __syncthreads();
        if (*sFlag == 1) {
            __threadfence();
            float* e1 = sRed;        // side 0 pooled
            float* e2 = sRed + 32;   // side 1 pooled
            if (tid < 64) {
                const int side = tid >> 5, a = tid & 31;
                const float v = __ldcg(A.pooled + (static_cast<size_t>(g & ~1) + side) * kF3 + a);
                (side ? e2 : e1)[a] = v;
            }
            if (tid == 0) A.counters[g >> 1] = 0;
            __syncthreads();
            pair_head_cta(e1, e2, W, H, sScratch, A.score + (g >> 1), tid);
        }
    }
}

// One entry of the layer loop (the six EdgeConv layers, sg_net.py:84-102).
struct LayerDesc {
    const float* ab;       // alpha | beta
    const float* next_w;   // matrix (or matrices) to prefetch into sW once every warp's GEMM has consumed sW
    int next_bytes;
    int cin4;              // input channels / 4
    int cout;              // 64 or 32
};

// everything a warp's "front" needs, to keep the NR dispatch readable
struct FrontCtx {
    const float* sXt;      // this layer's input tile (node-major)
    const float* sXX;
    float* sY;
    const float* sW;
    uint16_t* sIdx;
    uint8_t* sCnt;
    uint8_t* trace_knn;
    int c4n, cout, R, N, k, KS, layer;
};

#define SGPR_NR_SWITCH(nr, CALL)                                                                   \
    switch (nr) {                                                                                  \
        case 1: { constexpr int NR = 1; CALL; } break;                                             \
        case 2: { constexpr int NR = 2; CALL; } break;                                             \
        case 3: { constexpr int NR = 3; CALL; } break;                                             \
        case 4: { constexpr int NR = 4; CALL; } break;                                             \
        case 5: { constexpr int NR = 5; CALL; } break;                                             \
        case 6: { constexpr int NR = 6; CALL; } break;                                             \
        case 7: { constexpr int NR = 7; CALL; } break;                                             \
        default: { constexpr int NR = 8; CALL; } break;                                            \
    }

// front of one pass (up to 8 own rows): distance rows -> selection -> GEMM rows.  Only the row-tiled pieces are
// specialised on the row count; the selection network exists once.
template <int NPL, int TIES>
__device__ __forceinline__ void front_pass(const FrontCtx& F, int r0, int nr, int lane, uint64_t* barW, uint32_t& phW,
                                           bool& waited) {
    // column blocks holding active nodes: all of them, or (small graphs) the lower half
    constexpr int NQH = (NPL >= 2) ? NPL / 2 : 1;
    if (F.R <= 32 * NQH) { SGPR_NR_SWITCH(nr, (gram_rows<NPL, NQH, NR>(F.sXt, F.sXX, F.sY, F.c4n, F.R, F.N, r0, lane))) }
    else                 { SGPR_NR_SWITCH(nr, (gram_rows<NPL, NPL, NR>(F.sXt, F.sXX, F.sY, F.c4n, F.R, F.N, r0, lane))) }
    __syncwarp();
    SGPR_TL(8 + F.layer * 8 + 1);
    select_rows<NPL, TIES>(F.sY, F.sIdx, F.sCnt, F.trace_knn, F.R, F.N, F.k, F.KS, (F.layer == 0) ? XS : YS, r0, nr, lane);
    __syncwarp();                                      // the distance rows are dead; A|B may overwrite them
    SGPR_TL(8 + F.layer * 8 + 2);
    if (F.layer != 0) {
        if (!waited) { mbar_wait(barW, phW); phW ^= 1; waited = true; }
        if (F.cout == 64) { SGPR_NR_SWITCH(nr, (gemm_rows<NR, 4, 0>(F.sXt, F.sW, F.sY, YS, nullptr, F.c4n, r0, lane))) }
        else              { SGPR_NR_SWITCH(nr, (gemm_rows<NR, 2, 0>(F.sXt, F.sW, F.sY, YS, nullptr, F.c4n, r0, lane))) }
    }
}

__device__ __forceinline__ void conv_end_dispatch(const float* sCat, const float* sWend, float* sE, const float* ab, int r0,
                                                  int nr, int lane) {
    SGPR_NR_SWITCH(nr, (gemm_rows<NR, 1, 1>(sCat, sWend, sE, XS, ab, 16, r0, lane)))
}

// ------------------------------------------------------------------------------------------------------------
// The fused kernel.
// ------------------------------------------------------------------------------------------------------------
#ifndef SGPR_MINBLOCKS_SMALL
#define SGPR_MINBLOCKS_SMALL 2
#endif
// TIES: 0 = lowest index first (ATen CUDA topk), 1 = ATen CPU nth_element order.  SPLIT: 1 = one branch of a graph per
// work unit (EmbedArgs::split) — its own instantiation so that the whole-graph kernel's register allocation is untouched.
// RICH: 1 = compiled for ONE CTA per SM (up to 255 registers, no spills) — launches whose work units all have an SM to
// themselves anyway (5 % off the small-batch latency floor).
template <int NPL, int TIES = 0, int SPLIT = 0, int RICH = 0>
__global__ void __launch_bounds__(kThreads, (NPL <= 2 && !RICH) ? SGPR_MINBLOCKS_SMALL : 1)
sgpr_embed_kernel(const EmbedArgs A, const PackedWeights W, const HeadParams H) {
    constexpr int NMAX = 32 * NPL;
    SGPR_DYN_SMEM(smem);
    const SmemLayout L = make_layout(NMAX, A.KS);
    float* sW = reinterpret_cast<float*>(smem + L.w);
    float* sIn = reinterpret_cast<float*>(smem + L.in);
    float* sX = reinterpret_cast<float*>(smem + L.x);
    float* sY = reinterpret_cast<float*>(smem + L.y);
    float* sCat = reinterpret_cast<float*>(smem + L.cat);
    float* sXX = reinterpret_cast<float*>(smem + L.xx);        // [NMAX] current layer input norms
    float* sXX0 = sXX + NMAX;                                  // [NMAX] norms of the layer-0 coordinates
    float* sRed = reinterpret_cast<float*>(smem + L.red);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar);
    uint16_t* sIdx = reinterpret_cast<uint16_t*>(smem + L.idx);
    uint8_t* sCnt = smem + L.cnt;
    __shared__ int sFlag;
    __shared__ int sLast[kWarps];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = A.N, k = A.k, KS = A.KS;
    uint64_t* barIn = bars;
    uint64_t* barW = bars + 1;
    uint32_t phIn = 0, phW = 0;

#ifdef SGPR_TIMELINE
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        g_smid[blockIdx.x] = static_cast<int>(sm);
        long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_cta_t[2 * blockIdx.x] = t;
    }
#endif
    if (tid == 0) { mbar_init(barIn, 1); mbar_init(barW, 1); fence_mbar_init(); }
    // zero the feature tiles once so rows beyond the active ones never hold junk
    for (int e = tid; e < NMAX * XS; e += kThreads) { sX[e] = 0.0f; sCat[e] = 0.0f; }
    for (int e = tid; e < 2 * NMAX; e += kThreads) sXX[e] = 0.0f;
    __syncthreads();

    const uint32_t inBytes = A.compact ? static_cast<uint32_t>(A.compact) : static_cast<uint32_t>(kInCh * N * 4);

    __shared__ int sSlot;
#pragma unroll 1
    for (int it = blockIdx.x;; it += gridDim.x) {
        int slot = it;
        if (A.work_ctr) {                               // dynamic: next heaviest unprocessed graph
            if (tid == 0) sSlot = atomicAdd(A.work_ctr, 1);
            __syncthreads();
            slot = sSlot;
            __syncthreads();
        }
        if (slot >= (SPLIT ? 2 * A.G : A.G)) break;
        const int br = SPLIT ? (slot & 1) : -1;         // -1: the whole graph; 0: xyz branch; 1: semantic branch
        if (SPLIT) slot >>= 1;
        const int g = A.order ? __ldg(A.order + slot) : slot;
        const float* gin = reinterpret_cast<const float*>(
            reinterpret_cast<const unsigned char*>((A.pairs && (g & 1)) ? A.g1 : A.g0) + static_cast<size_t>(A.pairs ? (g >> 1) : g) * inBytes);
        const bool bulk_ok = ((inBytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(gin) & 15u) == 0);
        uint8_t* tk = A.trace_knn ? A.trace_knn + static_cast<size_t>(g) * 6 * N * k : nullptr;
        float* tl = A.trace_layers ? A.trace_layers + static_cast<size_t>(g) * 6 * N * 64 : nullptr;

        SGPR_TL(0);
        // ---- stage the input block and the first GEMM's weights (TMA bulk copies, mbarrier completion) ----
        if (tid == 0) {
            if (bulk_ok) { mbar_expect_tx(barIn, inBytes); bulk_g2s(sIn, gin, inBytes, barIn); }
            if (br == 1) { mbar_expect_tx(barW, 12 * 128 * 4); bulk_g2s(sW, W.w_f1, 12 * 128 * 4, barW); }
            else         { mbar_expect_tx(barW, 64 * 128 * 4); bulk_g2s(sW, W.w_s2, 64 * 128 * 4, barW); }
        }
        if (bulk_ok) { mbar_wait(barIn, phIn); phIn ^= 1; }
        else { for (int e = tid; e < static_cast<int>(inBytes / 4); e += kThreads) sIn[e] = __ldg(gin + e); __syncthreads(); }
        if (A.compact) {
            // expand the record in place into the [15][N] block the rest of the kernel reads: rows 0-2 (xyz) already sit
            // where they belong, the N label bytes behind them become the twelve one-hot rows (sg_net.py:270-286)
            const int lab = (tid < N) ? reinterpret_cast<const uint8_t*>(sIn)[12 * N + tid] : 255;
            __syncthreads();
            if (tid < N) {
#pragma unroll
                for (int c = 0; c < kLabels; ++c) sIn[(3 + c) * N + tid] = (lab == c) ? 1.0f : 0.0f;
            }
            __syncthreads();
        }

        SGPR_TL(1);
        // ---- layer-0 tile (x, y, z, 0) + squared norms for every node, and the last non-zero node ----
        int last = -1;
        for (int n = tid; n < N; n += kThreads) {
            const float x = sIn[n], y = sIn[N + n], z = sIn[2 * N + n];
            *reinterpret_cast<float4*>(sCat + n * XS) = make_float4(x, y, z, 0.0f);
            sXX0[n] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
            uint32_t bits = 0;
#pragma unroll
            for (int c = 0; c < kInCh; ++c) bits |= __float_as_uint(sIn[c * N + n]);
            if ((bits << 1) != 0u) last = n;                  // +0.0 and -0.0 are both "zero": they compare and add alike
        }
        last = __reduce_max_sync(0xffffffffu, last);
        if (lane == 0) sLast[warp] = last;
        __syncthreads();
        int R = N;
        if (A.dedup) {
            int m = -1;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) m = max(m, sLast[w]);
            R = min(m + 2, N);      // nodes up to the last non-zero one, plus one representative of the trailing zero pads
        }
#ifdef SGPR_TIMELINE
        if (tid == 0 && blockIdx.x < 1024) g_cta_g[blockIdx.x] = g | (R << 16);
#endif
        // rows owned by this warp for the whole graph
        const int rpw = (R + kWarps - 1) / kWarps;
        const int w0 = min(R, warp * rpw), w1 = min(R, w0 + rpw);

        // ================= the six EdgeConv layers: xyz 1,2,3 then sem 1,2,3 (sg_net.py:84-102) =================
        // the semantic branch's input: rows [n][0..11] from input rows 3..14 (sg_net.py:82,94), own rows
        auto stage_semantic = [&]() {
            for (int i = w0; i < w1; ++i)
                if (lane < 16) sX[i * XS + lane] = (lane < kLabels) ? sIn[(3 + lane) * N + i] : 0.0f;
            __syncwarp();
            norms_rows(sX, sXX, 3, w0, w1, lane);
        };
        if (br == 1) { stage_semantic(); __syncthreads(); }
#pragma unroll 1
        for (int l = (br == 1) ? 3 : 0; l < ((br == 0) ? 3 : 6); ++l) {
            LayerDesc D;
            switch (l) {
                case 0:  D = LayerDesc{nullptr, nullptr, 0, 1, 64}; break;
                case 1:  D = LayerDesc{W.ab_s2, W.w_s3, 64 * 64 * 4, 16, 64}; break;
                case 2:  D = LayerDesc{W.ab_s3, (br == 0) ? nullptr : W.w_f1, 12 * 128 * 4, 16, 32}; break;
                case 3:  D = LayerDesc{W.ab_f1, W.w_f2, 64 * 128 * 4, 3, 64}; break;
                case 4:  D = LayerDesc{W.ab_f2, W.w_f3, (64 * 64 + 64 * 32) * 4, 16, 64}; break;   // w_f3 + w_end (adjacent)
                default: D = LayerDesc{W.ab_f3, nullptr, 0, 16, 32}; break;
            }
            uint8_t* tkl = tk ? tk + l * N * k : nullptr;
            float* tr = tl ? tl + l * N * 64 : nullptr;

            // ---- front: distance rows -> selection -> (GEMM rows) for own rows, no CTA barrier ----
            FrontCtx F{(l == 0) ? sCat : sX, (l == 0) ? sXX0 : sXX, sY, sW, sIdx, sCnt, tkl, D.cin4, D.cout, R, N, k, KS, l};
            bool waited = false;
            SGPR_TL(8 + l * 8 + 0);
            for (int r0 = w0; r0 < w1; r0 += 8) front_pass<NPL, TIES>(F, r0, min(8, w1 - r0), lane, barW, phW, waited);
            SGPR_TL(8 + l * 8 + 3);
            if (l != 0 && !waited) { mbar_wait(barW, phW); phW ^= 1; }       // warps without rows still track the phase

            if (l == 0) {
                // xyz layer 1 gathers from the input block itself (read-only): no barrier needed before it
                xyz_rows(sCat, sIdx, sCnt, KS, W.s1, sX, tr, w0, w1, lane);
                __syncwarp();
                norms_rows(sX, sXX, 16, w0, w1, lane);
            } else {
                __syncthreads();                               // barrier B: every A|B row is in place, sW is consumed
                SGPR_TL(8 + l * 8 + 4);
                if (tid == 0 && D.next_w) {
                    // an xyz-branch unit may be the one that merges: conv_end's matrix rides along with layer 3's
                    const int extra = (br == 0 && l == 1) ? 64 * 32 * 4 : 0;
                    mbar_expect_tx(barW, D.next_bytes + extra);
                    bulk_g2s(sW, D.next_w, D.next_bytes, barW);
                    if (extra) bulk_g2s(sW + 64 * 64, W.w_end, extra, barW);
                }
                // ---- back: gather-max for own rows ----
                if (D.cout == 64) {
                    gather_rows<64>(sY, sIdx, sCnt, KS, D.ab, sX, tr, w0, w1, lane);
                    __syncwarp();
                    norms_rows(sX, sXX, 16, w0, w1, lane);
                } else {
                    gather_rows<32>(sY, sIdx, sCnt, KS, D.ab, (l == 2) ? sCat : sCat + 32, tr, w0, w1, lane);
                    if (l == 2) {
                        if (br < 0) stage_semantic();
                    } else if (br < 0) {   // l == 5: conv_end on own rows (sg_net.py:104-109): cat(xyz3, sem3) [.,64] -> [.,32]
                        __syncwarp();
                        for (int r0 = w0; r0 < w1; r0 += 8)
                            conv_end_dispatch(sCat, sW + 64 * 64, sX, W.ab_end, r0, min(8, w1 - r0), lane);
                    }
                }
            }
            SGPR_TL(8 + l * 8 + 5);
            __syncthreads();                                   // barrier A: the next layer's input (or sE) is complete
            SGPR_TL(8 + l * 8 + 6);
            if (tr || tkl) {                                   // debug taps: every trailing pad is a copy of row R-1
                if (tr) for (int e = tid; e < (N - R) * 64; e += kThreads) tr[(R + e / 64) * 64 + (e & 63)] = tr[(R - 1) * 64 + (e & 63)];
                if (tkl) for (int e = tid; e < (N - R) * k; e += kThreads) tkl[(R + e / k) * k + (e % k)] = tkl[(R - 1) * k + (e % k)];
            }
        }

        if constexpr (SPLIT != 0) {
            // ---- branch unit: hand the 32 channels over; whichever unit of the graph arrives second merges ----
            float* mine = A.halves + (static_cast<size_t>(g) * 2 + br) * N * kF3;
            for (int e = tid; e < R * kF3; e += kThreads) mine[e] = sCat[(e >> 5) * XS + br * 32 + (e & 31)];
            __threadfence();
            __syncthreads();
            if (tid == 0) sFlag = atomicAdd(A.gctr + g, 1);
            __syncthreads();
            if (sFlag == 0) { __syncthreads(); continue; }
            __threadfence();
            const float* other = A.halves + (static_cast<size_t>(g) * 2 + (br ^ 1)) * N * kF3;
            for (int e = tid; e < R * kF3; e += kThreads) sCat[(e >> 5) * XS + (br ^ 1) * 32 + (e & 31)] = __ldcg(other + e);
            if (tid == 0) A.gctr[g] = 0;
            __syncthreads();
            for (int r0 = w0; r0 < w1; r0 += 8) conv_end_dispatch(sCat, sW + 64 * 64, sX, W.ab_end, r0, min(8, w1 - r0), lane);
            __syncthreads();
        }
        SGPR_TL(58);
        finish_graph(A, W, H, sX, sRed, sY, &sFlag, g, N, R, tid, warp, lane);
        __syncthreads();
        SGPR_TL(62);
    }
#ifdef SGPR_TIMELINE
    if (tid == 0 && blockIdx.x < 1024) {
        long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_cta_t[2 * blockIdx.x + 1] = t;
    }
#endif
}

}  // namespace sgpr
