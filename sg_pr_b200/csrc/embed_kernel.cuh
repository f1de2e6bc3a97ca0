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
// so each layer is: k-NN (Gram tile -> per-row sorting network -> threshold select) -> per-node GEMM
// [R x C]x[C x 2C'] -> gather-max over the neighbour rows of A.  Layer 1 of the xyz branch keeps the reference's
// direct form W_a (x_j - x_i): metre-scale coordinates would lose ~5 bits to cancellation in A_j - A_i.
//
// Zero padding (sg_net.py:258-262, 276-278: graphs are padded to node_num with all-zero nodes): trailing all-zero
// nodes have bit-identical inputs, hence bit-identical features, distances and neighbour sets in every layer.  The
// kernel works on R = (#nodes up to the last non-zero one) + 1 rows — the first pad stands for its whole class, with
// its multiplicity kept in the k-NN selection — and replicates that row before the attention stage.  Results are
// bit-identical to processing every pad; graphs without trailing zero nodes simply have R = N.
//
// The kernel is issue/latency-bound (~1.6 kflop per input byte), so the design minimises instruction count:
// one copy of the phase code in a runtime layer loop (instruction cache), packed fp32 FMA (fma.rn.f32x2, SASS
// FFMA2) for every dot product with the even/odd-channel partial sums in the two halves, and an in-register
// Batcher network for the k-th-largest threshold instead of a shuffle-heavy warp sort.
#pragma once
#include "common.cuh"

namespace sgpr {

struct EmbedArgs {
    const float* g0;        // graphs of side 0 (or all graphs when !pairs)   [*, 15, N]
    const float* g1;        // graphs of side 1 (pairs mode)
    int G;                  // number of graphs to embed (2*B in pairs mode: g = 2*b + side)
    int N, k, KS;           // KS = k rounded up to 4 (neighbour-list row stride in bytes)
    int pairs;
    int dedup;              // 1: collapse trailing all-zero nodes into one row (exact); 0: process every node
    float* pooled;          // [G][32]
    float* att0;            // pairs: [B][N] side 0 ; else [G][N]   (may be null)
    float* att1;            // pairs: [B][N] side 1                 (may be null)
    float* emb;             // [G][N][32] or null
    float* score;           // [B]  (pairs)
    int* counters;          // [B]  arrival counters, zero on entry, zero on exit
    uint8_t* trace_knn;     // [G][6][N][k] or null
    float* trace_layers;    // [G][6][N][64] or null
};

struct SmemLayout {
    int w, in, x, y, cat, xx, red, bar, idx, cnt, total;   // byte offsets
};

__host__ __device__ inline SmemLayout make_layout(int nmax, int ks) {
    SmemLayout L;
    int o = 0;
    L.w = o;   o += 64 * 128 * 4;              // largest packed layer matrix (64 in x 128 out)
    L.in = o;  o += ((kInCh * nmax * 4 + 15) / 16) * 16;
    L.x = o;   o += nmax * XS * 4;
    L.y = o;   o += nmax * YS * 4;             // A|B tile; doubles as the distance tile during the k-NN phase
    L.cat = o; o += nmax * XS * 4;
    L.xx = o;  o += nmax * 4;                  // squared norms, then per-row thresholds, then attention scores
    L.red = o; o += (kWarps * 32 + 64) * 4;
    L.bar = o; o += 16;
    L.idx = o; o += ((nmax * ks + 15) / 16) * 16;
    L.cnt = o; o += ((nmax + 15) / 16) * 16;
    L.total = o;
    return L;
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ void cmpx(float& a, float& b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}

// ------------------------------------------------------------------------------------------------------------
// Distance tile: pd[i][c] = (2*dot(x_i,x_c) - xx_c) - xx_i  ==  -xx - inner - xx^T with inner = -2*dot, same
// rounding order as dgcnn.py:15-17.  Rows and columns < R.  Stored as sP[i*PS + (c>>5)*33 + (c&31)], PS = 33*NPL
// (the +1 skew per 32-column block keeps both the lane<->column stores here and the lane<->row loads of the
// selection phase conflict-free).  A warp owns RG rows x all columns at a time.
// ------------------------------------------------------------------------------------------------------------
template <int NPL, int RG>
__device__ __forceinline__ void knn_gram(const float* __restrict__ sX, const float* __restrict__ sXX,
                                         float* __restrict__ sP, int c4n, int R, int warp, int lane) {
    constexpr int PS = 33 * NPL;
    const int ngroups = (R + RG - 1) / RG;
    const int nq = (R + 31) >> 5;
#pragma unroll 1
    for (int g = warp; g < ngroups; g += kWarps) {
        const int i0 = g * RG;
        float2 acc[RG][NPL];
#pragma unroll
        for (int r = 0; r < RG; ++r)
#pragma unroll
            for (int q = 0; q < NPL; ++q) acc[r][q] = make_float2(0.0f, 0.0f);
        const float* pa = sX + i0 * XS;
        const float* pb = sX + lane * XS;
#pragma unroll 2
        for (int c = 0; c < c4n; ++c) {
            float4 a[RG];
#pragma unroll
            for (int r = 0; r < RG; ++r) a[r] = *reinterpret_cast<const float4*>(pa + r * XS + 4 * c);
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                if (q < nq) {
                    const float4 b = *reinterpret_cast<const float4*>(pb + 32 * q * XS + 4 * c);
#pragma unroll
                    for (int r = 0; r < RG; ++r) {
                        acc[r][q] = ffma2(make_float2(a[r].x, a[r].y), make_float2(b.x, b.y), acc[r][q]);
                        acc[r][q] = ffma2(make_float2(a[r].z, a[r].w), make_float2(b.z, b.w), acc[r][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int c = lane + 32 * q;
            if (c < R) {
                const float xxc = sXX[c];
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    const int i = i0 + r;
                    if (i < R) {
                        const float dot = __fadd_rn(acc[r][q].x, acc[r][q].y);
                        const float t = __fsub_rn(__fmul_rn(2.0f, dot), xxc);
                        sP[i * PS + q * 33 + lane] = __fsub_rn(t, sXX[i]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k-NN selection (dgcnn.py:19 `topk`): for every row < R pick the k largest of its N distances, ties at the k-th
// value to the lowest column index.  Columns >= R that are < N all carry the pad-class value pd[i][R-1]; columns
// >= N are -inf.  Only the SET matters downstream (max over neighbours), so the list is emitted in ascending
// column order with the pad class, if selected, represented once by column R-1, then padded to a multiple of 4.
//
// NPL lanes cooperate on a row, each holding 32 consecutive columns in registers: in-register Batcher sort of 32,
// log2(NPL) cross-lane bitonic merges, threshold = element NMAX-k of the sorted row, then count / emit passes over
// the unsorted values.  Per-row thresholds go to sThr (for the debug trace).
// ------------------------------------------------------------------------------------------------------------
template <int NPL>
__device__ __forceinline__ void knn_select(const float* __restrict__ sP, uint8_t* __restrict__ sIdx,
                                           uint8_t* __restrict__ sCnt, float* __restrict__ sThr, int R, int N, int k,
                                           int KS, int warp, int lane) {
    constexpr int PS = 33 * NPL;
    constexpr int NMAX = 32 * NPL;
    constexpr int RPW = 32 / NPL;                     // rows per warp task
    const int sub = lane / RPW;                       // which 32-column block of the row this lane holds
    const int rl = lane % RPW;
    const int ntasks = (R + RPW - 1) / RPW;
#pragma unroll 1
    for (int task = warp; task < ntasks; task += kWarps) {
        const int row = task * RPW + rl;
        const int rowc = min(row, R - 1);             // inactive lanes shadow the last row (never store)
        const float* prow = sP + rowc * PS;
        const float padval = prow[((R - 1) >> 5) * 33 + ((R - 1) & 31)];
        float o[32], v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int c = sub * 32 + j;
            const float x = (c < R) ? prow[sub * 33 + j] : ((c < N) ? padval : -INFINITY);
            o[j] = x;
            v[j] = x;
        }
        // ---- sort the lane's 32 values ascending ----
#define SGPR_CX(i, j) cmpx(v[i], v[j]);
#include "sortnet32.inc"
#undef SGPR_CX
        // ---- merge across the NPL lanes of the row (bitonic "flip" merges) ----
#pragma unroll
        for (int m = 1; m < NPL; m <<= 1) {           // m = number of lanes per sorted run being merged
            {   // mirror exchange with lane sub ^ (2m-1): element j meets the partner's element 31-j
                const bool keep_min = (sub & m) == 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float pa = __shfl_xor_sync(0xffffffffu, v[31 - j], (2 * m - 1) * RPW);
                    const float pb = __shfl_xor_sync(0xffffffffu, v[j], (2 * m - 1) * RPW);
                    v[j] = keep_min ? fminf(v[j], pa) : fmaxf(v[j], pa);
                    v[31 - j] = keep_min ? fminf(v[31 - j], pb) : fmaxf(v[31 - j], pb);
                }
            }
#pragma unroll
            for (int dl = m >> 1; dl >= 1; dl >>= 1) {  // cross-lane half-cleaners
                const bool keep_min = (sub & dl) == 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float p = __shfl_xor_sync(0xffffffffu, v[j], dl * RPW);
                    v[j] = keep_min ? fminf(v[j], p) : fmaxf(v[j], p);
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1)          // in-register half-cleaners
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if ((j & d) == 0) cmpx(v[j], v[j | d]);
        }
        // ---- threshold: k-th largest = sorted position NMAX - k ----
        const int P = NMAX - k;
        float tl;
        {   // v[P & 31] with a warp-uniform index: 31 selects instead of a local-memory array
            const int jt = P & 31;
            float s16[16], s8[8], s4[4], s2[2];
#pragma unroll
            for (int i = 0; i < 16; ++i) s16[i] = (jt & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
            for (int i = 0; i < 8; ++i) s8[i] = (jt & 2) ? s16[2 * i + 1] : s16[2 * i];
#pragma unroll
            for (int i = 0; i < 4; ++i) s4[i] = (jt & 4) ? s8[2 * i + 1] : s8[2 * i];
#pragma unroll
            for (int i = 0; i < 2; ++i) s2[i] = (jt & 8) ? s4[2 * i + 1] : s4[2 * i];
            tl = (jt & 16) ? s2[1] : s2[0];
        }
        const float thr = __shfl_sync(0xffffffffu, tl, rl + (P >> 5) * RPW);
        // ---- count pass over the unsorted values ----
        int gt_all = 0, gt_e = 0, eq_e = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int c = sub * 32 + j;
            const bool g = o[j] > thr, e = o[j] == thr, em = c < R;
            gt_all += g;
            gt_e += (g && em);
            eq_e += (e && em);
        }
        int gt_tot = gt_all, eq_before = 0;
#pragma unroll
        for (int s = 1; s < NPL; ++s) {
            const int src = rl + ((sub + s) % NPL) * RPW;
            gt_tot += __shfl_sync(0xffffffffu, gt_all, src);
        }
        const int need = k - gt_tot;                  // tied-at-threshold elements to take, lowest columns first (>= 1)
        int eqs[NPL], gts[NPL];
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            eqs[s] = __shfl_sync(0xffffffffu, eq_e, rl + s * RPW);
            gts[s] = __shfl_sync(0xffffffffu, gt_e, rl + s * RPW);
        }
        int base = 0, total = 0;
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            int eb = 0;
#pragma unroll
            for (int s2 = 0; s2 < s; ++s2) eb += eqs[s2];
            const int take = gts[s] + max(0, min(eqs[s], need - eb));
            if (s < sub) base += take;
            if (s == sub) eq_before = eb;
            total += take;
        }
        // ---- emit pass ----
        if (row < R) {
            uint8_t* list = sIdx + row * KS;
            int pos = base, eq_left = need - eq_before;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int c = sub * 32 + j;
                if (c < R) {
                    const bool tk = (o[j] > thr) || (o[j] == thr && eq_left > 0);
                    if (o[j] == thr) --eq_left;
                    if (tk) { list[pos] = static_cast<uint8_t>(c); ++pos; }
                }
            }
            // the lane that wrote the last entry pads the list to a multiple of 4 (repeats are harmless under max)
            if (pos == total && pos > base) {
                const uint8_t last = list[pos - 1];
                for (int e = total; e < ((total + 3) & ~3); ++e) list[e] = last;
            }
            if (sub == 0) { sCnt[row] = static_cast<uint8_t>((total + 3) & ~3); sThr[row] = thr; }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Per-node GEMM: out[n][co] = sum_ci X[n][ci] * W[ci][co], co in [0, 32*CPL), rows n < R.
// A warp owns NT nodes x all outputs; a lane owns CPL consecutive outputs.  W is packed in channel PAIRS
// (pack.hpp::pair_index): row p holds, for every output, (W[2p][co], W[2p+1][co]) so that one FFMA2 advances the
// even- and odd-channel partial sums of an output.  EPI: 0 = store raw, 1 = BN(alpha,beta)+LeakyReLU (conv_end).
// ------------------------------------------------------------------------------------------------------------
template <int NT, int CPL, int EPI>
__device__ __forceinline__ void node_gemm_nt(const float* __restrict__ sXin, const float* __restrict__ sW,
                                             float* __restrict__ sOut, int outStride, const float* __restrict__ ab,
                                             int cin4, int R, int warp, int lane) {
    constexpr int CO = 32 * CPL;
    constexpr int ROW = 2 * CO;                  // floats per channel-pair row
    const int nchunks = (R + NT - 1) / NT;
#pragma unroll 1
    for (int ch = warp; ch < nchunks; ch += kWarps) {
        const int n0 = ch * NT;
        float2 acc[NT][CPL];
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[n][j] = make_float2(0.0f, 0.0f);

        const float* px = sXin + n0 * XS;
#pragma unroll 2
        for (int c4 = 0; c4 < cin4; ++c4) {
            float4 x[NT];
#pragma unroll
            for (int n = 0; n < NT; ++n) x[n] = *reinterpret_cast<const float4*>(px + n * XS + 4 * c4);
#pragma unroll
            for (int h = 0; h < 2; ++h) {        // channel pair 2*c4 + h
                float2 w[CPL];
                const float* wp = sW + (2 * c4 + h) * ROW;
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
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const float2 xv = h ? make_float2(x[n].z, x[n].w) : make_float2(x[n].x, x[n].y);
#pragma unroll
                    for (int j = 0; j < CPL; ++j) acc[n][j] = ffma2(xv, w[j], acc[n][j]);
                }
            }
        }

        float al[CPL], be[CPL];
        if constexpr (EPI == 1) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) { al[j] = __ldg(ab + lane * CPL + j); be[j] = __ldg(ab + CO + lane * CPL + j); }
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            float y[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                y[j] = __fadd_rn(acc[n][j].x, acc[n][j].y);
                if constexpr (EPI == 1) y[j] = lrelu(fmaf(y[j], al[j], be[j]));
            }
            float* op = sOut + (n0 + n) * outStride + lane * CPL;
            if constexpr (CPL == 4) *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
            else if constexpr (CPL == 2) *reinterpret_cast<float2*>(op) = make_float2(y[0], y[1]);
            else *op = y[0];
        }
    }
}

// rows per warp task chosen so that ceil(R/NT) fills the 8 warps as evenly as the three variants allow
template <int CPL, int EPI>
__device__ __forceinline__ void node_gemm(const float* __restrict__ sXin, const float* __restrict__ sW,
                                          float* __restrict__ sOut, int outStride, const float* __restrict__ ab,
                                          int cin4, int R, int warp, int lane) {
    const int per = (R + kWarps - 1) / kWarps;
    if (per <= 4) node_gemm_nt<4, CPL, EPI>(sXin, sW, sOut, outStride, ab, cin4, R, warp, lane);
    else if (per <= 6) node_gemm_nt<6, CPL, EPI>(sXin, sW, sOut, outStride, ab, cin4, R, warp, lane);
    else node_gemm_nt<8, CPL, EPI>(sXin, sW, sOut, outStride, ab, cin4, R, warp, lane);
}

// ------------------------------------------------------------------------------------------------------------
// Gather-max + BN + LeakyReLU (sg_net.py:85-86 etc.): for node i < R and channel c
//     out = LReLU(alpha_c * ((max_{j in knn(i)} A[j][c] - A[i][c]) + B[i][c]) + beta_c)
// sY row = [A(0..COUT) | B(COUT..2COUT)].  A warp owns a node, a lane owns COUT/32 channels.
// ------------------------------------------------------------------------------------------------------------
template <int COUT>
__device__ __forceinline__ void gather_max_bn(const float* __restrict__ sY, const uint8_t* __restrict__ sIdx,
                                              const uint8_t* __restrict__ sCnt, int KS, const float* __restrict__ ab,
                                              float* __restrict__ sDst, int dstStride, float* __restrict__ trace, int R,
                                              int warp, int lane) {
    constexpr int CPL = COUT / 32;
    float al[CPL], be[CPL];
#pragma unroll
    for (int p = 0; p < CPL; ++p) { al[p] = __ldg(ab + lane * CPL + p); be[p] = __ldg(ab + COUT + lane * CPL + p); }

#pragma unroll 1
    for (int i = warp; i < R; i += kWarps) {
        float m[CPL];
#pragma unroll
        for (int p = 0; p < CPL; ++p) m[p] = -INFINITY;
        const uint8_t* row = sIdx + i * KS;
        const int cnt = sCnt[i];
        const float* base = sY + lane * CPL;
#pragma unroll 2
        for (int t = 0; t < cnt; t += 4) {
            const uchar4 jj = *reinterpret_cast<const uchar4*>(row + t);
            if constexpr (CPL == 2) {
                const float2 a0 = *reinterpret_cast<const float2*>(base + jj.x * YS);
                const float2 a1 = *reinterpret_cast<const float2*>(base + jj.y * YS);
                const float2 a2 = *reinterpret_cast<const float2*>(base + jj.z * YS);
                const float2 a3 = *reinterpret_cast<const float2*>(base + jj.w * YS);
                m[0] = fmaxf(fmaxf(m[0], fmaxf(a0.x, a1.x)), fmaxf(a2.x, a3.x));
                m[1] = fmaxf(fmaxf(m[1], fmaxf(a0.y, a1.y)), fmaxf(a2.y, a3.y));
            } else {
                const float a0 = base[jj.x * YS], a1 = base[jj.y * YS];
                const float a2 = base[jj.z * YS], a3 = base[jj.w * YS];
                m[0] = fmaxf(fmaxf(m[0], fmaxf(a0, a1)), fmaxf(a2, a3));
            }
        }
#pragma unroll
        for (int p = 0; p < CPL; ++p) {
            const float ai = base[i * YS + p];
            const float bi = base[i * YS + COUT + p];
            const float y = __fadd_rn(__fsub_rn(m[p], ai), bi);
            const float z = lrelu(fmaf(y, al[p], be[p]));
            sDst[i * dstStride + lane * CPL + p] = z;
            if (trace) trace[i * 64 + lane * CPL + p] = z;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// xyz layer 1 (3 -> 64) in the reference's direct form (sg_net.py:84-86): per edge
//     e = wa0*d0 + wa1*d1 + wa2*d2  (d = x_j - x_i, sequential FMA), max over the edges, then the centre
//     terms wb.x_i appended in the same sequential order (monotone in e, so they commute with the max).
// sIn is the channel-major input block [15][N]; a lane owns output channels 2*lane, 2*lane+1.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void xyz_layer1(const float* __restrict__ sIn, const uint8_t* __restrict__ sIdx,
                                           const uint8_t* __restrict__ sCnt, int KS, const float* __restrict__ s1,
                                           float* __restrict__ sDst, float* __restrict__ trace, int N, int R, int warp,
                                           int lane) {
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane) * 2);
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane) * 2 + 1);
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane + 1) * 2);
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane + 1) * 2 + 1);
    // p0 = {wa0,wa1,wa2,wb0}, p1 = {wb1,wb2,alpha,beta} of channel 2*lane; r0/r1 the same for 2*lane+1
    const float2 wa0 = make_float2(p0.x, r0.x), wa1 = make_float2(p0.y, r0.y), wa2 = make_float2(p0.z, r0.z);
#pragma unroll 1
    for (int i = warp; i < R; i += kWarps) {
        const float xi0 = sIn[i], xi1 = sIn[N + i], xi2 = sIn[2 * N + i];
        float m0 = -INFINITY, m1 = -INFINITY;
        const uint8_t* row = sIdx + i * KS;
        const int cnt = sCnt[i];
#pragma unroll 1
        for (int t = 0; t < cnt; t += 4) {
            const uchar4 jj = *reinterpret_cast<const uchar4*>(row + t);
            const int js[4] = {jj.x, jj.y, jj.z, jj.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = js[u];
                const float d0 = __fsub_rn(sIn[j], xi0);
                const float d1 = __fsub_rn(sIn[N + j], xi1);
                const float d2 = __fsub_rn(sIn[2 * N + j], xi2);
                float2 e = __fmul2_rn(wa0, make_float2(d0, d0));
                e = ffma2(wa1, make_float2(d1, d1), e);
                e = ffma2(wa2, make_float2(d2, d2), e);
                m0 = fmaxf(m0, e.x);
                m1 = fmaxf(m1, e.y);
            }
        }
        float y0 = fmaf(p0.w, xi0, m0); y0 = fmaf(p1.x, xi1, y0); y0 = fmaf(p1.y, xi2, y0);
        float y1 = fmaf(r0.w, xi0, m1); y1 = fmaf(r1.x, xi1, y1); y1 = fmaf(r1.y, xi2, y1);
        const float z0 = lrelu(fmaf(y0, p1.z, p1.w));
        const float z1 = lrelu(fmaf(y1, r1.z, r1.w));
        *reinterpret_cast<float2*>(sDst + i * XS + 2 * lane) = make_float2(z0, z1);
        if (trace) { trace[i * 64 + 2 * lane] = z0; trace[i * 64 + 2 * lane + 1] = z1; }
    }
}

// squared norms per node: xx = sum_c x_c^2 (dgcnn.py:16; products rounded, then added).  4 threads per node, each
// sums a quarter of the channels in order, partials combined pairwise.
__device__ __forceinline__ void sq_norms(const float* __restrict__ sX, float* __restrict__ sXX, int c4n, int R, int tid) {
    const int part = tid & 3;
    const int per = (c4n + 3) >> 2;                     // float4 groups per quarter
    for (int n0 = 0; n0 < R; n0 += kThreads / 4) {      // warp-uniform trip count
        const int n = n0 + (tid >> 2);
        float s = 0.0f;
        if (n < R) {
            for (int g = part * per; g < min(c4n, (part + 1) * per); ++g) {
                const float4 x = *reinterpret_cast<const float4*>(sX + n * XS + 4 * g);
                s = __fadd_rn(s, __fmul_rn(x.x, x.x));
                s = __fadd_rn(s, __fmul_rn(x.y, x.y));
                s = __fadd_rn(s, __fmul_rn(x.z, x.z));
                s = __fadd_rn(s, __fmul_rn(x.w, x.w));
            }
        }
        s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
        s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
        if (n < R && part == 0) sXX[n] = s;
    }
}

// Debug taps (parity tests only): expand the per-row selection into the k column indices the undeduplicated kernel
// would list (pads expanded, lowest index first), one thread per row, from the distance tile and the thresholds.
template <int NPL>
__device__ __forceinline__ void trace_knn_rows(uint8_t* __restrict__ dst, const float* __restrict__ sP,
                                               const float* __restrict__ sThr, int N, int R, int k, int tid) {
    if (!dst) return;
    constexpr int PS = 33 * NPL;
    for (int row = tid; row < N; row += kThreads) {
        const int rr = min(row, R - 1);
        const float* prow = sP + rr * PS;
        const float thr = sThr[rr];
        const float padval = prow[((R - 1) >> 5) * 33 + ((R - 1) & 31)];
        int ngt = 0;
        for (int c = 0; c < N; ++c) {
            const float v = (c < R) ? prow[(c >> 5) * 33 + (c & 31)] : padval;
            ngt += v > thr;
        }
        int need = k - ngt, pos = 0;
        for (int c = 0; c < N && pos < k; ++c) {
            const float v = (c < R) ? prow[(c >> 5) * 33 + (c & 31)] : padval;
            if (v > thr || (v == thr && need-- > 0)) dst[row * k + pos++] = static_cast<uint8_t>(c);
        }
    }
}

__device__ __forceinline__ void trace_replicate_rows(float* __restrict__ trace, int N, int R, int tid) {
    if (!trace) return;
    for (int e = tid; e < (N - R) * 64; e += kThreads) trace[(R + e / 64) * 64 + (e & 63)] = trace[(R - 1) * 64 + (e & 63)];
}

// ------------------------------------------------------------------------------------------------------------
// Pair head: NTN + FC + sigmoid for one ordered pair (layers_batch.py:70-83, sg_net.py:131-136).
// e1/e2: 32 pooled floats each (shared memory). scratch: >= 512+64 floats of shared memory.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_acc(float z) { return 1.0f / (1.0f + expf(-z)); }

__device__ __forceinline__ void pair_head_cta(const float* __restrict__ e1, const float* __restrict__ e2,
                                              const PackedWeights& W, const HeadParams& H, float* __restrict__ scratch,
                                              float* __restrict__ score_out, int tid) {
    float* P = scratch;          // [512]  P[b*16+t] = sum_a e1[a] * W[a][b*16+t]   (layers_batch.py:78)
    float* s = scratch + 512;    // [16]
    float* h = scratch + 528;    // [16]
    for (int c = tid; c < 512; c += kThreads) {
        float acc = 0.0f;
#pragma unroll 8
        for (int a = 0; a < kF3; ++a) acc = fmaf(e1[a], __ldg(W.ntn_w + a * 512 + c), acc);
        P[c] = acc;
    }
    __syncthreads();
    if (tid < kT) {
        float acc = 0.0f;
        for (int b = 0; b < kF3; ++b) acc = fmaf(P[b * kT + tid], e2[b], acc);      // layers_batch.py:79
        float blk = 0.0f;
        for (int c = 0; c < kF3; ++c) blk = fmaf(__ldg(W.ntn_v + tid * 64 + c), e1[c], blk);          // :80-81
        for (int c = 0; c < kF3; ++c) blk = fmaf(__ldg(W.ntn_v + tid * 64 + 32 + c), e2[c], blk);
        s[tid] = fmaxf(__fadd_rn(__fadd_rn(acc, blk), __ldg(W.ntn_b + tid)), 0.0f);                 // :82
    }
    __syncthreads();
    if (tid < kBn) {
        float acc = 0.0f;
        for (int t = 0; t < kT; ++t) acc = fmaf(s[t], H.fc1_w[tid * kT + t], acc);                   // sg_net.py:134
        h[tid] = fmaxf(__fadd_rn(acc, H.fc1_b[tid]), 0.0f);
    }
    __syncthreads();
    if (tid == 0) {
        float acc = 0.0f;
        for (int u = 0; u < kBn; ++u) acc = fmaf(h[u], H.fc2_w[u], acc);                             // sg_net.py:136
        *score_out = sigmoidf_acc(__fadd_rn(acc, H.fc2_b));
    }
}

// One entry of the layer loop (the five GEMM-form EdgeConv layers, sg_net.py:87-102).
struct LayerDesc {
    const float* ab;       // alpha | beta
    const float* next_w;   // matrix to prefetch into sW once this layer's GEMM has consumed sW
    int next_bytes;
    int cin4;              // input channels / 4
    int cout;              // 64 or 32
};

template <int NPL>
__device__ __forceinline__ void knn_gram_dispatch(const float* sX, const float* sXX, float* sP, int c4n, int R, int warp,
                                                  int lane) {
    const int per = (R + kWarps - 1) / kWarps;        // rows per warp if spread evenly
    if (NPL >= 4 || per <= 4) knn_gram<NPL, 4>(sX, sXX, sP, c4n, R, warp, lane);
    else if (per <= 6) knn_gram<NPL, 6>(sX, sXX, sP, c4n, R, warp, lane);
    else knn_gram<NPL, 8>(sX, sXX, sP, c4n, R, warp, lane);
}

// ------------------------------------------------------------------------------------------------------------
// The fused kernel.
// ------------------------------------------------------------------------------------------------------------
#ifndef SGPR_MINBLOCKS_SMALL
#define SGPR_MINBLOCKS_SMALL 2
#endif
template <int NPL>
__global__ void __launch_bounds__(kThreads, (NPL <= 2) ? SGPR_MINBLOCKS_SMALL : 1)
sgpr_embed_kernel(const EmbedArgs A, const PackedWeights W, const HeadParams H) {
    constexpr int NMAX = 32 * NPL;
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L = make_layout(NMAX, A.KS);
    float* sW = reinterpret_cast<float*>(smem + L.w);
    float* sIn = reinterpret_cast<float*>(smem + L.in);
    float* sX = reinterpret_cast<float*>(smem + L.x);
    float* sY = reinterpret_cast<float*>(smem + L.y);
    float* sP = sY;                                   // distance tile lives in the A|B tile between GEMMs
    float* sCat = reinterpret_cast<float*>(smem + L.cat);
    float* sXX = reinterpret_cast<float*>(smem + L.xx);
    float* sRed = reinterpret_cast<float*>(smem + L.red);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar);
    uint8_t* sIdx = smem + L.idx;
    uint8_t* sCnt = smem + L.cnt;
    __shared__ int sFlag;
    __shared__ int sLast[kWarps];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = A.N, k = A.k, KS = A.KS;
    uint64_t* barIn = bars;
    uint64_t* barW = bars + 1;
    uint32_t phIn = 0, phW = 0;

    if (tid == 0) { mbar_init(barIn, 1); mbar_init(barW, 1); fence_mbar_init(); }
    // zero the feature tiles once so rows beyond the active ones never hold junk
    for (int e = tid; e < NMAX * XS; e += kThreads) { sX[e] = 0.0f; sCat[e] = 0.0f; }
    for (int e = tid; e < NMAX; e += kThreads) sXX[e] = 0.0f;
    __syncthreads();

    const uint32_t inBytes = static_cast<uint32_t>(kInCh * N * 4);

#pragma unroll 1
    for (int g = blockIdx.x; g < A.G; g += gridDim.x) {
        const float* gin = A.pairs ? (((g & 1) ? A.g1 : A.g0) + static_cast<size_t>(g >> 1) * kInCh * N)
                                   : (A.g0 + static_cast<size_t>(g) * kInCh * N);
        const bool bulk_ok = ((inBytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(gin) & 15u) == 0);
        uint8_t* tk = A.trace_knn ? A.trace_knn + static_cast<size_t>(g) * 6 * N * k : nullptr;
        float* tl = A.trace_layers ? A.trace_layers + static_cast<size_t>(g) * 6 * N * 64 : nullptr;

        // ---- stage the input block and the first GEMM's weights (TMA bulk copies, mbarrier completion) ----
        if (tid == 0) {
            if (bulk_ok) { mbar_expect_tx(barIn, inBytes); bulk_g2s(sIn, gin, inBytes, barIn); }
            mbar_expect_tx(barW, 64 * 128 * 4);
            bulk_g2s(sW, W.w_s2, 64 * 128 * 4, barW);
        }
        if (bulk_ok) { mbar_wait(barIn, phIn); phIn ^= 1; }
        else { for (int e = tid; e < kInCh * N; e += kThreads) sIn[e] = __ldg(gin + e); __syncthreads(); }

        // ---- active rows: nodes up to the last non-zero one, plus one representative of the trailing zero pads ----
        int R = N;
        if (A.dedup) {
            int last = -1;
            for (int n = tid; n < N; n += kThreads) {
                uint32_t bits = 0;
#pragma unroll
                for (int c = 0; c < kInCh; ++c) bits |= __float_as_uint(sIn[c * N + n]);
                if ((bits << 1) != 0u) last = n;          // +0.0 and -0.0 are both "zero": they compare and add alike
            }
            last = __reduce_max_sync(0xffffffffu, last);
            if (lane == 0) sLast[warp] = last;
            __syncthreads();
            int m = -1;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) m = max(m, sLast[w]);
            R = min(m + 2, N);
        }

        // ================= the six EdgeConv layers: xyz 1,2,3 then sem 1,2,3 (sg_net.py:84-102) =================
#pragma unroll 1
        for (int l = 0; l < 6; ++l) {
            LayerDesc D;
            switch (l) {
                case 0:  D = LayerDesc{nullptr, nullptr, 0, 1, 64}; break;
                case 1:  D = LayerDesc{W.ab_s2, W.w_s3, 64 * 64 * 4, 16, 64}; break;
                case 2:  D = LayerDesc{W.ab_s3, W.w_f1, 12 * 128 * 4, 16, 32}; break;
                case 3:  D = LayerDesc{W.ab_f1, W.w_f2, 64 * 128 * 4, 3, 64}; break;
                case 4:  D = LayerDesc{W.ab_f2, W.w_f3, 64 * 64 * 4, 16, 64}; break;
                default: D = LayerDesc{W.ab_f3, W.w_end, 64 * 32 * 4, 16, 32}; break;
            }
            // ---- this layer's input as a node-major tile + squared norms ----
            if (l == 0) {          // xyz coordinates, (x, y, z, 0) per node (sg_net.py:81)
                for (int n = tid; n < R; n += kThreads) {
                    const float x = sIn[n], y = sIn[N + n], z = sIn[2 * N + n];
                    *reinterpret_cast<float4*>(sX + n * XS) = make_float4(x, y, z, 0.0f);
                    sXX[n] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
                }
            } else {
                if (l == 3) {      // semantic channels: node-major [n][12] from input rows 3..14 (sg_net.py:82,94)
                    for (int e = tid; e < R * kLabels; e += kThreads) {
                        const int n = e % R, c = e / R;
                        sX[n * XS + c] = sIn[(3 + c) * N + n];
                    }
                    __syncthreads();
                }
                sq_norms(sX, sXX, D.cin4, R, tid);
            }
            __syncthreads();
            // ---- dynamic graph: k nearest neighbours in this layer's feature space (dgcnn.py:14-20) ----
            knn_gram_dispatch<NPL>(sX, sXX, sP, D.cin4, R, warp, lane);
            __syncthreads();
            knn_select<NPL>(sP, sIdx, sCnt, sXX, R, N, k, KS, warp, lane);
            if (tk) { __syncthreads(); trace_knn_rows<NPL>(tk + l * N * k, sP, sXX, N, R, k, tid); }
            __syncthreads();                                   // distance tile is dead: sY may take the GEMM output
            float* tr = tl ? tl + l * N * 64 : nullptr;
            if (l == 0) {
                xyz_layer1(sIn, sIdx, sCnt, KS, W.s1, sX, tr, N, R, warp, lane);
            } else {
                mbar_wait(barW, phW); phW ^= 1;
                if (D.cout == 64) node_gemm<4, 0>(sX, sW, sY, YS, nullptr, D.cin4, R, warp, lane);
                else              node_gemm<2, 0>(sX, sW, sY, YS, nullptr, D.cin4, R, warp, lane);
                __syncthreads();
                if (tid == 0) { mbar_expect_tx(barW, D.next_bytes); bulk_g2s(sW, D.next_w, D.next_bytes, barW); }
                if (D.cout == 64) gather_max_bn<64>(sY, sIdx, sCnt, KS, D.ab, sX, XS, tr, R, warp, lane);
                else              gather_max_bn<32>(sY, sIdx, sCnt, KS, D.ab, (l == 2) ? sCat : sCat + 32, XS, tr, R, warp, lane);
            }
            __syncthreads();
            trace_replicate_rows(tr, N, R, tid);
        }

        // ================= conv_end (sg_net.py:104-109): cat(xyz3, sem3) [R,64] -> [R,32] =================
        mbar_wait(barW, phW); phW ^= 1;
        float* sE = sX;   // node embeddings, stride XS (first 32 columns)
        node_gemm_nt<4, 1, 1>(sCat, sW, sE, XS, W.ab_end, 16, R, warp, lane);
        __syncthreads();
        // every trailing pad is a copy of row R-1
        for (int e = tid; e < (N - R) * kF3; e += kThreads) sE[(R + (e >> 5)) * XS + (e & 31)] = sE[(R - 1) * XS + (e & 31)];
        __syncthreads();
        if (A.emb) {
            float* eo = A.emb + static_cast<size_t>(g) * N * kF3;
            for (int e = tid; e < N * kF3; e += kThreads) eo[e] = sE[(e >> 5) * XS + (e & 31)];
        }

        // ================= attention pooling over all N nodes (layers_batch.py:28-39) =================
        // ctx[b] = tanh(mean_n sum_a E[n][a] Watt[a][b]): lane = b, warp strides over nodes
        {
            float colsum = 0.0f;
#pragma unroll 1
            for (int n = warp; n < N; n += kWarps) {
                float t = 0.0f;
#pragma unroll 4
                for (int a4 = 0; a4 < kF3 / 4; ++a4) {
                    const float4 e = *reinterpret_cast<const float4*>(sE + n * XS + 4 * a4);
                    t = fmaf(e.x, __ldg(W.att_w + (4 * a4 + 0) * kF3 + lane), t);
                    t = fmaf(e.y, __ldg(W.att_w + (4 * a4 + 1) * kF3 + lane), t);
                    t = fmaf(e.z, __ldg(W.att_w + (4 * a4 + 2) * kF3 + lane), t);
                    t = fmaf(e.w, __ldg(W.att_w + (4 * a4 + 3) * kF3 + lane), t);
                }
                colsum = __fadd_rn(colsum, t);
            }
            sRed[warp * 32 + lane] = colsum;
        }
        __syncthreads();
        float* sCtx = sRed + kWarps * 32;        // [32]
        float* sPool = sRed + kWarps * 32 + 32;  // [32]
        if (tid < kF3) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s = __fadd_rn(s, sRed[w * 32 + tid]);
            sCtx[tid] = tanhf(s / static_cast<float>(N));
        }
        __syncthreads();
        // att[n] = sigmoid(E[n] . ctx)
        float* sAtt = sXX;
        for (int n = tid; n < N; n += kThreads) {
            float s = 0.0f;
#pragma unroll 8
            for (int b = 0; b < kF3; ++b) s = fmaf(sE[n * XS + b], sCtx[b], s);
            const float a = sigmoidf_acc(s);
            sAtt[n] = a;
            float* ao = A.pairs ? ((g & 1) ? A.att1 : A.att0) : A.att0;
            if (ao) ao[static_cast<size_t>(A.pairs ? (g >> 1) : g) * N + n] = a;
        }
        __syncthreads();
        // pooled[a] = sum_n E[n][a] att[n]
        if (tid < kF3) {
            float s = 0.0f;
            for (int n = 0; n < N; ++n) s = fmaf(sE[n * XS + tid], sAtt[n], s);
            sPool[tid] = s;
            A.pooled[static_cast<size_t>(g) * kF3 + tid] = s;
        }

        // ================= pair head, run by whichever CTA of the pair finishes last =================
        if (A.pairs) {
            __threadfence();
            __syncthreads();
            if (tid == 0) sFlag = atomicAdd(A.counters + (g >> 1), 1);
            __syncthreads();
            if (sFlag == 1) {
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
                pair_head_cta(e1, e2, W, H, sY, A.score + (g >> 1), tid);
            }
        }
        __syncthreads();
    }
}

}  // namespace sgpr
