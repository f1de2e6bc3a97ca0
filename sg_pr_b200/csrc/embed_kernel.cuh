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
// Arithmetic form (exact refactor of the reference's 1x1 conv over the materialised [2C, N, k] edge tensor):
//     W [x_j - x_i ; x_i] = (A_j - A_i) + B_i,   A = Wa x (per node),  B = Wb x (per node)
//     max_j LReLU(BN(.))  = LReLU(alpha * ((max_j A_j - A_i) + B_i) + beta)   with alpha >= 0 (sign folded into W)
// so each layer is: k-NN (Gram tile in registers -> bitonic threshold select) -> per-node GEMM [N x C]x[C x 2C']
// -> gather-max over the k neighbour rows of A.  Layer 1 of the xyz branch keeps the reference's direct form
// W_a (x_j - x_i) because metre-scale coordinates would lose ~5 bits to cancellation in A_j - A_i.
//
// The five GEMM-form layers run through ONE copy of the phase code (a runtime layer loop): the kernel is
// issue/latency-bound, so instruction-cache footprint and instruction count matter more than anything else.
// Dot products use the packed fp32 FMA of sm_100 (fma.rn.f32x2, SASS FFMA2): the two halves of a pair carry the
// even-channel and odd-channel partial sums, added once at the end.
#pragma once
#include "common.cuh"

namespace sgpr {

struct EmbedArgs {
    const float* g0;        // graphs of side 0 (or all graphs when !pairs)   [*, 15, N]
    const float* g1;        // graphs of side 1 (pairs mode)
    int G;                  // number of graphs to embed (2*B in pairs mode: g = 2*b + side)
    int N, k, KS;           // KS = k rounded up to 4 (neighbour-list row stride in bytes)
    int pairs;
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
    int w, in, x, y, cat, xx, red, bar, idx, total;   // byte offsets
};

__host__ __device__ inline SmemLayout make_layout(int nmax, int ks) {
    SmemLayout L;
    int o = 0;
    L.w = o;   o += 64 * 128 * 4;              // largest packed layer matrix (64 in x 128 out)
    L.in = o;  o += ((kInCh * nmax * 4 + 15) / 16) * 16;
    L.x = o;   o += nmax * XS * 4;
    L.y = o;   o += nmax * YS * 4;
    L.cat = o; o += nmax * XS * 4;
    L.xx = o;  o += nmax * 4;
    L.red = o; o += (kWarps * 32 + 64) * 4;
    L.bar = o; o += 16;
    L.idx = o; o += ((nmax * ks + 15) / 16) * 16;
    L.total = o;
    return L;
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ------------------------------------------------------------------------------------------------------------
// Bitonic sort (ascending) of 32*NPL floats held NPL per lane, element e = q*32 + lane.
// "Flip" formulation: every merge starts with a mirror exchange (partner e ^ (size-1)) and continues with
// half-cleaners (partner e ^ d); every exchange is ascending, so the keep-min predicate is one lane bit.
// ------------------------------------------------------------------------------------------------------------
template <int NPL>
__device__ __forceinline__ void bitonic_sort_asc(float (&v)[NPL], int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * NPL; size <<= 1) {
        // ---- mirror step ----
        if (size <= 32) {
            const bool keep_min = (lane & (size >> 1)) == 0;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const float p = __shfl_xor_sync(0xffffffffu, v[q], size - 1);
                v[q] = keep_min ? fminf(v[q], p) : fmaxf(v[q], p);
            }
        } else {
            const int mq = (size >> 5) - 1;          // register mirror mask
            const int hb = size >> 6;                // q bit that decides lower/upper half of the block
            float p[NPL];
#pragma unroll
            for (int q = 0; q < NPL; ++q) p[q] = __shfl_xor_sync(0xffffffffu, v[q ^ mq], 31);
#pragma unroll
            for (int q = 0; q < NPL; ++q) v[q] = ((q & hb) == 0) ? fminf(v[q], p[q]) : fmaxf(v[q], p[q]);
        }
        // ---- half-cleaners ----
#pragma unroll
        for (int d = size >> 2; d >= 1; d >>= 1) {
            if (d >= 32) {
                const int dq = d >> 5;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    if ((q & dq) == 0) {
                        const float lo = fminf(v[q], v[q | dq]);
                        const float hi = fmaxf(v[q], v[q | dq]);
                        v[q] = lo;
                        v[q | dq] = hi;
                    }
                }
            } else {
                const bool keep_min = (lane & d) == 0;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const float p = __shfl_xor_sync(0xffffffffu, v[q], d);
                    v[q] = keep_min ? fminf(v[q], p) : fmaxf(v[q], p);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k-NN of every node over the node-major tile sX[n][4*c4n] (dgcnn.py:14-20).
//   pd[i][j] = (2*dot(x_i,x_j) - xx_j) - xx_i     == -xx - inner - xx^T with inner = -2*dot, same rounding order
//   select the k largest per row; ties at the k-th value go to the lowest indices.  Only the SET matters
//   downstream (max over neighbours), so the list is written in ascending index order.
// A warp owns RG rows at a time: Gram tile in registers (RG x NPL per lane), sort a copy, threshold, compact.
// ------------------------------------------------------------------------------------------------------------
template <int NPL>
__device__ __forceinline__ void knn_phase(const float* __restrict__ sX, const float* __restrict__ sXX,
                                          uint8_t* __restrict__ sIdx, int c4n, int N, int k, int KS, int warp,
                                          int lane) {
    constexpr int NMAX = 32 * NPL;
    constexpr int RG = (NPL >= 4) ? 2 : 4;
    const int ngroups = (N + RG - 1) / RG;
    const uint32_t lt = (1u << lane) - 1u;

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
            float4 a[RG], b[NPL];
#pragma unroll
            for (int r = 0; r < RG; ++r) a[r] = *reinterpret_cast<const float4*>(pa + r * XS + 4 * c);
#pragma unroll
            for (int q = 0; q < NPL; ++q) b[q] = *reinterpret_cast<const float4*>(pb + 32 * q * XS + 4 * c);
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    acc[r][q] = ffma2(make_float2(a[r].x, a[r].y), make_float2(b[q].x, b[q].y), acc[r][q]);
                    acc[r][q] = ffma2(make_float2(a[r].z, a[r].w), make_float2(b[q].z, b[q].w), acc[r][q]);
                }
        }

        float xxj[NPL];
#pragma unroll
        for (int q = 0; q < NPL; ++q) xxj[q] = sXX[lane + 32 * q];

#pragma unroll
        for (int r = 0; r < RG; ++r) {
            const int i = i0 + r;
            if (i >= N) break;   // warp-uniform
            const float xxi = sXX[i];
            float o[NPL], v[NPL];
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const float dot = __fadd_rn(acc[r][q].x, acc[r][q].y);
                const float t = __fsub_rn(__fmul_rn(2.0f, dot), xxj[q]);
                const float pd = __fsub_rn(t, xxi);
                o[q] = (lane + 32 * q < N) ? pd : -INFINITY;
                v[q] = o[q];
            }
            bitonic_sort_asc<NPL>(v, lane);
            // k-th largest value = sorted position NMAX - k
            const int P = NMAX - k;
            float thr = 0.0f;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const float cand = __shfl_sync(0xffffffffu, v[q], P & 31);
                if (q == (P >> 5)) thr = cand;
            }
            uint32_t gt[NPL], eq[NPL];
            int ngt = 0;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                gt[q] = __ballot_sync(0xffffffffu, o[q] > thr);
                eq[q] = __ballot_sync(0xffffffffu, o[q] == thr);
                ngt += __popc(gt[q]);
            }
            const int need = k - ngt;        // how many of the tied-at-threshold elements to take (>= 1)
            int eq_before = 0, base = 0;
            uint8_t* row = sIdx + i * KS;
            int first = 0;
            bool have_first = false;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const bool mine_eq = ((eq[q] >> lane) & 1u) && (eq_before + __popc(eq[q] & lt) < need);
                const uint32_t sel = gt[q] | __ballot_sync(0xffffffffu, mine_eq);
                if ((sel >> lane) & 1u) row[base + __popc(sel & lt)] = static_cast<uint8_t>(lane + 32 * q);
                if (!have_first && sel) { first = (__ffs(sel) - 1) + 32 * q; have_first = true; }
                eq_before += __popc(eq[q]);
                base += __popc(sel);
            }
            // pad the list to a multiple of 4 with a repeat of its first entry (harmless under max)
            if (lane < KS - k) row[k + lane] = static_cast<uint8_t>(first);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Per-node GEMM: out[n][co] = sum_ci X[n][ci] * W[ci][co], co in [0, 32*CPL).
// A warp owns NT nodes x all outputs; a lane owns CPL consecutive outputs.  W is packed in channel PAIRS
// (pack.hpp: pack_pairs): row p holds, for every output, (W[2p][co], W[2p+1][co]) so that one FFMA2 advances the
// even- and odd-channel partial sums of an output.  EPI: 0 = store raw, 1 = BN(alpha,beta)+LeakyReLU (conv_end).
// ------------------------------------------------------------------------------------------------------------
template <int CPL, int EPI>
__device__ __forceinline__ void node_gemm(const float* __restrict__ sXin, const float* __restrict__ sW,
                                          float* __restrict__ sOut, int outStride, const float* __restrict__ ab,
                                          int cin4, int N, int warp, int lane) {
    constexpr int NT = 8;
    constexpr int CO = 32 * CPL;
    constexpr int ROW = 2 * CO;                  // floats per channel-pair row
    const int nchunks = (N + NT - 1) / NT;
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

// ------------------------------------------------------------------------------------------------------------
// Gather-max + BN + LeakyReLU (sg_net.py:85-86 etc.): for node i and channel c
//     out = LReLU(alpha_c * ((max_{j in knn(i)} A[j][c] - A[i][c]) + B[i][c]) + beta_c)
// sY row = [A(0..COUT) | B(COUT..2COUT)].  A warp owns a node, a lane owns COUT/32 channels.
// ------------------------------------------------------------------------------------------------------------
template <int COUT>
__device__ __forceinline__ void gather_max_bn(const float* __restrict__ sY, const uint8_t* __restrict__ sIdx, int KS,
                                              const float* __restrict__ ab, float* __restrict__ sDst, int dstStride,
                                              float* __restrict__ trace, int N, int warp, int lane) {
    constexpr int CPL = COUT / 32;
    float al[CPL], be[CPL];
#pragma unroll
    for (int p = 0; p < CPL; ++p) { al[p] = __ldg(ab + lane * CPL + p); be[p] = __ldg(ab + COUT + lane * CPL + p); }

#pragma unroll 1
    for (int i = warp; i < N; i += kWarps) {
        float m[CPL];
#pragma unroll
        for (int p = 0; p < CPL; ++p) m[p] = -INFINITY;
        const uint8_t* row = sIdx + i * KS;
        const float* base = sY + lane * CPL;
#pragma unroll 2
        for (int t = 0; t < KS; t += 4) {
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
//     e = wa0*d0 + wa1*d1 + wa2*d2  (d = x_j - x_i, sequential FMA), max over the k edges, then the centre
//     terms wb.x_i appended in the same sequential order (monotone in e, so they commute with the max).
// sIn is the channel-major input block [15][N]; a lane owns output channels 2*lane, 2*lane+1.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void xyz_layer1(const float* __restrict__ sIn, const uint8_t* __restrict__ sIdx, int KS,
                                           const float* __restrict__ s1, float* __restrict__ sDst,
                                           float* __restrict__ trace, int N, int warp, int lane) {
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane) * 2);
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane) * 2 + 1);
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane + 1) * 2);
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(s1) + (2 * lane + 1) * 2 + 1);
    // p0 = {wa0,wa1,wa2,wb0}, p1 = {wb1,wb2,alpha,beta} of channel 2*lane; r0/r1 the same for 2*lane+1
    const float2 wa0 = make_float2(p0.x, r0.x), wa1 = make_float2(p0.y, r0.y), wa2 = make_float2(p0.z, r0.z);
#pragma unroll 1
    for (int i = warp; i < N; i += kWarps) {
        const float xi0 = sIn[i], xi1 = sIn[N + i], xi2 = sIn[2 * N + i];
        float m0 = -INFINITY, m1 = -INFINITY;
        const uint8_t* row = sIdx + i * KS;
#pragma unroll 1
        for (int t = 0; t < KS; t += 4) {
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
__device__ __forceinline__ void sq_norms(const float* __restrict__ sX, float* __restrict__ sXX, int c4n, int N, int tid) {
    const int part = tid & 3;
    const int per = (c4n + 3) >> 2;                     // float4 groups per quarter
    for (int n0 = 0; n0 < N; n0 += kThreads / 4) {      // warp-uniform trip count
        const int n = n0 + (tid >> 2);
        float s = 0.0f;
        if (n < N) {
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
        if (n < N && part == 0) sXX[n] = s;
    }
}

__device__ __forceinline__ void trace_knn_rows(uint8_t* __restrict__ dst, const uint8_t* __restrict__ sIdx, int N,
                                               int k, int KS, int tid) {
    if (!dst) return;
    for (int e = tid; e < N * k; e += kThreads) dst[e] = sIdx[(e / k) * KS + (e % k)];
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
    const float* w;        // this layer's packed matrix (already requested into sW by the previous step)
    const float* ab;       // alpha | beta
    const float* next_w;   // matrix to prefetch into sW once this layer's GEMM has consumed sW
    int next_bytes;
    int cin4;              // input channels / 4
    int cout;              // 64 or 32
};

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
    float* sCat = reinterpret_cast<float*>(smem + L.cat);
    float* sXX = reinterpret_cast<float*>(smem + L.xx);
    float* sRed = reinterpret_cast<float*>(smem + L.red);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar);
    uint8_t* sIdx = smem + L.idx;
    __shared__ int sFlag;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = A.N, k = A.k, KS = A.KS;
    uint64_t* barIn = bars;
    uint64_t* barW = bars + 1;
    uint32_t phIn = 0, phW = 0;

    if (tid == 0) { mbar_init(barIn, 1); mbar_init(barW, 1); fence_mbar_init(); }
    // zero the feature tiles once so rows >= N never hold junk
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

        // ================= xyz layer 1 (direct form) =================
        for (int n = tid; n < N; n += kThreads) {
            const float x = sIn[n], y = sIn[N + n], z = sIn[2 * N + n];
            *reinterpret_cast<float4*>(sX + n * XS) = make_float4(x, y, z, 0.0f);
            sXX[n] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        }
        __syncthreads();
        knn_phase<NPL>(sX, sXX, sIdx, 1, N, k, KS, warp, lane);
        __syncthreads();
        trace_knn_rows(tk, sIdx, N, k, KS, tid);
        xyz_layer1(sIn, sIdx, KS, W.s1, sX, tl, N, warp, lane);
        __syncthreads();

        // ================= the five GEMM-form EdgeConv layers: xyz 2,3 then sem 1,2,3 =================
#pragma unroll 1
        for (int l = 1; l < 6; ++l) {
            LayerDesc D;
            switch (l) {
                case 1:  D = LayerDesc{W.w_s2, W.ab_s2, W.w_s3, 64 * 64 * 4, 16, 64}; break;
                case 2:  D = LayerDesc{W.w_s3, W.ab_s3, W.w_f1, 12 * 128 * 4, 16, 32}; break;
                case 3:  D = LayerDesc{W.w_f1, W.ab_f1, W.w_f2, 64 * 128 * 4, 3, 64}; break;
                case 4:  D = LayerDesc{W.w_f2, W.ab_f2, W.w_f3, 64 * 64 * 4, 16, 64}; break;
                default: D = LayerDesc{W.w_f3, W.ab_f3, W.w_end, 64 * 32 * 4, 16, 32}; break;
            }
            if (l == 3) {   // semantic branch input: node-major [n][12] from input rows 3..14 (sg_net.py:82,94)
                for (int e = tid; e < N * kLabels; e += kThreads) {
                    const int n = e % N, c = e / N;
                    sX[n * XS + c] = sIn[(3 + c) * N + n];
                }
                __syncthreads();
            }
            sq_norms(sX, sXX, D.cin4, N, tid);
            __syncthreads();
            knn_phase<NPL>(sX, sXX, sIdx, D.cin4, N, k, KS, warp, lane);
            mbar_wait(barW, phW); phW ^= 1;
            if (D.cout == 64) node_gemm<4, 0>(sX, sW, sY, YS, nullptr, D.cin4, N, warp, lane);
            else              node_gemm<2, 0>(sX, sW, sY, YS, nullptr, D.cin4, N, warp, lane);
            __syncthreads();
            if (tid == 0) { mbar_expect_tx(barW, D.next_bytes); bulk_g2s(sW, D.next_w, D.next_bytes, barW); }
            trace_knn_rows(tk ? tk + l * N * k : nullptr, sIdx, N, k, KS, tid);
            float* tr = tl ? tl + l * N * 64 : nullptr;
            if (D.cout == 64) gather_max_bn<64>(sY, sIdx, KS, D.ab, sX, XS, tr, N, warp, lane);
            else              gather_max_bn<32>(sY, sIdx, KS, D.ab, (l == 2) ? sCat : sCat + 32, XS, tr, N, warp, lane);
            __syncthreads();
        }

        // ================= conv_end (sg_net.py:104-109): cat(xyz3, sem3) [N,64] -> [N,32] =================
        mbar_wait(barW, phW); phW ^= 1;
        float* sE = sX;   // node embeddings, stride XS (first 32 columns)
        node_gemm<1, 1>(sCat, sW, sE, XS, W.ab_end, 16, N, warp, lane);
        __syncthreads();
        if (A.emb) {
            float* eo = A.emb + static_cast<size_t>(g) * N * kF3;
            for (int e = tid; e < N * kF3; e += kThreads) eo[e] = sE[(e >> 5) * XS + (e & 31)];
        }

        // ================= attention pooling (layers_batch.py:28-39) =================
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
