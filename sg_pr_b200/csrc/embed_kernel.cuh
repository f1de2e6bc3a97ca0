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
    L.w = o;   o += 64 * 128 * 4;              // largest packed layer matrix [64][128]
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

// ------------------------------------------------------------------------------------------------------------
// Bitonic sort (ascending) of 32*NPL floats held NPL per lane, element e = q*32 + lane.
// ------------------------------------------------------------------------------------------------------------
template <int NPL>
__device__ __forceinline__ void bitonic_sort_asc(float (&v)[NPL], int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * NPL; size <<= 1) {
#pragma unroll
        for (int d = size >> 1; d >= 1; d >>= 1) {
            if (d >= 32) {
                const int dq = d >> 5;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    if ((q & dq) == 0) {
                        const bool asc = ((q * 32) & size) == 0;
                        const float lo = fminf(v[q], v[q | dq]);
                        const float hi = fmaxf(v[q], v[q | dq]);
                        v[q] = asc ? lo : hi;
                        v[q | dq] = asc ? hi : lo;
                    }
                }
            } else {
                const bool lower = (lane & d) == 0;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const float p = __shfl_xor_sync(0xffffffffu, v[q], d);
                    const bool asc = (((q * 32) | lane) & size) == 0;
                    v[q] = (lower == asc) ? fminf(v[q], p) : fmaxf(v[q], p);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k-NN of every node over the node-major tile sX[n][4*C4] (dgcnn.py:14-20).
//   pd[i][j] = (2*dot(x_i,x_j) - xx_j) - xx_i     == -xx - inner - xx^T with inner = -2*dot, same rounding order
//   select the k largest per row; ties at the k-th value go to the lowest indices.  Only the SET matters
//   downstream (max over neighbours), so the list is written in ascending index order.
// A warp owns RG rows at a time: Gram tile in registers (RG x NPL per lane), sort a copy, threshold, compact.
// ------------------------------------------------------------------------------------------------------------
template <int NPL, int C4>
__device__ __forceinline__ void knn_phase(const float* __restrict__ sX, const float* __restrict__ sXX,
                                          uint8_t* __restrict__ sIdx, int N, int k, int KS, int warp, int lane) {
    constexpr int NMAX = 32 * NPL;
    constexpr int RG = (NPL >= 4) ? 2 : 4;
    const int ngroups = (N + RG - 1) / RG;
    const uint32_t lt = (1u << lane) - 1u;

    for (int g = warp; g < ngroups; g += kWarps) {
        const int i0 = g * RG;
        float acc[RG][NPL];
#pragma unroll
        for (int r = 0; r < RG; ++r)
#pragma unroll
            for (int q = 0; q < NPL; ++q) acc[r][q] = 0.0f;

#pragma unroll
        for (int c = 0; c < C4; ++c) {
            float4 a[RG], b[NPL];
#pragma unroll
            for (int r = 0; r < RG; ++r) a[r] = *reinterpret_cast<const float4*>(sX + (i0 + r) * XS + 4 * c);
#pragma unroll
            for (int q = 0; q < NPL; ++q) b[q] = *reinterpret_cast<const float4*>(sX + (lane + 32 * q) * XS + 4 * c);
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    float s = acc[r][q];
                    s = fmaf(a[r].x, b[q].x, s);
                    s = fmaf(a[r].y, b[q].y, s);
                    s = fmaf(a[r].z, b[q].z, s);
                    s = fmaf(a[r].w, b[q].w, s);
                    acc[r][q] = s;
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
                const float t = __fsub_rn(__fmul_rn(2.0f, acc[r][q]), xxj[q]);
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
// Per-node GEMM: out[n][co] = sum_ci X[n][ci] * W[ci][co], co in [0, 32*CPL), sequential FMA over ci ascending.
// A warp owns NT nodes x all outputs; a lane owns CPL consecutive outputs.  EPI: 0 = store raw,
// 1 = BN(alpha,beta)+LeakyReLU (conv_end, sg_net.py:74-76,105).
// ------------------------------------------------------------------------------------------------------------
template <int CIN4, int CPL, int EPI>
__device__ __forceinline__ void node_gemm(const float* __restrict__ sXin, const float* __restrict__ sW,
                                          float* __restrict__ sOut, int outStride, const float* __restrict__ ab,
                                          int N, int warp, int lane) {
    constexpr int NT = 8;
    constexpr int CO = 32 * CPL;
    const int nchunks = (N + NT - 1) / NT;
    for (int ch = warp; ch < nchunks; ch += kWarps) {
        const int n0 = ch * NT;
        float acc[NT][CPL];
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int p = 0; p < CPL; ++p) acc[n][p] = 0.0f;

#pragma unroll 4
        for (int c4 = 0; c4 < CIN4; ++c4) {
            float4 x[NT];
#pragma unroll
            for (int n = 0; n < NT; ++n) x[n] = *reinterpret_cast<const float4*>(sXin + (n0 + n) * XS + 4 * c4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float w[CPL];
                const float* wp = sW + (4 * c4 + q) * CO + lane * CPL;
                if constexpr (CPL == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(wp);
                    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
                } else if constexpr (CPL == 2) {
                    const float2 t = *reinterpret_cast<const float2*>(wp);
                    w[0] = t.x; w[1] = t.y;
                } else {
                    w[0] = *wp;
                }
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const float xv = (q == 0) ? x[n].x : (q == 1) ? x[n].y : (q == 2) ? x[n].z : x[n].w;
#pragma unroll
                    for (int p = 0; p < CPL; ++p) acc[n][p] = fmaf(xv, w[p], acc[n][p]);
                }
            }
        }

        float al[CPL], be[CPL];
        if constexpr (EPI == 1) {
#pragma unroll
            for (int p = 0; p < CPL; ++p) { al[p] = __ldg(ab + lane * CPL + p); be[p] = __ldg(ab + CO + lane * CPL + p); }
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            float* op = sOut + (n0 + n) * outStride + lane * CPL;
#pragma unroll
            for (int p = 0; p < CPL; ++p) {
                float y = acc[n][p];
                if constexpr (EPI == 1) y = lrelu(fmaf(y, al[p], be[p]));
                acc[n][p] = y;
            }
            if constexpr (CPL == 4) *reinterpret_cast<float4*>(op) = make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]);
            else if constexpr (CPL == 2) *reinterpret_cast<float2*>(op) = make_float2(acc[n][0], acc[n][1]);
            else *op = acc[n][0];
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

    for (int i = warp; i < N; i += kWarps) {
        float m[CPL];
#pragma unroll
        for (int p = 0; p < CPL; ++p) m[p] = -INFINITY;
        const uint8_t* row = sIdx + i * KS;
        for (int t = 0; t < KS; t += 4) {
            const uchar4 jj = *reinterpret_cast<const uchar4*>(row + t);
            if constexpr (CPL == 2) {
                const float2 a0 = *reinterpret_cast<const float2*>(sY + jj.x * YS + 2 * lane);
                const float2 a1 = *reinterpret_cast<const float2*>(sY + jj.y * YS + 2 * lane);
                const float2 a2 = *reinterpret_cast<const float2*>(sY + jj.z * YS + 2 * lane);
                const float2 a3 = *reinterpret_cast<const float2*>(sY + jj.w * YS + 2 * lane);
                m[0] = fmaxf(fmaxf(m[0], fmaxf(a0.x, a1.x)), fmaxf(a2.x, a3.x));
                m[1] = fmaxf(fmaxf(m[1], fmaxf(a0.y, a1.y)), fmaxf(a2.y, a3.y));
            } else {
                const float a0 = sY[jj.x * YS + lane], a1 = sY[jj.y * YS + lane];
                const float a2 = sY[jj.z * YS + lane], a3 = sY[jj.w * YS + lane];
                m[0] = fmaxf(fmaxf(m[0], fmaxf(a0, a1)), fmaxf(a2, a3));
            }
        }
#pragma unroll
        for (int p = 0; p < CPL; ++p) {
            const float ai = sY[i * YS + lane * CPL + p];
            const float bi = sY[i * YS + COUT + lane * CPL + p];
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
    // p0 = {wa0,wa1,wa2,wb0}, p1 = {wb1,wb2,alpha,beta}
    for (int i = warp; i < N; i += kWarps) {
        const float xi0 = sIn[i], xi1 = sIn[N + i], xi2 = sIn[2 * N + i];
        float m0 = -INFINITY, m1 = -INFINITY;
        const uint8_t* row = sIdx + i * KS;
        for (int t = 0; t < KS; t += 4) {
            const uchar4 jj = *reinterpret_cast<const uchar4*>(row + t);
            const int js[4] = {jj.x, jj.y, jj.z, jj.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = js[u];
                const float d0 = __fsub_rn(sIn[j], xi0);
                const float d1 = __fsub_rn(sIn[N + j], xi1);
                const float d2 = __fsub_rn(sIn[2 * N + j], xi2);
                float e0 = __fmul_rn(p0.x, d0); e0 = fmaf(p0.y, d1, e0); e0 = fmaf(p0.z, d2, e0);
                float e1 = __fmul_rn(r0.x, d0); e1 = fmaf(r0.y, d1, e1); e1 = fmaf(r0.z, d2, e1);
                m0 = fmaxf(m0, e0);
                m1 = fmaxf(m1, e1);
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

// squared norms per node, sequential over channels: xx = sum_c x_c^2   (dgcnn.py:16, products rounded, then added)
template <int C>
__device__ __forceinline__ void sq_norms(const float* __restrict__ sX, float* __restrict__ sXX, int N, int tid) {
    for (int n = tid; n < N; n += kThreads) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) { const float x = sX[n * XS + c]; s = __fadd_rn(s, __fmul_rn(x, x)); }
        sXX[n] = s;
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

// ------------------------------------------------------------------------------------------------------------
// The fused kernel.
// ------------------------------------------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(kThreads, (NPL <= 2) ? 2 : 1)
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
    __syncthreads();

    const uint32_t inBytes = static_cast<uint32_t>(kInCh * N * 4);

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

        // ================= xyz branch =================
        // layer 1 operand: node-major [n][4] = (x, y, z, 0)
        for (int n = tid; n < N; n += kThreads) {
            const float x = sIn[n], y = sIn[N + n], z = sIn[2 * N + n];
            *reinterpret_cast<float4*>(sX + n * XS) = make_float4(x, y, z, 0.0f);
            sXX[n] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        }
        __syncthreads();
        knn_phase<NPL, 1>(sX, sXX, sIdx, N, k, KS, warp, lane);
        __syncthreads();
        trace_knn_rows(tk, sIdx, N, k, KS, tid);
        xyz_layer1(sIn, sIdx, KS, W.s1, sX, tl, N, warp, lane);
        __syncthreads();

        // layer 2: 64 -> 64
        sq_norms<64>(sX, sXX, N, tid);
        __syncthreads();
        knn_phase<NPL, 16>(sX, sXX, sIdx, N, k, KS, warp, lane);
        mbar_wait(barW, phW); phW ^= 1;
        node_gemm<16, 4, 0>(sX, sW, sY, YS, nullptr, N, warp, lane);
        __syncthreads();
        if (tid == 0) { mbar_expect_tx(barW, 64 * 64 * 4); bulk_g2s(sW, W.w_s3, 64 * 64 * 4, barW); }
        trace_knn_rows(tk ? tk + 1 * N * k : nullptr, sIdx, N, k, KS, tid);
        gather_max_bn<64>(sY, sIdx, KS, W.ab_s2, sX, XS, tl ? tl + 1 * N * 64 : nullptr, N, warp, lane);
        __syncthreads();

        // layer 3: 64 -> 32, result into sCat[:, 0:32]
        sq_norms<64>(sX, sXX, N, tid);
        __syncthreads();
        knn_phase<NPL, 16>(sX, sXX, sIdx, N, k, KS, warp, lane);
        mbar_wait(barW, phW); phW ^= 1;
        node_gemm<16, 2, 0>(sX, sW, sY, YS, nullptr, N, warp, lane);
        __syncthreads();
        if (tid == 0) { mbar_expect_tx(barW, 12 * 128 * 4); bulk_g2s(sW, W.w_f1, 12 * 128 * 4, barW); }
        trace_knn_rows(tk ? tk + 2 * N * k : nullptr, sIdx, N, k, KS, tid);
        gather_max_bn<32>(sY, sIdx, KS, W.ab_s3, sCat, XS, tl ? tl + 2 * N * 64 : nullptr, N, warp, lane);
        __syncthreads();

        // ================= semantic branch =================
        // layer 1 operand: node-major [n][12] from input rows 3..14
        for (int e = tid; e < N * kLabels; e += kThreads) {
            const int n = e % N, c = e / N;
            sX[n * XS + c] = sIn[(3 + c) * N + n];
        }
        __syncthreads();
        sq_norms<kLabels>(sX, sXX, N, tid);
        __syncthreads();
        knn_phase<NPL, 3>(sX, sXX, sIdx, N, k, KS, warp, lane);
        mbar_wait(barW, phW); phW ^= 1;
        node_gemm<3, 4, 0>(sX, sW, sY, YS, nullptr, N, warp, lane);
        __syncthreads();
        if (tid == 0) { mbar_expect_tx(barW, 64 * 128 * 4); bulk_g2s(sW, W.w_f2, 64 * 128 * 4, barW); }
        trace_knn_rows(tk ? tk + 3 * N * k : nullptr, sIdx, N, k, KS, tid);
        gather_max_bn<64>(sY, sIdx, KS, W.ab_f1, sX, XS, tl ? tl + 3 * N * 64 : nullptr, N, warp, lane);
        __syncthreads();

        // layer 2
        sq_norms<64>(sX, sXX, N, tid);
        __syncthreads();
        knn_phase<NPL, 16>(sX, sXX, sIdx, N, k, KS, warp, lane);
        mbar_wait(barW, phW); phW ^= 1;
        node_gemm<16, 4, 0>(sX, sW, sY, YS, nullptr, N, warp, lane);
        __syncthreads();
        if (tid == 0) { mbar_expect_tx(barW, 64 * 64 * 4); bulk_g2s(sW, W.w_f3, 64 * 64 * 4, barW); }
        trace_knn_rows(tk ? tk + 4 * N * k : nullptr, sIdx, N, k, KS, tid);
        gather_max_bn<64>(sY, sIdx, KS, W.ab_f2, sX, XS, tl ? tl + 4 * N * 64 : nullptr, N, warp, lane);
        __syncthreads();

        // layer 3, result into sCat[:, 32:64]
        sq_norms<64>(sX, sXX, N, tid);
        __syncthreads();
        knn_phase<NPL, 16>(sX, sXX, sIdx, N, k, KS, warp, lane);
        mbar_wait(barW, phW); phW ^= 1;
        node_gemm<16, 2, 0>(sX, sW, sY, YS, nullptr, N, warp, lane);
        __syncthreads();
        if (tid == 0) { mbar_expect_tx(barW, 64 * 32 * 4); bulk_g2s(sW, W.w_end, 64 * 32 * 4, barW); }
        trace_knn_rows(tk ? tk + 5 * N * k : nullptr, sIdx, N, k, KS, tid);
        gather_max_bn<32>(sY, sIdx, KS, W.ab_f3, sCat + 32, XS, tl ? tl + 5 * N * 64 : nullptr, N, warp, lane);
        __syncthreads();

        // ================= conv_end (sg_net.py:104-109): cat(xyz3, sem3) [N,64] -> [N,32] =================
        mbar_wait(barW, phW); phW ^= 1;
        float* sE = sX;   // node embeddings, stride XS (first 32 columns)
        node_gemm<16, 1, 1>(sCat, sW, sE, XS, W.ab_end, N, warp, lane);
        __syncthreads();
        if (A.emb) {
            float* eo = A.emb + static_cast<size_t>(g) * N * kF3;
            for (int e = tid; e < N * kF3; e += kThreads) eo[e] = sE[(e >> 5) * XS + (e & 31)];
        }

        // ================= attention pooling (layers_batch.py:28-39) =================
        // ctx[b] = tanh(mean_n sum_a E[n][a] Watt[a][b]): lane = b, warp strides over nodes
        {
            float wcol[kF3];
#pragma unroll
            for (int a = 0; a < kF3; ++a) wcol[a] = __ldg(W.att_w + a * kF3 + lane);
            float colsum = 0.0f;
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
#pragma unroll
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
