// Training step kernels: SGTrainer.process_batch(batch, training=True) from the point where the batch tensors exist
// (/root/reference/sg_net.py:332-338): SG.forward in TRAIN mode (sg_net.py:112-138 with the seven BatchNorm layers of
// sg_net.py:50-76 normalising by batch statistics, one BatchNorm batch per side because dgcnn_conv_pass is called once
// per side, sg_net.py:123-124), mean binary cross entropy (sg_net.py:335), backward, Adam with L2 weight decay
// (sg_net.py:351-352), running-statistics update (momentum 0.1, unbiased variance).
//
// Train-mode BatchNorm couples every graph of a side through the per-channel batch statistics, so — unlike the eval
// path's single fused kernel — a layer cannot start before the previous layer's statistics are complete.  The step is a
// short chain of launches with a grid-wide dependency (the kernel boundary) exactly where BatchNorm puts one:
//
//   pack | edge_fwd l=0,1,2 (xyz and sem branches of both sides in one grid) | end_fwd
//        | pool_head (conv_end BN, attention fwd, pair head fwd + loss + bwd, attention bwd)
//        | end_bwd | edge_bwd l=2,1,0 | adam (+ running statistics, loss)
//
// Arithmetic form (tests/train_model.py holds the same algebra in torch, checked against autograd on the CPU):
//   forward   y_ij = (A_j - A_i) + B_i,  A = Wa x, B = Wb x per node; BN+LeakyReLU is monotone per channel, so the max over
//             neighbours is taken on y (max if gamma >= 0 else min) and only [node][channel] tensors are ever stored;
//             sum_ij y and sum_ij y^2 accumulate in fp64 per (side, layer, channel).
//   backward  dy_ij = s*gz_i*[j = ext] - r - q*y_ij  (s = gamma*istd, q = s*dgamma/e*istd, r = s*dbeta/e - q*mu) folds the
//             BatchNorm backward into per-node terms: dB_i = T_i = s*gz_i - k*r - q*sum_j y_ij and
//             dA_n = -T_n + s*S1_n - deg_n*(r + q*A_n) - q*S2_n with S1_n = sum of gz_i over the nodes i whose extreme
//             neighbour is n, S2_n = sum of D_i = B_i - A_i over the nodes i that list n, deg_n = how many do.
//   k-NN indices carry no gradient (topk indices, dgcnn.py:19).
//
// Per-parameter gradients are accumulated in registers across the graphs a persistent CTA processes and written as
// per-CTA partials; the Adam kernel sums the partials in a fixed order (deterministic, no float atomics).
#pragma once
#include "../../include/sgpr_b200.h"
#include "common.cuh"
#include "embed_kernel.cuh"

namespace sgpr {
namespace train {

// ---- flat parameter vector (floats): conv weights in their state_dict shapes, BN gamma|beta per layer, head --------
constexpr int P_S1W = 0;                    // dgcnn_s_conv1.0.weight [64][6]
constexpr int P_S2W = P_S1W + 64 * 6;       // dgcnn_s_conv2.0.weight [64][128]
constexpr int P_S3W = P_S2W + 64 * 128;     // dgcnn_s_conv3.0.weight [32][128]
constexpr int P_F1W = P_S3W + 32 * 128;     // dgcnn_f_conv1.0.weight [64][24]
constexpr int P_F2W = P_F1W + 64 * 24;      // dgcnn_f_conv2.0.weight [64][128]
constexpr int P_F3W = P_F2W + 64 * 128;     // dgcnn_f_conv3.0.weight [32][128]
constexpr int P_ENDW = P_F3W + 32 * 128;    // dgcnn_conv_end.0.weight [32][64]
constexpr int P_BN = P_ENDW + 32 * 64;      // 7 x (gamma[C] | beta[C]), layers s1 s2 s3 f1 f2 f3 end
constexpr int kBnFloats = 2 * (64 + 64 + 32 + 64 + 64 + 32 + 32);   // 704
constexpr int P_ATT = P_BN + kBnFloats;     // attention.weight_matrix [32][32]
constexpr int P_NTNW = P_ATT + 32 * 32;     // tensor_network.weight_matrix [32][32][16]
constexpr int P_NTNV = P_NTNW + 32 * 32 * 16;   // tensor_network.weight_matrix_block [16][64]
constexpr int P_NTNB = P_NTNV + 16 * 64;    // tensor_network.bias [16]
constexpr int P_FC1W = P_NTNB + 16;         // fully_connected_first.weight [16][16]
constexpr int P_FC1B = P_FC1W + 256;
constexpr int P_FC2W = P_FC1B + 16;         // scoring_layer.weight [1][16]
constexpr int P_FC2B = P_FC2W + 16;
constexpr int P_TOTAL = P_FC2B + 1;         // 47,985 trainable floats
constexpr int R_OFF = P_TOTAL;              // 7 x (running_mean[C] | running_var[C])
constexpr int STATE_TOTAL = P_TOTAL + kBnFloats;
constexpr int kHeadFloats = P_TOTAL - P_NTNW;   // 17,713: the head's parameters are contiguous

// packed (channel-pair layout, pack.hpp::pair_index) copies of the GEMM matrices, rebuilt every step
constexpr int WPK_S2 = 0, WPK_S3 = WPK_S2 + 64 * 128, WPK_F1 = WPK_S3 + 64 * 64, WPK_F2 = WPK_F1 + 12 * 128,
              WPK_F3 = WPK_F2 + 64 * 128, WPK_END = WPK_F3 + 64 * 64, WPK_TOTAL = WPK_END + 64 * 32;

// backward copies for dX = dA Wa + dB Wb of the layers with 64 input channels: per layer two [cout][64] matrices
// (input = this layer's output channel, output = its input channel) in the same channel-pair layout
constexpr int WT_S2 = WPK_TOTAL, WT_S3 = WT_S2 + 2 * 64 * 64, WT_F2 = WT_S3 + 2 * 32 * 64, WT_F3 = WT_F2 + 2 * 64 * 64,
              WPK_ALL = WT_F3 + 2 * 32 * 64;
__host__ __device__ inline int wt_off(int L) { return L == 1 ? WT_S2 : (L == 2 ? WT_S3 : (L == 4 ? WT_F2 : WT_F3)); }

constexpr float kMomentum = 0.1f;
// head kernel modes / optimiser kernel modes
constexpr int kHeadFused = 0;       // forward + mean BCE + backward (sgpr_train_step)
constexpr int kHeadForward = 1;     // predictions only (sgpr_train_forward)
constexpr int kHeadBackward = 2;    // forward recomputed, backward from T.dpred (sgpr_train_backward)
constexpr int kApplyNone = 0;       // gradients only
constexpr int kApplyAll = 1;        // Adam update + running statistics
constexpr int kApplyRunning = 2;    // running statistics only (a train-mode forward without an optimiser step)           // nn.BatchNorm default (sg_net.py:52)

// layer index L: 0-2 xyz EdgeConv 1-3, 3-5 sem EdgeConv 1-3, 6 conv_end
__host__ __device__ inline int layer_cin(int L) { return L == 0 ? 3 : (L == 3 ? 12 : 64); }
__host__ __device__ inline int layer_cout(int L) { return (L == 2 || L == 5 || L == 6) ? 32 : 64; }
__host__ __device__ inline int conv_off(int L) {
    switch (L) {
        case 0: return P_S1W; case 1: return P_S2W; case 2: return P_S3W; case 3: return P_F1W;
        case 4: return P_F2W; case 5: return P_F3W; default: return P_ENDW;
    }
}
__host__ __device__ inline int conv_size(int L) { return L == 6 ? 32 * 64 : layer_cout(L) * 2 * layer_cin(L); }
__host__ __device__ inline int bn_off(int L) {          // offset of gamma within the BN block; beta follows at +C
    switch (L) {
        case 0: return 0; case 1: return 128; case 2: return 256; case 3: return 320;
        case 4: return 448; case 5: return 576; default: return 640;
    }
}
__host__ __device__ inline int wpk_off(int L) {
    switch (L) {
        case 1: return WPK_S2; case 2: return WPK_S3; case 3: return WPK_F1; case 4: return WPK_F2;
        case 5: return WPK_F3; default: return WPK_END;
    }
}

struct Segment {            // one contiguous slice of the parameter vector whose gradient arrives as per-CTA partials
    int off, size, count, stride;
    const float* part;
};
constexpr int kMaxSeg = 12;

struct TrainWs {
    int G;                  // pairs per step = graphs per side
    int S;                  // BatchNorm batches processed: 2 (one per side), or 1 in mirrored mode
    int mirrored;           // features_2[p] == features_1[p ^ 1] (process_batch's doubling, sg_net.py:324-331): side 2 holds
                            // the same graphs as side 1, hence the same batch statistics, activations and — summed over
                            // both roles of a graph — gradients; only side 1 is computed
    int N, k, KS;
    float eps;
    const float* f[2];      // features_1 / features_2   [G][15][N]
    const float* target;    // [G]  (fused step)
    const float* dpred;     // [G]  d loss / d prediction handed in by the caller's autograd (sgpr_train_backward)
    float* state;           // [STATE_TOTAL] parameters | running statistics
    float* wpk;             // [WPK_TOTAL]
    // per EdgeConv layer L = 0..5, every tensor [2*G][N][cout]  (side-major: sg = side*G + g)
    float* yext[6];         // extreme over the neighbours of the pre-BN activation
    float* a[6];            // A_i
    float* d[6];            // B_i - A_i
    float* sumy[6];         // sum_j y_ij
    float* gz[6];           // d loss / d z (z = BN output), written by the backward of the layer above
    uint8_t* enode[6];      // which neighbour achieved the extreme
    uint8_t* idx[6];        // [2*G][N][k] neighbour lists
    float* yend;            // [2*G][N][32] conv_end pre-BN
    float* gzend;           // [2*G][N][32]
    float* pooled;          // [2*G][32]
    float* att;             // [2*G][N]
    float* dpooled;         // [2*G][32]
    double* stats;          // [2][7][2][64]  sum y | sum y^2
    double* bsum;           // [2][7][2][64]  dbeta | dgamma
    float* pred;            // [G]
    float* losspart;        // [head grid]
    float* grads;           // [P_TOTAL]
    float* adam_m;
    float* adam_v;
    float* loss;            // [1]
    int head_grid;
    int nseg;
    Segment seg[kMaxSeg];
};

__device__ __forceinline__ double* stat_ptr(double* base, int side, int L) { return base + (side * 7 + L) * 128; }

// batch mean and 1/sqrt(var + eps) of channel c from the fp64 sums (biased variance, like nn.BatchNorm in train mode)
__device__ __forceinline__ void bn_coef(const double* st, int c, double e, float eps, float& mu, float& istd) {
    const double m = st[c] / e;
    double var = st[64 + c] / e - m * m;
    var = var > 0.0 ? var : 0.0;
    mu = static_cast<float>(m);
    istd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}
__device__ __forceinline__ float bn_act(float y, float mu, float istd, float gamma, float beta) {
    return lrelu(fmaf(__fmul_rn(__fsub_rn(y, mu), istd), gamma, beta));
}
__device__ __forceinline__ float slope_of(float x) { return x > 0.0f ? 1.0f : kSlope; }   // LeakyReLU'(z); sign(z) = sign(x)

__device__ __forceinline__ int pair_index_dev(int ci, int co, int CO) {       // pack.hpp::pair_index
    const int cpl = CO / 32;
    const int p = ci >> 1, r = ci & 1;
    const int lane = co / cpl, j = co % cpl;
    return p * 2 * CO + (j >> 1) * 128 + lane * (cpl >= 2 ? 4 : 2) + (j & 1) * 2 + r;
}

// ---- pack: natural conv weights -> channel-pair GEMM layout [cin][A half | B half] ------------------------------------
// element e of layer L's natural matrix (value v) -> its slots in the packed copies the GEMMs read
__device__ __forceinline__ void pack_element(float* __restrict__ wpk, int L, int e, float v) {
    if (L == 6) {                                   // conv_end [32][64]: out f, in c
        const int f = e >> 6, c = e & 63;
        wpk[wpk_off(6) + pair_index_dev(c, f, 32)] = v;
        return;
    }
    const int cin = layer_cin(L), cout = layer_cout(L);
    const int c = e / (2 * cin), col = e % (2 * cin);
    const int ci = col < cin ? col : col - cin;
    const int co = col < cin ? c : cout + c;
    wpk[wpk_off(L) + pair_index_dev(ci, co, 2 * cout)] = v;
    if (cin == 64)      // transposed halves for the backward: MA[c][ci] = Wa[c][ci], MB[c][ci] = Wb[c][ci]
        wpk[wt_off(L) + (col < cin ? 0 : cout * 64) + pair_index_dev(c, ci, 64)] = v;
}

// whole-state pack (after sgpr_train_set_state*); during training the optimiser kernel keeps the copies current
#ifndef SGPR_TRAIN_INST_ONLY   // non-template kernels live in train.cu only (train_inst.cu holds the per-NPL templates)
__global__ void __launch_bounds__(kThreads) sgpr_train_pack_kernel(const TrainWs T) {
    const int stride = gridDim.x * blockDim.x;
    for (int L = 1; L <= 6; ++L) {
        const float* w = T.state + conv_off(L);
        const int total = conv_size(L);
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) pack_element(T.wpk, L, e, w[e]);
    }
}
#endif  // !SGPR_TRAIN_INST_ONLY

// ---- shared-memory carve-outs ----------------------------------------------------------------------------------------
struct FwdSmem { int w, x, y, xx, idx, cnt, prm, stat, total; };
__host__ __device__ inline FwdSmem fwd_layout(int nmax, int ks) {
    FwdSmem L;
    int o = 0;
    L.w = o;    o += 64 * 128 * 4;
    L.x = o;    o += nmax * XS * 4;
    L.y = o;    o += nmax * YS * 4;
    L.xx = o;   o += nmax * 4;
    L.idx = o;  o += ((nmax * ks * 2 + 15) / 16) * 16;
    L.cnt = o;  o += ((nmax + 15) / 16) * 16;
    L.prm = o;  o += 4 * 64 * 4;
    L.stat = o; o += kWarps * 128 * 8;
    L.total = o;
    return L;
}

// BatchNorm terms of the layer below (mu | istd | gamma | beta per channel) for a CTA's (side, layer): once per CTA
__device__ __forceinline__ void load_prev_bn(const TrainWs& T, int Lprev, int side, float* sPrm, int tid) {
    if (tid < 64) {
        float mu, istd;
        bn_coef(stat_ptr(T.stats, side, Lprev), tid, static_cast<double>(T.G) * T.N * T.k, T.eps, mu, istd);
        sPrm[tid] = mu;
        sPrm[64 + tid] = istd;
        sPrm[128 + tid] = T.state[P_BN + bn_off(Lprev) + tid];
        sPrm[192 + tid] = T.state[P_BN + bn_off(Lprev) + 64 + tid];
    }
}

// fill the node-major input tile of EdgeConv layer l of branch br for graph (side, g); columns beyond cin stay zero.
// l > 0 needs load_prev_bn + a barrier beforehand.
__device__ __forceinline__ void fill_layer_input(const TrainWs& T, int l, int br, int side, int g, float* sX,
                                                 const float* sPrm, int tid) {
    const int N = T.N, L = br * 3 + l;
    if (l == 0) {
        const float* gin = T.f[side] + static_cast<size_t>(g) * kInCh * N;
        if (br == 0) {
            for (int n = tid; n < N; n += kThreads)
                *reinterpret_cast<float4*>(sX + n * XS) = make_float4(__ldg(gin + n), __ldg(gin + N + n), __ldg(gin + 2 * N + n), 0.0f);
        } else {
            for (int e = tid; e < N * kLabels; e += kThreads) {
                const int c = e / N, n = e - c * N;
                sX[n * XS + c] = __ldg(gin + (3 + c) * N + n);
            }
        }
    } else {
        const float4* yp = reinterpret_cast<const float4*>(T.yext[L - 1] + (static_cast<size_t>(side) * T.G + g) * N * 64);
#pragma unroll 4
        for (int e = tid; e < N * 16; e += kThreads) {
            const int n = e >> 4, c = (e & 15) * 4;
            const float4 y = __ldg(yp + e);
            float4 x;
            x.x = bn_act(y.x, sPrm[c], sPrm[64 + c], sPrm[128 + c], sPrm[192 + c]);
            x.y = bn_act(y.y, sPrm[c + 1], sPrm[65 + c], sPrm[129 + c], sPrm[193 + c]);
            x.z = bn_act(y.z, sPrm[c + 2], sPrm[66 + c], sPrm[130 + c], sPrm[194 + c]);
            x.w = bn_act(y.w, sPrm[c + 3], sPrm[67 + c], sPrm[131 + c], sPrm[195 + c]);
            *reinterpret_cast<float4*>(sX + n * XS + c) = x;
        }
    }
}

// gather over the neighbour lists for own rows: extreme / arg-extreme of A_j and the sums of y_ij = (A_j - A_i) + B_i.
// Per edge and channel only sum A_j and sum A_j^2 are accumulated; with D = B_i - A_i
//     sum_j y = SA + k D,      sum_j y^2 = SA2 + 2 D SA + k D^2
// (all terms are O(k) for BatchNorm-scaled features, so the expansion costs ~3 ulp of the sum).  The extreme is a max of
// sign-folded values (sign bit flipped where gamma < 0: min == max of the negated values, exact).
template <int COUT>
__device__ __forceinline__ void train_gather_rows(const float* __restrict__ sY, const uint16_t* __restrict__ sIdx, int KS,
                                                  int k, const float* __restrict__ gamma, float* __restrict__ yext,
                                                  float* __restrict__ ga, float* __restrict__ gd, float* __restrict__ gsum,
                                                  uint8_t* __restrict__ enode, int r0, int r1, int lane,
                                                  double (&acc1)[2], double (&acc2)[2]) {
    constexpr int CPL = COUT / 32;
    uint32_t flip[CPL];
#pragma unroll
    for (int p = 0; p < CPL; ++p) flip[p] = gamma[lane * CPL + p] >= 0.0f ? 0u : 0x80000000u;
    const float kf = static_cast<float>(k);
    const float* base = sY + lane * CPL;
    for (int i = r0; i < r1; ++i) {
        float best[CPL], sa[CPL], sa2[CPL];
        int bj[CPL];
#pragma unroll
        for (int p = 0; p < CPL; ++p) { best[p] = -INFINITY; sa[p] = 0.0f; sa2[p] = 0.0f; bj[p] = 0; }
        const uint16_t* list = sIdx + i * KS;
#pragma unroll 4
        for (int e = 0; e < k; ++e) {
            const int j = list[e];
            float v[CPL];
            if constexpr (CPL == 2) {
                const float2 t = *reinterpret_cast<const float2*>(base + j * YS);
                v[0] = t.x; v[1] = t.y;
            } else {
                v[0] = base[j * YS];
            }
#pragma unroll
            for (int p = 0; p < CPL; ++p) {
                sa[p] = __fadd_rn(sa[p], v[p]);
                sa2[p] = fmaf(v[p], v[p], sa2[p]);
                const float vs = __uint_as_float(__float_as_uint(v[p]) ^ flip[p]);
                bj[p] = vs > best[p] ? j : bj[p];
                best[p] = fmaxf(best[p], vs);
            }
        }
#pragma unroll
        for (int p = 0; p < CPL; ++p) {
            const size_t o = static_cast<size_t>(i) * COUT + lane * CPL + p;
            const float ai = base[i * YS + p], bi = base[i * YS + COUT + p];
            const float dv = __fsub_rn(bi, ai);
            const float ext = __uint_as_float(__float_as_uint(best[p]) ^ flip[p]);
            const float s1 = fmaf(kf, dv, sa[p]);
            const float s2 = fmaf(kf * dv, dv, fmaf(2.0f * dv, sa[p], sa2[p]));
            yext[o] = __fadd_rn(__fsub_rn(ext, ai), bi);
            ga[o] = ai;
            gd[o] = dv;
            gsum[o] = s1;
            enode[o] = static_cast<uint8_t>(bj[p]);
            acc1[p] += static_cast<double>(s1);
            acc2[p] += static_cast<double>(s2);
        }
    }
}

// xyz layer 1 in the reference's direct per-edge form W_a (x_j - x_i) + W_b x_i (metre-scale coordinates would lose
// ~5 bits in A_j - A_i; same choice as embed_kernel.cuh::xyz_rows).  Per edge only e_ij = W_a (x_j - x_i) is formed; the
// centre term c_i = W_b x_i joins per node: sum y = SE + k c, sum y^2 = SE2 + 2 c SE + k c^2.  sW: natural [64][6].
__device__ __forceinline__ void train_xyz_rows(const float* __restrict__ sX, const uint16_t* __restrict__ sIdx, int KS, int k,
                                               const float* __restrict__ sW, const float* __restrict__ gamma,
                                               float* __restrict__ yext, float* __restrict__ ga, float* __restrict__ gd,
                                               float* __restrict__ gsum, uint8_t* __restrict__ enode, int r0, int r1,
                                               int lane, double (&acc1)[2], double (&acc2)[2]) {
    float wa[2][3], wb[2][3];
    uint32_t flip[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int c = 2 * lane + p;
#pragma unroll
        for (int u = 0; u < 3; ++u) { wa[p][u] = sW[c * 6 + u]; wb[p][u] = sW[c * 6 + 3 + u]; }
        flip[p] = gamma[c] >= 0.0f ? 0u : 0x80000000u;
    }
    const float kf = static_cast<float>(k);
    for (int i = r0; i < r1; ++i) {
        const float4 xi = *reinterpret_cast<const float4*>(sX + i * XS);
        float best[2], se[2], se2[2];
        int bj[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) { best[p] = -INFINITY; se[p] = 0.0f; se2[p] = 0.0f; bj[p] = 0; }
        const uint16_t* list = sIdx + i * KS;
#pragma unroll 4
        for (int e = 0; e < k; ++e) {
            const int j = list[e];
            const float4 xj = *reinterpret_cast<const float4*>(sX + j * XS);
            const float d0 = __fsub_rn(xj.x, xi.x), d1 = __fsub_rn(xj.y, xi.y), d2 = __fsub_rn(xj.z, xi.z);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const float ev = fmaf(wa[p][2], d2, fmaf(wa[p][1], d1, __fmul_rn(wa[p][0], d0)));
                se[p] = __fadd_rn(se[p], ev);
                se2[p] = fmaf(ev, ev, se2[p]);
                const float vs = __uint_as_float(__float_as_uint(ev) ^ flip[p]);
                bj[p] = vs > best[p] ? j : bj[p];
                best[p] = fmaxf(best[p], vs);
            }
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const size_t o = static_cast<size_t>(i) * 64 + 2 * lane + p;
            const float av = fmaf(wa[p][2], xi.z, fmaf(wa[p][1], xi.y, __fmul_rn(wa[p][0], xi.x)));
            const float cb = fmaf(wb[p][2], xi.z, fmaf(wb[p][1], xi.y, __fmul_rn(wb[p][0], xi.x)));
            const float ext = __uint_as_float(__float_as_uint(best[p]) ^ flip[p]);
            const float s1 = fmaf(kf, cb, se[p]);
            const float s2 = fmaf(kf * cb, cb, fmaf(2.0f * cb, se[p], se2[p]));
            yext[o] = __fadd_rn(ext, cb);
            ga[o] = av;
            gd[o] = __fsub_rn(cb, av);
            gsum[o] = s1;
            enode[o] = static_cast<uint8_t>(bj[p]);
            acc1[p] += static_cast<double>(s1);
            acc2[p] += static_cast<double>(s2);
        }
    }
}

// CTA-wide reduction of per-thread fp64 channel sums (thread owns `cpl` channels lane*cpl+p; one row of 128 per warp:
// [which][channel]) followed by one fp64 atomicAdd per channel into the (side, layer) statistics.
__device__ __forceinline__ void flush_channel_sums(double* sStat, double* dst, const double (&acc1)[2], const double (&acc2)[2],
                                                   int cpl, int cout, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    __syncthreads();
    for (int p = 0; p < cpl; ++p) {
        sStat[warp * 128 + lane * cpl + p] = acc1[p];
        sStat[warp * 128 + 64 + lane * cpl + p] = acc2[p];
    }
    __syncthreads();
    if (tid < 128 && (tid & 63) < cout) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += sStat[w * 128 + tid];
        atomicAdd(dst + tid, s);
    }
    __syncthreads();
}

// =====================================================================================================================
// EdgeConv layer l (0..2) forward for both branches and both sides: grid = multiple of 4, CTA -> (branch, side), loops g
// =====================================================================================================================
template <int NPL, int TIES = 0>      // TIES: k-NN tie rule, as in sgpr_embed_kernel (0 = ATen CUDA topk, 1 = ATen CPU topk)
__global__ void __launch_bounds__(kThreads, (NPL <= 2) ? 2 : 1) sgpr_train_edge_fwd(const TrainWs T, int l) {
    constexpr int NMAX = 32 * NPL;
    SGPR_DYN_SMEM(smem);
    const FwdSmem S = fwd_layout(NMAX, T.KS);
    float* sW = reinterpret_cast<float*>(smem + S.w);
    float* sX = reinterpret_cast<float*>(smem + S.x);
    float* sY = reinterpret_cast<float*>(smem + S.y);
    float* sXX = reinterpret_cast<float*>(smem + S.xx);
    uint16_t* sIdx = reinterpret_cast<uint16_t*>(smem + S.idx);
    uint8_t* sCnt = smem + S.cnt;
    float* sPrm = reinterpret_cast<float*>(smem + S.prm);
    double* sStat = reinterpret_cast<double*>(smem + S.stat);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int br = blockIdx.x & 1, side = (blockIdx.x >> 1) % T.S;
    const int L = br * 3 + l;
    const int cin4 = (l == 0) ? (br ? 3 : 1) : 16, cout = layer_cout(L);
    const int N = T.N, k = T.k, KS = T.KS;
    const bool direct = (L == 0);

    {   // this CTA's layer matrix, once
        const float* src = direct ? T.state + P_S1W : T.wpk + wpk_off(L);
        const int n = direct ? 64 * 6 : layer_cin(L) * 2 * cout;
        for (int e = tid; e < n; e += kThreads) sW[e] = src[e];
    }
    for (int e = tid; e < NMAX * XS; e += kThreads) sX[e] = 0.0f;
    for (int e = tid; e < NMAX; e += kThreads) sXX[e] = 0.0f;
    if (l > 0) load_prev_bn(T, L - 1, side, sPrm, tid);
    __syncthreads();

    const int rpw = (N + kWarps - 1) / kWarps;
    const int w0 = min(N, warp * rpw), w1 = min(N, w0 + rpw);
    const float* gamma = T.state + P_BN + bn_off(L);
    double acc1[2] = {0.0, 0.0}, acc2[2] = {0.0, 0.0};

    for (int g = (blockIdx.x >> 1) / T.S; g < T.G; g += (gridDim.x >> 1) / T.S) {
        const size_t sg = static_cast<size_t>(side) * T.G + g;
        fill_layer_input(T, l, br, side, g, sX, sPrm, tid);
        __syncthreads();
        norms_rows(sX, sXX, cin4, w0, w1, lane);
        __syncthreads();
        // ---- front: distance rows -> k-NN selection -> per-node GEMM rows, for own rows ----
        for (int r0 = w0; r0 < w1; r0 += 8) {
            const int nr = min(8, w1 - r0);
            SGPR_NR_SWITCH(nr, (gram_rows<NPL, NPL, NR>(sX, sXX, sY, cin4, N, N, r0, lane)))
            __syncwarp();
            select_rows<NPL, TIES>(sY, sIdx, sCnt, nullptr, N, N, k, KS, 1, r0, nr, lane);
            __syncwarp();
            if (!direct) {
                if (cout == 64) { SGPR_NR_SWITCH(nr, (gemm_rows<NR, 4, 0>(sX, sW, sY, YS, nullptr, cin4, r0, lane))) }
                else            { SGPR_NR_SWITCH(nr, (gemm_rows<NR, 2, 0>(sX, sW, sY, YS, nullptr, cin4, r0, lane))) }
            }
        }
        __syncthreads();
        // ---- back: gather over the neighbour lists for own rows ----
        const size_t o = sg * N * cout;
        if (direct)
            train_xyz_rows(sX, sIdx, KS, k, sW, gamma, T.yext[L] + o, T.a[L] + o, T.d[L] + o, T.sumy[L] + o, T.enode[L] + o,
                           w0, w1, lane, acc1, acc2);
        else if (cout == 64)
            train_gather_rows<64>(sY, sIdx, KS, k, gamma, T.yext[L] + o, T.a[L] + o, T.d[L] + o, T.sumy[L] + o,
                                  T.enode[L] + o, w0, w1, lane, acc1, acc2);
        else
            train_gather_rows<32>(sY, sIdx, KS, k, gamma, T.yext[L] + o, T.a[L] + o, T.d[L] + o, T.sumy[L] + o,
                                  T.enode[L] + o, w0, w1, lane, acc1, acc2);
        uint8_t* gi = T.idx[L] + sg * N * k;
        for (int i = w0; i < w1; ++i)
            for (int e = lane; e < k; e += 32) gi[i * k + e] = static_cast<uint8_t>(sIdx[i * KS + e]);
        __syncthreads();
    }
    flush_channel_sums(sStat, stat_ptr(T.stats, side, L), acc1, acc2, cout / 32, cout, tid);
}

// =====================================================================================================================
// conv_end forward (sg_net.py:104-105 before the BatchNorm): y = W_end . cat(xyz3, sem3) per node; statistics
// grid = multiple of 2, CTA -> side, loops g
// =====================================================================================================================
template <int NPL>
__global__ void __launch_bounds__(kThreads, (NPL <= 2) ? 2 : 1) sgpr_train_end_fwd(const TrainWs T) {
    constexpr int NMAX = 32 * NPL;
    SGPR_DYN_SMEM(smem);
    float* sW = reinterpret_cast<float*>(smem);                 // [64][32] pair layout
    float* sX = sW + 64 * 32;                                   // [NMAX][XS] cat(xyz3, sem3)
    float* sO = sX + NMAX * XS;                                 // [NMAX][XS] (first 32 columns)
    float* sPrm = sO + NMAX * XS;                               // [4][64]
    double* sStat = reinterpret_cast<double*>(sPrm + 256);      // [kWarps][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int side = blockIdx.x % T.S;
    const int N = T.N;
    for (int e = tid; e < 64 * 32; e += kThreads) sW[e] = T.wpk[WPK_END + e];
    for (int e = tid; e < NMAX * XS; e += kThreads) sX[e] = 0.0f;
    if (tid < 64) {          // channel c < 32: xyz layer 3 (L = 2); c >= 32: sem layer 3 (L = 5)
        const int L = tid < 32 ? 2 : 5, c = tid & 31;
        float mu, istd;
        bn_coef(stat_ptr(T.stats, side, L), c, static_cast<double>(T.G) * N * T.k, T.eps, mu, istd);
        sPrm[tid] = mu;
        sPrm[64 + tid] = istd;
        sPrm[128 + tid] = T.state[P_BN + bn_off(L) + c];
        sPrm[192 + tid] = T.state[P_BN + bn_off(L) + 32 + c];
    }
    __syncthreads();
    const int rpw = (N + kWarps - 1) / kWarps;
    const int w0 = min(N, warp * rpw), w1 = min(N, w0 + rpw);
    double acc1[2] = {0.0, 0.0}, acc2[2] = {0.0, 0.0};
    for (int g = blockIdx.x / T.S; g < T.G; g += gridDim.x / T.S) {
        const size_t sg = static_cast<size_t>(side) * T.G + g;
        const float* y2 = T.yext[2] + sg * N * 32;
        const float* y5 = T.yext[5] + sg * N * 32;
#pragma unroll 4
        for (int e = tid; e < N * 64; e += kThreads) {
            const int n = e >> 6, c = e & 63;
            const float y = c < 32 ? y2[n * 32 + c] : y5[n * 32 + c - 32];
            sX[n * XS + c] = bn_act(y, sPrm[c], sPrm[64 + c], sPrm[128 + c], sPrm[192 + c]);
        }
        __syncthreads();
        for (int r0 = w0; r0 < w1; r0 += 8) {
            const int nr = min(8, w1 - r0);
            SGPR_NR_SWITCH(nr, (gemm_rows<NR, 1, 0>(sX, sW, sO, XS, nullptr, 16, r0, lane)))
        }
        __syncwarp();
        float* yo = T.yend + sg * N * 32;
        for (int i = w0; i < w1; ++i) {
            const float y = sO[i * XS + lane];
            yo[i * 32 + lane] = y;
            acc1[0] += static_cast<double>(y);
            acc2[0] += static_cast<double>(y) * static_cast<double>(y);
        }
        __syncthreads();
    }
    flush_channel_sums(sStat, stat_ptr(T.stats, side, 6), acc1, acc2, 1, 32, tid);
}

// =====================================================================================================================
// conv_end BatchNorm + LeakyReLU -> attention pooling (layers_batch.py:28-39) -> pair head (layers_batch.py:70-83,
// sg_net.py:131-136) -> loss -> head backward -> attention backward -> LeakyReLU of conv_end, in ONE kernel: after the
// conv_end statistics nothing on this stretch depends on another pair.  A CTA works on a GROUP = the two graphs of a
// pair: thread half h (128 threads) owns graph h for the attention stages, the whole CTA runs the head.
//   two-sided: group p = graphs (side 0, p) and (side 1, p), one ordered pair p
//   mirrored : group q = graphs 2q and 2q+1 of side 0, the two ordered pairs 2q = (2q, 2q+1) and 2q+1 = (2q+1, 2q);
//              each graph's pooled-vector gradient is the sum over its two roles.
// mode kHeadForward stops after the predictions (sgpr_train_forward); kHeadBackward takes d loss / d prediction from
// T.dpred instead of the BCE (sgpr_train_backward; the attention forward is simply recomputed from yend).
// Outputs: pred, att, pooled, dpooled (taps), gz_end = d loss / d (conv_end BN output) with its dbeta / dgamma sums;
// partial rows: part_head [grid][kHeadFloats], part_att [grid][1024].
// =====================================================================================================================
#ifndef SGPR_TRAIN_INST_ONLY   // non-template kernels live in train.cu only (train_inst.cu holds the per-NPL templates)
__global__ void __launch_bounds__(kThreads, 2) sgpr_train_pool_head_kernel(const TrainWs T, float* __restrict__ part_head,
                                                                         float* __restrict__ part_att, int mode) {
    SGPR_DYN_SMEM(smem);
    const int N = T.N, G = T.G;
    float* sE = reinterpret_cast<float*>(smem);            // [2][N][33]  node embeddings
    float* sYh = sE + 2 * N * 33;                          // [2][N][33]  conv_end BN-normalised pre-activations
    __shared__ float sPrm[2][4 * 32];
    __shared__ float sSum[2][32], sCtx[2][32], sPool[2][32], sDp[2][32], sDcbar[2][32], sV[2][32];
    __shared__ float sAtt[2][SGPR_MAX_NODES], sDsig[2][SGPR_MAX_NODES];
    __shared__ float e12[64];              // e1 | e2 (the V-block input cat(e1, e2), layers_batch.py:80)
    __shared__ float sP[512], sQ[512];
    __shared__ float sS[16], sNt[16], sHpre[16], sH[16], sDh[16], sDs[16];
    __shared__ float sDz;
    __shared__ double sRed[kWarps * 64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = tid >> 7, lt = tid & 127;
    const int t16 = tid >> 4, part16 = tid & 15;
    const int side = T.mirrored ? 0 : half;
    const float* watt = T.state + P_ATT;
    const float* W = T.state + P_NTNW;
    const float* V = T.state + P_NTNV;
    const float* nb = T.state + P_NTNB;
    const float* w1 = T.state + P_FC1W;
    const float* b1 = T.state + P_FC1B;
    const float* w2 = T.state + P_FC2W;
    const float* b2 = T.state + P_FC2B;
    float gW[64], gV[4], gA[4], gW1 = 0.0f, gNb = 0.0f, gB1 = 0.0f, gW2 = 0.0f, gB2 = 0.0f, loss = 0.0f;
#pragma unroll
    for (int m = 0; m < 64; ++m) gW[m] = 0.0f;
#pragma unroll
    for (int m = 0; m < 4; ++m) { gV[m] = 0.0f; gA[m] = 0.0f; }
    double accb = 0.0, accg = 0.0;                     // conv_end channel lt & 31 of this half's side
    if (lt < 32) {
        float mu, istd;
        bn_coef(stat_ptr(T.stats, side, 6), lt, static_cast<double>(G) * N, T.eps, mu, istd);
        sPrm[half][lt] = mu;
        sPrm[half][32 + lt] = istd;
        sPrm[half][64 + lt] = T.state[P_BN + bn_off(6) + lt];
        sPrm[half][96 + lt] = T.state[P_BN + bn_off(6) + 32 + lt];
    }
    __syncthreads();
    const int groups = T.mirrored ? G / 2 : G;
    const int npairs = T.mirrored ? 2 : 1;
    const float* prm = sPrm[half];
    float* hE = sE + half * N * 33;
    float* hYh = sYh + half * N * 33;

    for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const size_t sg = T.mirrored ? static_cast<size_t>(2 * grp + half) : static_cast<size_t>(half) * G + grp;
        // ---- conv_end BN + LeakyReLU ----
        {
            const float* y = T.yend + sg * N * 32;
#pragma unroll 4
            for (int e = lt; e < N * 32; e += 128) {
                const int n = e >> 5, c = e & 31;
                const float yh = __fmul_rn(__fsub_rn(__ldg(y + e), prm[c]), prm[32 + c]);
                hYh[n * 33 + c] = yh;
                hE[n * 33 + c] = lrelu(fmaf(yh, prm[64 + c], prm[96 + c]));
            }
        }
        __syncthreads();
        // ---- attention forward (layers_batch.py:35-38), one graph per thread half ----
        if (lt < 32) {
            float s = 0.0f;
            for (int n = 0; n < N; ++n) s = __fadd_rn(s, hE[n * 33 + lt]);
            sSum[half][lt] = s;
        }
        __syncthreads();
        if (lt < 32) {          // context = tanh(mean_n(E W)) = tanh((sum_n E) W / N)
            float s = 0.0f;
            for (int a = 0; a < 32; ++a) s = fmaf(sSum[half][a], watt[a * 32 + lt], s);
            sCtx[half][lt] = tanhf(s / static_cast<float>(N));
        }
        __syncthreads();
        if (lt < N) {
            float s = 0.0f;
            for (int a = 0; a < 32; ++a) s = fmaf(hE[lt * 33 + a], sCtx[half][a], s);
            const float av = sigmoidf_acc(s);
            sAtt[half][lt] = av;
            T.att[sg * N + lt] = av;
        }
        __syncthreads();
        if (lt < 32) {
            float s = 0.0f;
            for (int n = 0; n < N; ++n) s = fmaf(sAtt[half][n], hE[n * 33 + lt], s);
            sPool[half][lt] = s;
            sDp[half][lt] = 0.0f;
            T.pooled[sg * 32 + lt] = s;
        }
        __syncthreads();
        // ---- pair head: forward, loss, backward ----
        for (int q = 0; q < npairs; ++q) {
            const int p = T.mirrored ? 2 * grp + q : grp;       // pair index
            const int g1 = q, g2 = q ^ 1;                       // which half holds e1 / e2 of this pair
            if (tid < 64) e12[tid] = tid < 32 ? sPool[g1][tid] : sPool[g2][tid - 32];
            __syncthreads();
            const float* e1 = e12;
            const float* e2 = e12 + 32;
            for (int m = 0; m < 2; ++m) {                       // P[b*16+t] = sum_a e1[a] W[a][b][t]
                const int bt = tid + 256 * m;
                float acc = 0.0f;
#pragma unroll 8
                for (int a = 0; a < 32; ++a) acc = fmaf(e1[a], W[a * 512 + bt], acc);
                sP[bt] = acc;
            }
            __syncthreads();
            {   // s[t] = sum_b P[b][t] e2[b] + V[t] . cat + bias[t];  thread (t16, part16) sums a slice
                float bil = fmaf(sP[(2 * part16 + 1) * 16 + t16], e2[2 * part16 + 1], __fmul_rn(sP[(2 * part16) * 16 + t16], e2[2 * part16]));
                float blk = 0.0f;
#pragma unroll
                for (int u = 0; u < 4; ++u) blk = fmaf(V[t16 * 64 + part16 * 4 + u], e12[part16 * 4 + u], blk);
                bil = group16_sum(bil);
                blk = group16_sum(blk);
                if (part16 == 0) {
                    const float s = __fadd_rn(__fadd_rn(bil, blk), nb[t16]);
                    sS[t16] = s;
                    sNt[t16] = fmaxf(s, 0.0f);
                }
            }
            __syncthreads();
            {
                float h = group16_sum(__fmul_rn(sNt[part16], w1[t16 * 16 + part16]));
                if (part16 == 0) {
                    h = __fadd_rn(h, b1[t16]);
                    sHpre[t16] = h;
                    sH[t16] = fmaxf(h, 0.0f);
                }
            }
            __syncthreads();
            if (warp == 0) {
                float z = lane < 16 ? __fmul_rn(sH[lane], w2[lane]) : 0.0f;
                z = warp_sum(z);
                if (lane == 0) {
                    const float pr = sigmoidf_acc(__fadd_rn(z, b2[0]));
                    T.pred[p] = pr;
                    if (mode == kHeadFused) {
                        const float tg = T.target[p];
                        // torch.nn.functional.binary_cross_entropy clamps both logs at -100
                        loss += -(tg * fmaxf(logf(pr), -100.0f) + (1.0f - tg) * fmaxf(logf(1.0f - pr), -100.0f));
                        sDz = (pr - tg) / static_cast<float>(G);           // d mean-BCE / d (pre-sigmoid score)
                    } else if (mode == kHeadBackward) {
                        sDz = T.dpred[p] * pr * (1.0f - pr);               // through the sigmoid (sg_net.py:136)
                    }
                }
            }
            __syncthreads();
            if (mode == kHeadForward) continue;                            // uniform: the whole CTA skips the backward
            const float dz = sDz;
            if (tid < 16) sDh[tid] = sHpre[tid] > 0.0f ? dz * w2[tid] : 0.0f;
            __syncthreads();
            if (tid < 16) {
                float s = 0.0f;
                for (int u = 0; u < 16; ++u) s = fmaf(sDh[u], w1[u * 16 + tid], s);
                sDs[tid] = sS[tid] > 0.0f ? s : 0.0f;
            }
            __syncthreads();
            const float q0 = e2[tid >> 4] * sDs[part16];               // q[b*16+t] = e2[b] ds[t], bt = tid
            const float q1 = e2[16 + (tid >> 4)] * sDs[part16];        // bt = tid + 256
            sQ[tid] = q0;
            sQ[256 + tid] = q1;
            // ---- head parameter gradients ----
#pragma unroll
            for (int m = 0; m < 64; ++m) gW[m] = fmaf(e1[m >> 1], (m & 1) ? q1 : q0, gW[m]);    // flat index tid + 256 m
#pragma unroll
            for (int m = 0; m < 4; ++m) { const int e = tid + 256 * m; gV[m] = fmaf(sDs[e >> 6], e12[e & 63], gV[m]); }
            gW1 = fmaf(sDh[t16], sNt[part16], gW1);
            if (tid < 16) {
                gNb += sDs[tid];
                gB1 += sDh[tid];
                gW2 = fmaf(dz, sH[tid], gW2);
            }
            if (tid == 0) gB2 += dz;
            __syncthreads();
            // ---- d loss / d pooled vectors, added to the graph that played the role ----
            for (int u = 0; u < 4; ++u) {                        // de1[a] = sum_bt W[a][bt] q[bt] + sum_t ds[t] V[t][a]
                const int a = warp * 4 + u;
                float s = 0.0f;
#pragma unroll 4
                for (int i = 0; i < 16; ++i) s = fmaf(W[a * 512 + lane + 32 * i], sQ[lane + 32 * i], s);
                s = warp_sum(s);
                if (lane == 0) {
                    for (int t = 0; t < 16; ++t) s = fmaf(sDs[t], V[t * 64 + a], s);
                    T.dpooled[static_cast<size_t>(p) * 32 + a] = s;
                    sDp[g1][a] = __fadd_rn(sDp[g1][a], s);
                }
            }
            __syncthreads();                                     // (mirrored: g2 of this pair was g1's array a moment ago)
            if (tid < 32) {                                      // de2[b] = sum_t P[b][t] ds[t] + sum_t ds[t] V[t][32+b]
                float s = 0.0f;
                for (int t = 0; t < 16; ++t) s = fmaf(sP[tid * 16 + t], sDs[t], s);
                for (int t = 0; t < 16; ++t) s = fmaf(sDs[t], V[t * 64 + 32 + tid], s);
                T.dpooled[(static_cast<size_t>(G) + p) * 32 + tid] = s;
                sDp[g2][tid] = __fadd_rn(sDp[g2][tid], s);
            }
            __syncthreads();
        }
        if (mode == kHeadForward) continue;
        // ---- attention backward, one graph per thread half ----
        if (lt < N) {
            float s = 0.0f;
            for (int a = 0; a < 32; ++a) s = fmaf(hE[lt * 33 + a], sDp[half][a], s);
            const float av = sAtt[half][lt];
            sDsig[half][lt] = s * av * (1.0f - av);
        }
        __syncthreads();
        if (lt < 32) {
            float s = 0.0f;
            for (int n = 0; n < N; ++n) s = fmaf(sDsig[half][n], hE[n * 33 + lt], s);
            const float c = sCtx[half][lt];
            sDcbar[half][lt] = s * (1.0f - c * c) / static_cast<float>(N);
        }
        __syncthreads();
        if (lt < 32) {
            float s = 0.0f;
            for (int b = 0; b < 32; ++b) s = fmaf(sDcbar[half][b], watt[lt * 32 + b], s);
            sV[half][lt] = s;
        }
        __syncthreads();
        {
            float* gzo = T.gzend + sg * N * 32;
            for (int e = lt; e < N * 32; e += 128) {
                const int n = e >> 5, c = e & 31;
                const float de = fmaf(sAtt[half][n], sDp[half][c], fmaf(sDsig[half][n], sCtx[half][c], sV[half][c]));
                const float gzv = de * slope_of(hE[n * 33 + c]);
                gzo[e] = gzv;
                accb += static_cast<double>(gzv);
                accg += static_cast<double>(gzv) * static_cast<double>(hYh[n * 33 + c]);
            }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {                            // d attention.weight_matrix[a][b] += esum[a] dcbar[b], both graphs
            const int e = tid + 256 * m;
            gA[m] = fmaf(sSum[0][e >> 5], sDcbar[0][e & 31], gA[m]);
            gA[m] = fmaf(sSum[1][e >> 5], sDcbar[1][e & 31], gA[m]);
        }
        __syncthreads();
    }
    if (mode == kHeadForward) return;
    float* row = part_head + static_cast<size_t>(blockIdx.x) * kHeadFloats;
#pragma unroll
    for (int m = 0; m < 64; ++m) row[tid + 256 * m] = gW[m];
#pragma unroll
    for (int m = 0; m < 4; ++m) row[16384 + tid + 256 * m] = gV[m];
    row[16384 + 1024 + 16 + tid] = gW1;
    if (tid < 16) {
        row[16384 + 1024 + tid] = gNb;
        row[16384 + 1024 + 16 + 256 + tid] = gB1;
        row[16384 + 1024 + 16 + 256 + 16 + tid] = gW2;
    }
    if (tid == 0) {
        row[kHeadFloats - 1] = gB2;
        T.losspart[blockIdx.x] = loss;
    }
    float* arow = part_att + static_cast<size_t>(blockIdx.x) * 1024;
#pragma unroll
    for (int m = 0; m < 4; ++m) arow[tid + 256 * m] = gA[m];
    // dbeta / dgamma of the conv_end BatchNorm: warps 0-3 hold graph half 0, warps 4-7 half 1
    sRed[warp * 64 + lane] = accb;
    sRed[warp * 64 + 32 + lane] = accg;
    __syncthreads();
    if (tid < 128) {
        const int sd = tid >> 6, which = (tid >> 5) & 1, c = tid & 31;
        if (T.mirrored) {
            if (sd == 0) {
                double s = 0.0;
                for (int w = 0; w < kWarps; ++w) s += sRed[w * 64 + which * 32 + c];
                atomicAdd(stat_ptr(T.bsum, 0, 6) + which * 64 + c, s);
            }
        } else {
            double s = 0.0;
            for (int w = 4 * sd; w < 4 * sd + 4; ++w) s += sRed[w * 64 + which * 32 + c];
            atomicAdd(stat_ptr(T.bsum, sd, 6) + which * 64 + c, s);
        }
    }
}
#endif  // !SGPR_TRAIN_INST_ONLY

// =====================================================================================================================
// conv_end backward: BatchNorm backward with the complete dbeta/dgamma, d cat(xyz3, sem3), dW_end, and the gz of the
// two third EdgeConv layers (+ their dbeta/dgamma sums).  grid = multiple of 2, CTA -> side.  part: [grid][2048].
// =====================================================================================================================
template <int NPL>
__global__ void __launch_bounds__(kThreads, (NPL <= 2) ? 2 : 1) sgpr_train_end_bwd(const TrainWs T, float* __restrict__ part) {
    constexpr int NMAX = 32 * NPL;
    constexpr int DS = 36;                                        // row stride of the dy tile
    SGPR_DYN_SMEM(smem);
    float* sWn = reinterpret_cast<float*>(smem);                // [32][64] natural
    float* sX = sWn + 32 * 64;                                  // [NMAX][XS] cat(xyz3, sem3)
    float* sDy = sX + NMAX * XS;                                // [NMAX][DS]
    float* sPrm = sDy + NMAX * DS;                              // [4][64] previous-layer BN terms
    float* sEnd = sPrm + 256;                                   // [4][32]: mu, istd, s, (unused) | dbeta/e, dgamma/e
    double* sRed = reinterpret_cast<double*>(sEnd + 192);       // [4][128]
    const int tid = threadIdx.x;
    const int side = blockIdx.x % T.S;
    const int N = T.N;
    const double e_end = static_cast<double>(T.G) * N;
    for (int e = tid; e < 32 * 64; e += kThreads) sWn[e] = T.state[P_ENDW + e];
    for (int e = tid; e < NMAX * XS; e += kThreads) sX[e] = 0.0f;
    if (tid < 64) {
        const int L = tid < 32 ? 2 : 5, c = tid & 31;
        float mu, istd;
        bn_coef(stat_ptr(T.stats, side, L), c, static_cast<double>(T.G) * N * T.k, T.eps, mu, istd);
        sPrm[tid] = mu;
        sPrm[64 + tid] = istd;
        sPrm[128 + tid] = T.state[P_BN + bn_off(L) + c];
        sPrm[192 + tid] = T.state[P_BN + bn_off(L) + 32 + c];
    }
    if (tid < 32) {
        float mu, istd;
        bn_coef(stat_ptr(T.stats, side, 6), tid, e_end, T.eps, mu, istd);
        const double* bs = stat_ptr(T.bsum, side, 6);
        sEnd[tid] = mu;
        sEnd[32 + tid] = istd;
        sEnd[64 + tid] = T.state[P_BN + bn_off(6) + tid] * istd;
        sEnd[96 + tid] = static_cast<float>(bs[tid] / e_end);
        sEnd[128 + tid] = static_cast<float>(bs[64 + tid] / e_end);
    }
    __syncthreads();
    float gW[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) gW[m] = 0.0f;
    double accb = 0.0, accg = 0.0;                     // channel tid & 63 of cat
    const int f0 = 2 * (tid >> 4), c0 = 4 * (tid & 15);
    for (int g = blockIdx.x / T.S; g < T.G; g += gridDim.x / T.S) {
        const size_t sg = static_cast<size_t>(side) * T.G + g;
        const float* y2 = T.yext[2] + sg * N * 32;
        const float* y5 = T.yext[5] + sg * N * 32;
#pragma unroll 4
        for (int e = tid; e < N * 64; e += kThreads) {
            const int n = e >> 6, c = e & 63;
            const float y = c < 32 ? y2[n * 32 + c] : y5[n * 32 + c - 32];
            sX[n * XS + c] = bn_act(y, sPrm[c], sPrm[64 + c], sPrm[128 + c], sPrm[192 + c]);
        }
        const float* ye = T.yend + sg * N * 32;
        const float* gze = T.gzend + sg * N * 32;
#pragma unroll 4
        for (int e = tid; e < N * 32; e += kThreads) {
            const int n = e >> 5, c = e & 31;
            const float yh = __fmul_rn(__fsub_rn(ye[e], sEnd[c]), sEnd[32 + c]);
            sDy[n * DS + c] = sEnd[64 + c] * (gze[e] - sEnd[96 + c] - yh * sEnd[128 + c]);
        }
        __syncthreads();
        float* gz2 = T.gz[2] + sg * N * 32;
        float* gz5 = T.gz[5] + sg * N * 32;
        {   // d cat[n][c] = sum_f dy[n][f] W[f][c]: a thread owns channel c = tid & 63 and walks nodes four at a time
            const int c = tid & 63;
            for (int n0 = (tid >> 6) * 4; n0 < N; n0 += 16) {
                float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                int nn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) nn[u] = min(n0 + u, N - 1);
#pragma unroll
                for (int f4 = 0; f4 < 8; ++f4) {
                    const float w0 = sWn[(4 * f4) * 64 + c], w1 = sWn[(4 * f4 + 1) * 64 + c];
                    const float w2 = sWn[(4 * f4 + 2) * 64 + c], w3 = sWn[(4 * f4 + 3) * 64 + c];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 dy = *reinterpret_cast<const float4*>(sDy + nn[u] * DS + 4 * f4);
                        s[u] = fmaf(dy.w, w3, fmaf(dy.z, w2, fmaf(dy.y, w1, fmaf(dy.x, w0, s[u]))));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = n0 + u;
                    if (n < N) {
                        const float gzv = s[u] * slope_of(sX[n * XS + c]);
                        const float yprev = c < 32 ? y2[n * 32 + c] : y5[n * 32 + c - 32];
                        const float yh = __fmul_rn(__fsub_rn(yprev, sPrm[c]), sPrm[64 + c]);
                        if (c < 32) gz2[n * 32 + c] = gzv; else gz5[n * 32 + c - 32] = gzv;
                        accb += static_cast<double>(gzv);
                        accg += static_cast<double>(gzv) * static_cast<double>(yh);
                    }
                }
            }
        }
        for (int n = 0; n < N; ++n) {                              // dW_end[f][c] += dy[n][f] cat[n][c]
            const float2 dy = *reinterpret_cast<const float2*>(sDy + n * DS + f0);
            const float4 x = *reinterpret_cast<const float4*>(sX + n * XS + c0);
            gW[0] = fmaf(dy.x, x.x, gW[0]); gW[1] = fmaf(dy.x, x.y, gW[1]); gW[2] = fmaf(dy.x, x.z, gW[2]); gW[3] = fmaf(dy.x, x.w, gW[3]);
            gW[4] = fmaf(dy.y, x.x, gW[4]); gW[5] = fmaf(dy.y, x.y, gW[5]); gW[6] = fmaf(dy.y, x.z, gW[6]); gW[7] = fmaf(dy.y, x.w, gW[7]);
        }
        __syncthreads();
    }
    float* row = part + static_cast<size_t>(blockIdx.x) * 2048;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) row[(f0 + u) * 64 + c0 + v] = gW[u * 4 + v];
    sRed[(tid >> 6) * 128 + (tid & 63)] = accb;
    sRed[(tid >> 6) * 128 + 64 + (tid & 63)] = accg;
    __syncthreads();
    if (tid < 128) {
        const int which = tid >> 6, c = tid & 63;
        double s = 0.0;
        for (int w = 0; w < 4; ++w) s += sRed[w * 128 + tid];
        atomicAdd(stat_ptr(T.bsum, side, c < 32 ? 2 : 5) + which * 64 + (c & 31), s);
    }
}

// =====================================================================================================================
// EdgeConv layer l backward for both branches and both sides (grid = multiple of 4, CTA -> (branch, side), loops g).
// part: [grid/2 per branch][cout * 2 cin] conv-weight partials in the state_dict layout.
// =====================================================================================================================
struct BwdSmem { int x, gz, d, da, wh, en, adj, ib, rl, c, prm, red, total; };
__host__ __device__ inline BwdSmem bwd_layout(int nmax, int ks) {
    BwdSmem L;
    int o = 0;
    L.x = o;    o += nmax * XS * 4;
    L.gz = o;   o += nmax * XS * 4;
    L.d = o;    o += nmax * XS * 4;
    L.da = o;   o += nmax * XS * 4;
    L.wh = o;   o += 64 * 64 * 4;                        // one transposed half (MA, then MB) for the dX GEMM
    L.en = o;   o += nmax * 64;
    L.adj = o;  o += nmax * 16;                          // 128-bit neighbour mask per node
    L.ib = o;   o += ((nmax * ks + 15) / 16) * 16;       // this graph's neighbour lists (bytes)
    L.rl = o;   o += nmax * nmax;                        // reverse lists, nmax entries reserved per node
    L.c = o;    o += 4 * 64 * 4;
    L.prm = o;  o += 4 * 64 * 4;
    L.red = o;  o += kWarps * 128 * 8;
    L.total = o;
    return L;
}

// dWa[c][ci] += sum_n dA[n][c] x[n][ci], dWb likewise with dB: thread tile CPT output channels x 4 input channels
template <int CPT>
__device__ __forceinline__ void accum_dw(const float* __restrict__ sX, const float* __restrict__ sDA,
                                         const float* __restrict__ sDB, int N, int c0, int ci0, float (&ga)[16], float (&gb)[16]) {
    for (int n = 0; n < N; ++n) {
        const float4 x = *reinterpret_cast<const float4*>(sX + n * XS + ci0);
        float da[CPT], db[CPT];
        if constexpr (CPT == 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(sDA + n * XS + c0), b4 = *reinterpret_cast<const float4*>(sDB + n * XS + c0);
            da[0] = a4.x; da[1] = a4.y; da[2] = a4.z; da[3] = a4.w;
            db[0] = b4.x; db[1] = b4.y; db[2] = b4.z; db[3] = b4.w;
        } else if constexpr (CPT == 2) {
            const float2 a2 = *reinterpret_cast<const float2*>(sDA + n * XS + c0), b2 = *reinterpret_cast<const float2*>(sDB + n * XS + c0);
            da[0] = a2.x; da[1] = a2.y; db[0] = b2.x; db[1] = b2.y;
        } else {
            da[0] = sDA[n * XS + c0]; db[0] = sDB[n * XS + c0];
        }
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            ga[u * 4 + 0] = fmaf(da[u], x.x, ga[u * 4 + 0]); ga[u * 4 + 1] = fmaf(da[u], x.y, ga[u * 4 + 1]);
            ga[u * 4 + 2] = fmaf(da[u], x.z, ga[u * 4 + 2]); ga[u * 4 + 3] = fmaf(da[u], x.w, ga[u * 4 + 3]);
            gb[u * 4 + 0] = fmaf(db[u], x.x, gb[u * 4 + 0]); gb[u * 4 + 1] = fmaf(db[u], x.y, gb[u * 4 + 1]);
            gb[u * 4 + 2] = fmaf(db[u], x.z, gb[u * 4 + 2]); gb[u * 4 + 3] = fmaf(db[u], x.w, gb[u * 4 + 3]);
        }
    }
}

// dX rows += In rows x M for 8 rows starting at r0 (rows beyond r1 repeat the last one; the caller drops them).
// In: node-major tile, c4n*4 channels; M: [c4n*4][64] in the channel-pair layout; a lane owns outputs 2*lane, 2*lane+1 and
// keeps the even-/odd-channel partial sums of each in the two halves of a float2 (FFMA2).
__device__ __forceinline__ void dx_accum(const float* __restrict__ sIn, const float* __restrict__ sM, float2 (&acc)[8][2],
                                         int c4n, int r0, int r1, int lane) {
    const float* px[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) px[n] = sIn + min(r0 + n, r1 - 1) * XS;
#pragma unroll 2
    for (int c4 = 0; c4 < c4n; ++c4) {
        const float4 w0 = *reinterpret_cast<const float4*>(sM + (2 * c4) * 128 + lane * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(sM + (2 * c4 + 1) * 128 + lane * 4);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float4 x = *reinterpret_cast<const float4*>(px[n] + 4 * c4);
            acc[n][0] = ffma2(make_float2(x.x, x.y), make_float2(w0.x, w0.y), acc[n][0]);
            acc[n][1] = ffma2(make_float2(x.x, x.y), make_float2(w0.z, w0.w), acc[n][1]);
            acc[n][0] = ffma2(make_float2(x.z, x.w), make_float2(w1.x, w1.y), acc[n][0]);
            acc[n][1] = ffma2(make_float2(x.z, x.w), make_float2(w1.z, w1.w), acc[n][1]);
        }
    }
}

template <int NPL>
__global__ void __launch_bounds__(kThreads, (NPL <= 2) ? 2 : 1) sgpr_train_edge_bwd(const TrainWs T, int l, float* __restrict__ part0,
                                                                                   float* __restrict__ part1) {
    constexpr int NMAX = 32 * NPL;
    constexpr int SLOTS = 4 * NPL;                    // nodes per warp
    constexpr int PASSES = (NPL == 4) ? 2 : 1;        // 8-row passes of the dX GEMM per warp
    SGPR_DYN_SMEM(smem);
    const BwdSmem S = bwd_layout(NMAX, T.KS);
    float* sX = reinterpret_cast<float*>(smem + S.x);
    float* sGZ = reinterpret_cast<float*>(smem + S.gz);         // gz, later dB
    float* sD = reinterpret_cast<float*>(smem + S.d);
    float* sDA = reinterpret_cast<float*>(smem + S.da);         // S1, then dA
    float* sWh = reinterpret_cast<float*>(smem + S.wh);
    uint8_t* sEN = smem + S.en;
    uint32_t* sAdj = reinterpret_cast<uint32_t*>(smem + S.adj); // [NMAX][4]
    uint8_t* sIb = smem + S.ib;
    uint8_t* sRl = smem + S.rl;
    float* sC = reinterpret_cast<float*>(smem + S.c);           // s | q | r | (spare)
    float* sPrm = reinterpret_cast<float*>(smem + S.prm);       // previous layer: mu | istd | gamma | beta
    double* sRed = reinterpret_cast<double*>(smem + S.red);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int br = blockIdx.x & 1, side = (blockIdx.x >> 1) % T.S;
    const int L = br * 3 + l;
    const int cin = layer_cin(L), cout = layer_cout(L), cpl = cout / 32;
    const int N = T.N, k = T.k;
    const double e_edge = static_cast<double>(T.G) * N * k;

    for (int e = tid; e < NMAX * XS; e += kThreads) sX[e] = 0.0f;
    if (tid < cout) {
        float mu, istd;
        bn_coef(stat_ptr(T.stats, side, L), tid, e_edge, T.eps, mu, istd);
        const double* bs = stat_ptr(T.bsum, side, L);
        const float s = T.state[P_BN + bn_off(L) + tid] * istd;
        const float pterm = s * static_cast<float>(bs[tid] / e_edge);
        const float q = s * static_cast<float>(bs[64 + tid] / e_edge) * istd;
        sC[tid] = s;
        sC[64 + tid] = q;
        sC[128 + tid] = pterm - q * mu;
    }
    if (l > 0) load_prev_bn(T, L - 1, side, sPrm, tid);
    __syncthreads();

    float ga[16], gb[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) { ga[m] = 0.0f; gb[m] = 0.0f; }
    double accb[2] = {0.0, 0.0}, accg[2] = {0.0, 0.0};          // previous layer's channels 2*lane, 2*lane+1
    const int rpw = (N + kWarps - 1) / kWarps;
    const int w0 = min(N, warp * rpw), w1 = min(N, w0 + rpw);
    const float kf = static_cast<float>(k);

    for (int g = (blockIdx.x >> 1) / T.S; g < T.G; g += (gridDim.x >> 1) / T.S) {
        const size_t sg = static_cast<size_t>(side) * T.G + g;
        const size_t o = sg * N * cout;
        fill_layer_input(T, l, br, side, g, sX, sPrm, tid);
        {
            const float4* gz = reinterpret_cast<const float4*>(T.gz[L] + o);
            const float4* gd = reinterpret_cast<const float4*>(T.d[L] + o);
            const uint32_t* en = reinterpret_cast<const uint32_t*>(T.enode[L] + o);
            const int q4 = cout >> 2;                        // float4 groups per node
#pragma unroll 4
            for (int e = tid; e < N * q4; e += kThreads) {
                const int n = e / q4, c = (e - n * q4) * 4;
                *reinterpret_cast<float4*>(sGZ + n * XS + c) = __ldg(gz + e);
                *reinterpret_cast<float4*>(sD + n * XS + c) = __ldg(gd + e);
                *reinterpret_cast<uint32_t*>(sEN + n * 64 + c) = __ldg(en + e);
                *reinterpret_cast<float4*>(sDA + n * XS + c) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            {   // this graph's neighbour lists, N*k bytes
                const uint8_t* gb8 = T.idx[L] + sg * N * k;
                if (((N * k) & 3) == 0) {
                    const uint32_t* gi = reinterpret_cast<const uint32_t*>(gb8);
                    for (int e = tid; e < (N * k) / 4; e += kThreads) reinterpret_cast<uint32_t*>(sIb)[e] = __ldg(gi + e);
                } else {
                    for (int e = tid; e < N * k; e += kThreads) sIb[e] = __ldg(gb8 + e);
                }
            }
        }
        __syncthreads();
        // neighbour bit masks, one warp per node (lanes = list entries, OR-reduced across the warp)
        for (int i = warp; i < N; i += kWarps) {
            uint32_t m[NPL];
#pragma unroll
            for (int w = 0; w < NPL; ++w) m[w] = 0u;
            for (int e = lane; e < k; e += 32) {
                const int j = sIb[i * k + e];
#pragma unroll
                for (int w = 0; w < NPL; ++w) m[w] |= (j >> 5) == w ? (1u << (j & 31)) : 0u;
            }
#pragma unroll
            for (int w = 0; w < NPL; ++w) {
                const uint32_t all = __reduce_or_sync(0xffffffffu, m[w]);
                if (lane == 0) sAdj[4 * i + w] = all;
            }
        }
        __syncthreads();
        // ---- reverse neighbourhoods of own nodes (n = warp, warp + 8, ...): bit i of ballot word q is set when node
        // 32 q + i lists n; the set lanes write themselves into n's reverse list (ascending).  Then S1 (scatter by the
        // extreme neighbour, one thread per channel: fixed order) runs beside S2 (gather over the reverse lists). ----
        int degr[SLOTS];
        float s2r[SLOTS][2], areg[SLOTS][2], syreg[SLOTS][2];
        {
            const float* ga_ = T.a[L] + o;
            const float* gs_ = T.sumy[L] + o;
#pragma unroll
            for (int slot = 0; slot < SLOTS; ++slot) {
                const int n = warp + kWarps * slot;
                int base = 0;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int i = 32 * q + lane;
                    const bool lists = (n < N && i < N) ? ((sAdj[4 * i + (n >> 5)] >> (n & 31)) & 1u) != 0u : false;
                    const uint32_t mask = __ballot_sync(0xffffffffu, lists);
                    if (lists) sRl[n * NMAX + base + __popc(mask & ((1u << lane) - 1u))] = static_cast<uint8_t>(i);
                    base += __popc(mask);
                }
                degr[slot] = base;
#pragma unroll
                for (int p = 0; p < 2; ++p) {                   // operands of the per-node terms, in flight during S2
                    const bool live = n < N && p < cpl;
                    areg[slot][p] = live ? __ldg(ga_ + n * cout + lane * cpl + p) : 0.0f;
                    syreg[slot][p] = live ? __ldg(gs_ + n * cout + lane * cpl + p) : 0.0f;
                }
            }
        }
        __syncwarp();                                           // a node's reverse list is written and read by one warp
        if (tid < cout) {
            for (int i = 0; i < N; ++i) {
                const int n = sEN[i * 64 + tid];
                sDA[n * XS + tid] = __fadd_rn(sDA[n * XS + tid], sGZ[i * XS + tid]);
            }
        }
        if (cpl == 2) {
#pragma unroll
            for (int slot = 0; slot < SLOTS; ++slot) {
                const uint8_t* rl = sRl + (warp + kWarps * slot) * NMAX;
                float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 4
                for (int t = 0; t < degr[slot]; ++t) {
                    const float2 dv = *reinterpret_cast<const float2*>(sD + rl[t] * XS + 2 * lane);
                    a0 = __fadd_rn(a0, dv.x); a1 = __fadd_rn(a1, dv.y);
                }
                s2r[slot][0] = a0; s2r[slot][1] = a1;
            }
        } else {
#pragma unroll
            for (int slot = 0; slot < SLOTS; ++slot) {
                const uint8_t* rl = sRl + (warp + kWarps * slot) * NMAX;
                float a0 = 0.0f;
#pragma unroll 4
                for (int t = 0; t < degr[slot]; ++t) a0 = __fadd_rn(a0, sD[rl[t] * XS + lane]);
                s2r[slot][0] = a0; s2r[slot][1] = 0.0f;
            }
        }
        __syncthreads();                                        // S1 complete (and every read of gz by the scatter)
#pragma unroll
        for (int slot = 0; slot < SLOTS; ++slot) {
            const int n = warp + kWarps * slot;
            if (n < N) {
                const float deg = static_cast<float>(degr[slot]);
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    if (p < cpl) {
                        const int c = lane * cpl + p;
                        const float s = sC[c], q = sC[64 + c], r = sC[128 + c];
                        const float t = s * sGZ[n * XS + c] - kf * r - q * syreg[slot][p];
                        sDA[n * XS + c] = -t + s * sDA[n * XS + c] - deg * (r + q * areg[slot][p]) - q * s2r[slot][p];
                        sGZ[n * XS + c] = t;                    // dB
                    }
                }
            }
        }
        __syncthreads();
        // ---- dX = dA Wa + dB Wb -> gz of the layer below (+ its dbeta / dgamma sums) ----
        if (l > 0) {
            float2 acc[PASSES][8][2];
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps)
#pragma unroll
                for (int n = 0; n < 8; ++n) { acc[ps][n][0] = make_float2(0.0f, 0.0f); acc[ps][n][1] = make_float2(0.0f, 0.0f); }
            for (int half = 0; half < 2; ++half) {
                if (half) __syncthreads();                      // every warp is done with the first matrix
                const float4* src = reinterpret_cast<const float4*>(T.wpk + wt_off(L) + half * cout * 64);
                for (int e = tid; e < cout * 16; e += kThreads) reinterpret_cast<float4*>(sWh)[e] = __ldg(src + e);
                __syncthreads();
                const float* sIn = half ? sGZ : sDA;
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps)
                    if (w0 + 8 * ps < w1) dx_accum(sIn, sWh, acc[ps], cout >> 2, w0 + 8 * ps, w1, lane);
            }
            const float* yp = T.yext[L - 1] + sg * N * 64;
            float* gzp = T.gz[L - 1] + sg * N * 64;
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int n = w0 + 8 * ps + u;
                    if (n < w1) {
                        const float2 yv = __ldg(reinterpret_cast<const float2*>(yp + n * 64 + 2 * lane));
                        const float2 xv = *reinterpret_cast<const float2*>(sX + n * XS + 2 * lane);
                        const float dx0 = __fadd_rn(acc[ps][u][0].x, acc[ps][u][0].y), dx1 = __fadd_rn(acc[ps][u][1].x, acc[ps][u][1].y);
                        const float g0 = dx0 * slope_of(xv.x), g1 = dx1 * slope_of(xv.y);
                        *reinterpret_cast<float2*>(gzp + n * 64 + 2 * lane) = make_float2(g0, g1);
                        const float yh0 = __fmul_rn(__fsub_rn(yv.x, sPrm[2 * lane]), sPrm[64 + 2 * lane]);
                        const float yh1 = __fmul_rn(__fsub_rn(yv.y, sPrm[2 * lane + 1]), sPrm[65 + 2 * lane]);
                        accb[0] += static_cast<double>(g0);
                        accb[1] += static_cast<double>(g1);
                        accg[0] += static_cast<double>(g0) * static_cast<double>(yh0);
                        accg[1] += static_cast<double>(g1) * static_cast<double>(yh1);
                    }
                }
            }
        }
        // ---- conv-weight gradient tiles ----
        if (cin == 64) {
            if (cout == 64) accum_dw<4>(sX, sDA, sGZ, N, 4 * (tid >> 4), 4 * (tid & 15), ga, gb);
            else            accum_dw<2>(sX, sDA, sGZ, N, 2 * (tid >> 4), 4 * (tid & 15), ga, gb);
        } else if (cin == 12) {
            accum_dw<1>(sX, sDA, sGZ, N, tid >> 2, 4 * (tid & 3), ga, gb);
        } else if (tid < 64) {
            accum_dw<1>(sX, sDA, sGZ, N, tid, 0, ga, gb);
        }
        __syncthreads();
    }
    // ---- partial rows: [cout][Wa(cin) | Wb(cin)] ----
    float* row = (br ? part1 : part0) + static_cast<size_t>(blockIdx.x >> 1) * (cout * 2 * cin);
    {
        int cpt, c0, ci0;
        bool active = true;
        if (cin == 64) { cpt = cout == 64 ? 4 : 2; c0 = cpt * (tid >> 4); ci0 = 4 * (tid & 15); }
        else if (cin == 12) { cpt = 1; c0 = tid >> 2; ci0 = 4 * (tid & 3); }
        else { cpt = 1; c0 = tid; ci0 = 0; active = tid < 64; }
        if (active) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u < cpt) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (ci0 + v < cin) {
                            row[(c0 + u) * 2 * cin + ci0 + v] = ga[u * 4 + v];
                            row[(c0 + u) * 2 * cin + cin + ci0 + v] = gb[u * 4 + v];
                        }
                    }
                }
            }
        }
    }
    if (l > 0) {
        __syncthreads();
        sRed[warp * 128 + 2 * lane] = accb[0];
        sRed[warp * 128 + 2 * lane + 1] = accb[1];
        sRed[warp * 128 + 64 + 2 * lane] = accg[0];
        sRed[warp * 128 + 64 + 2 * lane + 1] = accg[1];
        __syncthreads();
        if (tid < 128) {
            double s = 0.0;
            for (int w = 0; w < kWarps; ++w) s += sRed[w * 128 + tid];
            atomicAdd(stat_ptr(T.bsum, side, L - 1) + tid, s);
        }
    }
}

// =====================================================================================================================
// Optimiser: gradient = fixed-order sum of the partials (BN gamma/beta: the fp64 sums of both sides); torch.optim.Adam
// with L2 weight decay (sg_net.py:351-352); BatchNorm running statistics, side 1 then side 2; mean loss.
// =====================================================================================================================
struct AdamArgs {
    float lr, wd, b1, b2, eps;
    float bc1;          // 1 - b1^t
    float bc2_sqrt;     // sqrt(1 - b2^t)
    int apply;          // kApplyNone / kApplyAll / kApplyRunning
};

// One block per 32 consecutive state elements: lane = element, the 8 warps split the partial rows (j = warp, warp + 8,
// ...) and their sums are added in warp order — a fixed summation order whatever the grid did.
#ifndef SGPR_TRAIN_INST_ONLY   // non-template kernels live in train.cu only (train_inst.cu holds the per-NPL templates)
__global__ void __launch_bounds__(kThreads) sgpr_train_adam_kernel(const TrainWs T, const AdamArgs A) {
    __shared__ float sPart[kWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    const bool is_bn = e >= P_BN && e < P_BN + kBnFloats;
    float g = 0.0f;
    if (A.apply == kApplyRunning && blockIdx.x * 32 + 31 < P_TOTAL) return;      // only the running-statistics blocks work
    if (e < P_TOTAL && !is_bn && A.apply != kApplyRunning) {
        for (int s = 0; s < T.nseg; ++s) {
            const Segment& sg = T.seg[s];
            if (e >= sg.off && e < sg.off + sg.size) {
                const float* p = sg.part + (e - sg.off);
#pragma unroll 4
                for (int j = warp; j < sg.count; j += kWarps) g += __ldg(p + static_cast<size_t>(j) * sg.stride);
                break;
            }
        }
    }
    sPart[warp][lane] = g;
    __syncthreads();
    if (warp != 0) return;
    if (blockIdx.x == 0 && lane == 0 && A.apply != kApplyRunning) {
        float s = 0.0f;
        for (int i = 0; i < T.head_grid; ++i) s += T.losspart[i];
        T.loss[0] = s / static_cast<float>(T.G);
    }
    if (e < P_TOTAL) {
        if (A.apply == kApplyRunning) return;
        if (is_bn) {
            int L = 6;
            while (bn_off(L) > e - P_BN) --L;
            const int C = layer_cout(L), r = e - P_BN - bn_off(L);
            const int which = r < C ? 1 : 0, c = r < C ? r : r - C;        // gamma <- dgamma (slot 1), beta <- dbeta (slot 0)
            double sum = 0.0;
            for (int side = 0; side < T.S; ++side) sum += stat_ptr(T.bsum, side, L)[which * 64 + c];
            g = static_cast<float>(sum);
        } else {
            g = 0.0f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) g += sPart[w][lane];
        }
        T.grads[e] = g;
        if (A.apply == kApplyAll) {
            const float p = T.state[e];
            g = fmaf(A.wd, p, g);
            const float m = A.b1 * T.adam_m[e] + (1.0f - A.b1) * g;
            const float v = A.b2 * T.adam_v[e] + (1.0f - A.b2) * g * g;
            T.adam_m[e] = m;
            T.adam_v[e] = v;
            const float denom = sqrtf(v) / A.bc2_sqrt + A.eps;
            const float pn = p - (A.lr / A.bc1) * (m / denom);
            T.state[e] = pn;
            if (e >= P_S2W && e < P_BN) {               // a GEMM weight: refresh its packed copies (xyz layer 1 is read natural)
                int L = 6;
                while (conv_off(L) > e) --L;
                pack_element(T.wpk, L, e - conv_off(L), pn);
            }
        }
    } else if (e < STATE_TOTAL && A.apply != kApplyNone) {
        int L = 6;
        const int r0 = e - R_OFF;
        while (bn_off(L) > r0) --L;
        const int C = layer_cout(L), r = r0 - bn_off(L);
        const int is_var = r >= C, c = is_var ? r - C : r;
        const double cnt = (L == 6) ? static_cast<double>(T.G) * T.N : static_cast<double>(T.G) * T.N * T.k;
        float run = T.state[e];
        for (int side = 0; side < 2; ++side) {          // two BatchNorm calls per step, side 1 first (sg_net.py:123-124)
            const double* st = stat_ptr(T.stats, side < T.S ? side : T.S - 1, L);
            const double m = st[c] / cnt;
            double var = st[64 + c] / cnt - m * m;
            var = var > 0.0 ? var : 0.0;
            const float val = is_var ? static_cast<float>(var * cnt / (cnt - 1.0)) : static_cast<float>(m);
            run = (1.0f - kMomentum) * run + kMomentum * val;
        }
        T.state[e] = run;
    }
}
#endif  // !SGPR_TRAIN_INST_ONLY

// =====================================================================================================================
// Training-batch assembly + augmentation on the device.  Replaces, for graphs already resident in HBM, the host loop of
// process_batch (sg_net.py:316-331) with transfer_to_torch's training branch (sg_net.py:286-295) and the point-cloud
// augmentations of utils.py:91-178 (rotate about z, jitter sigma 0.01 clip 0.05, scale U[0.8, 1.25), rotation
// perturbation sigma 0.015 clip 0.045, shift U[-0.3, 0.3) — the reference's defaults, applied in that order to all
// node_num rows, zero pads included; one x-flip decision per listed pair shared by its two graphs, sg_net.py:288-291).
// Random numbers: Philox4x32-10 keyed by the seed, counter = (step, slot, node, stream) — reproducible from
// (seed, step), but NOT numpy's global Mersenne stream (the reference never seeds it).
// Output row 2p / 2p+1 = augmented graph a / b of listed pair p: features_1 of the mirrored batch.
// =====================================================================================================================
struct Philox { uint32_t v[4]; };
__host__ __device__ inline Philox philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0, n1 = static_cast<uint32_t>(p1);
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1, n3 = static_cast<uint32_t>(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}
__host__ __device__ inline float u01(uint32_t x) { return static_cast<float>(x >> 8) * (1.0f / 16777216.0f); }         // [0, 1)
__host__ __device__ inline float u01_open(uint32_t x) { return (static_cast<float>(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }  // (0, 1]
// two standard normals from two words (Box-Muller)
__device__ __forceinline__ void normal2(uint32_t a, uint32_t b, float& n0, float& n1) {
    const float r = sqrtf(-2.0f * logf(u01_open(a)));
    const float t = 6.283185307179586f * u01(b);
    n0 = r * cosf(t);
    n1 = r * sinf(t);
}

struct AssembleArgs {
    const float* graphs;      // [M][15][N] un-augmented padded blocks
    const int* pair_idx;      // [P][2]
    float* out;               // [2P][15][N]
    float* draws;             // optional [2P][12]: flip u, angle u, scale, 3 raw perturbation normals, 3 shifts (debug / tests)
    float* jitter;            // optional [2P][N][3] raw jitter normals (debug / tests)
    int M, N, P;
    uint32_t seed_lo, seed_hi, step_lo, step_hi;
};

#ifndef SGPR_TRAIN_INST_ONLY   // non-template kernels live in train.cu only (train_inst.cu holds the per-NPL templates)
__global__ void __launch_bounds__(kThreads) sgpr_train_assemble_kernel(const AssembleArgs A) {
    __shared__ float sPar[16];
    const int tid = threadIdx.x, N = A.N;
    for (int slot = blockIdx.x; slot < 2 * A.P; slot += gridDim.x) {
        const int p = slot >> 1;
        const int g = A.pair_idx[slot];
        if (tid == 0) {
            const Philox f = philox4x32(A.step_lo, A.step_hi, static_cast<uint32_t>(p), 0xF11Fu, A.seed_lo, A.seed_hi);
            const Philox a = philox4x32(A.step_lo, A.step_hi, static_cast<uint32_t>(slot), 0xA001u, A.seed_lo, A.seed_hi);
            const Philox b = philox4x32(A.step_lo, A.step_hi, static_cast<uint32_t>(slot), 0xA002u, A.seed_lo, A.seed_hi);
            float n0, n1, n2, n3;
            normal2(a.v[2], a.v[3], n0, n1);
            normal2(b.v[0], b.v[1], n2, n3);
            sPar[0] = u01(f.v[0]);                                   // flip when > 0.5 (random.random() > 0.5)
            sPar[1] = u01(a.v[0]);                                   // rotation angle / 2 pi
            sPar[2] = 0.8f + (1.25f - 0.8f) * u01(a.v[1]);           // scale
            sPar[3] = n0; sPar[4] = n1; sPar[5] = n2;                // raw perturbation angles (x sigma, clipped below)
            sPar[6] = -0.3f + 0.6f * u01(b.v[2]);                    // shifts
            sPar[7] = -0.3f + 0.6f * u01(b.v[3]);
            sPar[8] = -0.3f + 0.6f * u01(philox4x32(A.step_lo, A.step_hi, static_cast<uint32_t>(slot), 0xA003u, A.seed_lo, A.seed_hi).v[0]);
            if (A.draws) for (int i = 0; i < 9; ++i) A.draws[slot * 12 + i] = sPar[i];
        }
        __syncthreads();
        const float* src = A.graphs + static_cast<size_t>(g) * kInCh * N;
        float* dst = A.out + static_cast<size_t>(slot) * kInCh * N;
        for (int e = tid; e < kLabels * N; e += kThreads) dst[3 * N + e] = __ldg(src + 3 * N + e);     // label rows unchanged
        const float flip = sPar[0] > 0.5f ? -1.0f : 1.0f;
        const float ang = sPar[1] * 6.283185307179586f, c = cosf(ang), s = sinf(ang);
        const float scale = sPar[2];
        const float ax = fminf(fmaxf(0.015f * sPar[3], -0.045f), 0.045f), ay = fminf(fmaxf(0.015f * sPar[4], -0.045f), 0.045f),
                    az = fminf(fmaxf(0.015f * sPar[5], -0.045f), 0.045f);
        const float cx = cosf(ax), sx = sinf(ax), cy = cosf(ay), sy = sinf(ay), cz = cosf(az), sz = sinf(az);
        // R = Rz (Ry Rx), applied to row vectors: out = p R   (utils.py rotate_perturbation_point_cloud)
        const float r00 = cz * cy, r01 = cz * sy * sx - sz * cx, r02 = cz * sy * cx + sz * sx;
        const float r10 = sz * cy, r11 = sz * sy * sx + cz * cx, r12 = sz * sy * cx - cz * sx;
        const float r20 = -sy, r21 = cy * sx, r22 = cy * cx;
        for (int n = tid; n < N; n += kThreads) {
            float x = flip * __ldg(src + n), y = __ldg(src + N + n), z = __ldg(src + 2 * N + n);
            const float xr = x * c + y * s, yr = -x * s + y * c;                      // p . Rot_z(angle)
            const Philox q = philox4x32(A.step_lo, A.step_hi, static_cast<uint32_t>(slot), 0xB000u + static_cast<uint32_t>(n),
                                        A.seed_lo, A.seed_hi);
            float j0, j1, j2, j3;
            normal2(q.v[0], q.v[1], j0, j1);
            normal2(q.v[2], q.v[3], j2, j3);
            if (A.jitter) { float* jd = A.jitter + (static_cast<size_t>(slot) * N + n) * 3; jd[0] = j0; jd[1] = j1; jd[2] = j2; }
            x = (xr + fminf(fmaxf(0.01f * j0, -0.05f), 0.05f)) * scale;
            y = (yr + fminf(fmaxf(0.01f * j1, -0.05f), 0.05f)) * scale;
            z = (z + fminf(fmaxf(0.01f * j2, -0.05f), 0.05f)) * scale;
            dst[n] = x * r00 + y * r10 + z * r20 + sPar[6];
            dst[N + n] = x * r01 + y * r11 + z * r21 + sPar[7];
            dst[2 * N + n] = x * r02 + y * r12 + z * r22 + sPar[8];
        }
        __syncthreads();
    }
}
#endif  // !SGPR_TRAIN_INST_ONLY

}  // namespace train
}  // namespace sgpr
