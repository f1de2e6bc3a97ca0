// Host-side packing of the reference state_dict into the kernel layout (see PackedWeights in common.cuh).
// Replaces nothing the reference computes per call: nn.BatchNorm in eval mode evaluates
//   y = x * alpha + beta,  alpha = gamma / sqrt(running_var + eps),  beta = bias - running_mean * alpha
// on every forward (ATen batch_norm eval path; modules at /root/reference/sg_net.py:50-76); here it is done once.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

#include "../../include/sgpr_b200.h"
#include "common.cuh"

namespace sgpr {

struct PackOffsets {          // offsets in floats into the blob; every section is 16-byte aligned
    size_t s1, w_s2, w_s3, w_f1, w_f2, w_f3, w_end;
    size_t ab_s2, ab_s3, ab_f1, ab_f2, ab_f3, ab_end;
    size_t att_w, ntn_w, ntn_v, ntn_b;
    size_t wtc_s2, wtc_s3, wtc_f2, wtc_f3;   // tensor-core operand planes of the 64-channel EdgeConv layers (embed_tc_kernel.cuh)
    size_t total;
};

inline size_t align4(size_t x) { return (x + 3) & ~static_cast<size_t>(3); }

inline PackOffsets make_offsets() {
    PackOffsets o{};
    size_t p = 0;
    auto take = [&](size_t n) { size_t at = p; p = align4(p + n); return at; };
    o.s1 = take(64 * 8);
    o.w_s2 = take(64 * 128);
    o.w_s3 = take(64 * 64);
    o.w_f1 = take(12 * 128);
    o.w_f2 = take(64 * 128);
    o.w_f3 = take(64 * 64);
    o.w_end = take(64 * 32);
    o.ab_s2 = take(128);
    o.ab_s3 = take(64);
    o.ab_f1 = take(128);
    o.ab_f2 = take(128);
    o.ab_f3 = take(64);
    o.ab_end = take(64);
    o.att_w = take(32 * 32);
    o.ntn_w = take(32 * 512);
    o.ntn_v = take(16 * 64);
    o.ntn_b = take(16);
    o.wtc_s2 = take(2 * 128 * 64);
    o.wtc_s3 = take(2 * 128 * 64);
    o.wtc_f2 = take(2 * 128 * 64);
    o.wtc_f3 = take(2 * 128 * 64);
    o.total = p;
    return o;
}

inline void bn_terms(const sgpr_bn& bn, int c, float eps, float& alpha, float& beta) {
    const float inv_std = 1.0f / std::sqrt(bn.running_var[c] + eps);
    alpha = inv_std * bn.weight[c];
    beta = bn.bias[c] - bn.running_mean[c] * alpha;
}

// Index of element (ci, co) of a [cin][CO] matrix in the channel-PAIR layout the kernels read (CO = 32*CPL):
// row p = ci/2 holds, per output, the pair (W[2p][co], W[2p+1][co]); a lane owns CPL consecutive outputs and reads
// them as conflict-free 16-byte (CPL>=2) or 8-byte (CPL=1) vectors, outputs 2,3 of a lane in a second 128-float plane.
inline size_t pair_index(int ci, int co, int CO) {
    const int cpl = CO / 32;
    const int p = ci >> 1, r = ci & 1;
    const int lane = co / cpl, j = co % cpl;
    return static_cast<size_t>(p) * 2 * CO + (j >> 1) * 128 + lane * (cpl >= 2 ? 4 : 2) + (j & 1) * 2 + r;
}

// EdgeConv layer: conv weight [cout][2*cin] -> pair layout of the [cin][2*cout] matrix (A half | B half) with the
// BN-scale sign folded in.
inline void pack_edgeconv(const float* w, const sgpr_bn& bn, int cin, int cout, float eps, float* wt, float* ab) {
    const int CO = 2 * cout;
    for (int c = 0; c < cout; ++c) {
        float alpha, beta;
        bn_terms(bn, c, eps, alpha, beta);
        const float sign = (alpha < 0.0f) ? -1.0f : 1.0f;
        ab[c] = sign * alpha;
        ab[cout + c] = beta;
        for (int ci = 0; ci < cin; ++ci) {
            wt[pair_index(ci, c, CO)] = sign * w[static_cast<size_t>(c) * 2 * cin + ci];
            wt[pair_index(ci, cout + c, CO)] = sign * w[static_cast<size_t>(c) * 2 * cin + cin + ci];
        }
    }
}

// The same sign-folded [cin = 64][2*cout] matrix as pack_edgeconv builds, as the TMEM-resident A operand of the
// tensor-core GEMM D[m][node] = sum_k W[m][k] x_node[k]: row m < cout is the (x_j - x_i) half, cout <= m < 2*cout the x_i
// half, rows beyond 2*cout are zero; split x = big + small with big = x rounded to TF32 (ties away, cvt.rna.tf32.f32).
// Layout [plane: big, small][m = 128][k = 64] row-major — a thread (TMEM lane = m) reads its 64 values contiguously.
inline void pack_edgeconv_tc(const float* w, const sgpr_bn& bn, int cout, float eps, float* planes) {
    const int cin = 64;
    for (int i = 0; i < 2 * 128 * 64; ++i) planes[i] = 0.0f;
    for (int c = 0; c < cout; ++c) {
        float alpha, beta;
        bn_terms(bn, c, eps, alpha, beta);
        const float sign = (alpha < 0.0f) ? -1.0f : 1.0f;
        for (int half = 0; half < 2; ++half) {
            const int m = half * cout + c;
            for (int ci = 0; ci < cin; ++ci) {
                const float x = sign * w[static_cast<size_t>(c) * 2 * cin + half * cin + ci];
                uint32_t bits;
                std::memcpy(&bits, &x, 4);
                bits = (bits + 0x1000u) & 0xFFFFE000u;
                float big;
                std::memcpy(&big, &bits, 4);
                planes[m * 64 + ci] = big;
                planes[128 * 64 + m * 64 + ci] = x - big;
            }
        }
    }
}

inline int pack_weights(const sgpr_weights& hw, std::vector<float>& blob, HeadParams& hp, const PackOffsets& o) {
    blob.assign(o.total, 0.0f);
    const float eps = hw.bn_eps;
    // xyz layer 1: per output channel {wa0,wa1,wa2, wb0,wb1,wb2, alpha, beta}
    for (int c = 0; c < 64; ++c) {
        float alpha, beta;
        bn_terms(hw.s_bn[0], c, eps, alpha, beta);
        const float sign = (alpha < 0.0f) ? -1.0f : 1.0f;
        float* dst = blob.data() + o.s1 + c * 8;
        for (int q = 0; q < 6; ++q) dst[q] = sign * hw.s_conv_w[0][c * 6 + q];
        dst[6] = sign * alpha;
        dst[7] = beta;
    }
    pack_edgeconv(hw.s_conv_w[1], hw.s_bn[1], 64, 64, eps, blob.data() + o.w_s2, blob.data() + o.ab_s2);
    pack_edgeconv(hw.s_conv_w[2], hw.s_bn[2], 64, 32, eps, blob.data() + o.w_s3, blob.data() + o.ab_s3);
    pack_edgeconv(hw.f_conv_w[0], hw.f_bn[0], 12, 64, eps, blob.data() + o.w_f1, blob.data() + o.ab_f1);
    pack_edgeconv(hw.f_conv_w[1], hw.f_bn[1], 64, 64, eps, blob.data() + o.w_f2, blob.data() + o.ab_f2);
    pack_edgeconv(hw.f_conv_w[2], hw.f_bn[2], 64, 32, eps, blob.data() + o.w_f3, blob.data() + o.ab_f3);
    pack_edgeconv_tc(hw.s_conv_w[1], hw.s_bn[1], 64, eps, blob.data() + o.wtc_s2);
    pack_edgeconv_tc(hw.s_conv_w[2], hw.s_bn[2], 32, eps, blob.data() + o.wtc_s3);
    pack_edgeconv_tc(hw.f_conv_w[1], hw.f_bn[1], 64, eps, blob.data() + o.wtc_f2);
    pack_edgeconv_tc(hw.f_conv_w[2], hw.f_bn[2], 32, eps, blob.data() + o.wtc_f3);
    // conv_end [32][64] -> pair layout of [64][32]; BN sign kept (no max follows)
    for (int c = 0; c < 32; ++c) {
        float alpha, beta;
        bn_terms(hw.end_bn, c, eps, alpha, beta);
        blob[o.ab_end + c] = alpha;
        blob[o.ab_end + 32 + c] = beta;
        for (int ci = 0; ci < 64; ++ci) blob[o.w_end + pair_index(ci, c, 32)] = hw.end_conv_w[c * 64 + ci];
    }
    std::memcpy(blob.data() + o.att_w, hw.att_w, sizeof(float) * 32 * 32);
    std::memcpy(blob.data() + o.ntn_w, hw.ntn_w, sizeof(float) * 32 * 512);
    std::memcpy(blob.data() + o.ntn_v, hw.ntn_v, sizeof(float) * 16 * 64);
    std::memcpy(blob.data() + o.ntn_b, hw.ntn_b, sizeof(float) * 16);
    std::memcpy(hp.fc1_w, hw.fc1_w, sizeof(float) * 256);
    std::memcpy(hp.fc1_b, hw.fc1_b, sizeof(float) * 16);
    std::memcpy(hp.fc2_w, hw.fc2_w, sizeof(float) * 16);
    hp.fc2_b = hw.fc2_b[0];
    return 0;
}

}  // namespace sgpr
