// Host side of the tcgen05 score-matrix kernel (csrc/scoremat_umma.cuh).
#include "scoremat_umma.cuh"
#include "launchers.hpp"
#include <cstring>
#include <algorithm>
using std::min;

namespace sgpr {

cudaError_t score_matrix_umma_optin() {
    cudaError_t e = cudaFuncSetAttribute(umma::sgpr_score_matrix_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, umma::kSmemUmma);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(umma::sgpr_ntn_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, umma::kSplitSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(umma::sgpr_score_matrix_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, umma::kSmemUmma2);
    return e;
}

// FC1 weights as the B operand of the second UMMA (scoremat_umma.cuh, version 2): two [16][32] K-major planes in the
// 128-byte-swizzled shared-memory image — rows [W1_big | W1_big] and rows [W1_small | 0].  Host side of sgpr_set_weights.
void score_matrix_fc1_planes(const float* fc1_w /* [16][16] row = output */, float* planes /* [2][16][32] */) {
    for (int i = 0; i < 2 * 16 * 32; ++i) planes[i] = 0.0f;
    for (int u = 0; u < 16; ++u) {
        for (int t = 0; t < 16; ++t) {
            const float w = fc1_w[u * 16 + t];
            uint32_t bits;
            memcpy(&bits, &w, 4);
            bits = (bits + 0x1000u) & 0xFFFFE000u;               // cvt.rna.tf32.f32: nearest, ties away from zero
            float big;
            memcpy(&big, &bits, 4);
            const float small = w - big;
            auto at = [&](int f) { return u * 32 + ((((f >> 2) ^ (u & 7)) << 2) + (f & 3)); };
            planes[at(t)] = big;
            planes[at(16 + t)] = big;
            planes[512 + at(t)] = small;
        }
    }
}


size_t score_matrix_umma_scratch_floats(int R, int M) {
    const size_t r16 = (static_cast<size_t>(R) + umma::kTileI - 1) / umma::kTileI * umma::kTileI;
    const size_t m128 = (static_cast<size_t>(M) + umma::kTileJ - 1) / umma::kTileJ * umma::kTileJ;
    return 2 * r16 * 512 + 2 * m128 * 32 + static_cast<size_t>(R) * kT;
}

void score_matrix_umma_launch(int sm_count, cudaStream_t st, const float* pooled_rows, const float* pooled_cols, float* scratch,
                              float* scores, long long ld, int R, int M, const PackedWeights& pw, const HeadParams& hp,
                              const float* fc1_planes_dev, int version, float* const* outs, int n_out) {
    const int r16 = (R + umma::kTileI - 1) / umma::kTileI * umma::kTileI;
    const int m128 = (M + umma::kTileJ - 1) / umma::kTileJ * umma::kTileJ;
    float* proj_big = scratch;
    float* proj_small = proj_big + static_cast<size_t>(r16) * 512;
    float* cols_big = proj_small + static_cast<size_t>(r16) * 512;
    float* cols_small = cols_big + static_cast<size_t>(m128) * 32;
    float* rowblk = cols_small + static_cast<size_t>(m128) * 32;
    // one launch prepares both sides: CTAs [0, gr) split the row side (64 KB of shared memory: the transposed NTN tensor),
    // CTAs [gr, gr + gc) the column side
    // (8 graphs per CTA step; one CTA per SM on the row side amortises its 66 KB table load over more steps, the column
    // side — no table — gets half an SM count: the whole preparation is a single wave)
    const int gr = min((r16 + 7) / 8, sm_count), gc = min((m128 + 7) / 8, sm_count / 2);
    umma::sgpr_ntn_split_kernel<<<gr + gc, kThreads, umma::kSplitSmem, st>>>(pooled_rows, R, r16, proj_big, proj_small, rowblk,
                                                                          pooled_cols, M, m128, cols_big, cols_small, nullptr, gr, pw);
    if (version == 2) {
        umma::ScoreMat2Args a{};
        a.cols_big = cols_big; a.cols_small = cols_small; a.proj_big = proj_big; a.proj_small = proj_small;
        a.rowblk = rowblk; a.fc1_planes = fc1_planes_dev;
        a.scores = scores; a.ld = ld; a.R = R; a.M = M;
        a.n_ib = r16 / umma::kTileI2;                             // r16 is a multiple of 16, hence of 8
        a.n_tiles = (m128 / umma::kTileJ) * a.n_ib;
        const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
        umma::sgpr_score_matrix_umma2_kernel<<<grid, umma::kThreadsUmma2, umma::kSmemUmma2, st>>>(a, hp);
        return;
    }
    umma::ScoreMatArgs a{};
    a.cols_big = cols_big; a.cols_small = cols_small; a.proj_big = proj_big; a.proj_small = proj_small;
    a.rowblk = rowblk;
    a.scores = scores; a.ld = ld; a.R = R; a.M = M;
    a.n_out = n_out;
    for (int p = 0; p < n_out && p < 8; ++p) a.outs[p] = outs[p];
    a.n_ib = r16 / umma::kTileI;
    a.n_tiles = (m128 / umma::kTileJ) * a.n_ib;
    const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;      // persistent: one CTA per SM, contiguous tile ranges
    umma::sgpr_score_matrix_umma_kernel<<<grid, umma::kThreadsUmma, umma::kSmemUmma, st>>>(a, hp);
}

}  // namespace sgpr
