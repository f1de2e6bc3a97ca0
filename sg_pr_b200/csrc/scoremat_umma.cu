// Host side of the tcgen05 score-matrix kernel (csrc/scoremat_umma.cuh).
#include "scoremat_umma.cuh"
#include "launchers.hpp"

namespace sgpr {

cudaError_t score_matrix_umma_optin() {
    return cudaFuncSetAttribute(umma::sgpr_score_matrix_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, umma::kSmemUmma);
}

size_t score_matrix_umma_scratch_floats(int R, int M) {
    const size_t r16 = (static_cast<size_t>(R) + umma::kTileI - 1) / umma::kTileI * umma::kTileI;
    const size_t m128 = (static_cast<size_t>(M) + umma::kTileJ - 1) / umma::kTileJ * umma::kTileJ;
    return 2 * r16 * 512 + 2 * m128 * 32 + static_cast<size_t>(R) * kT + static_cast<size_t>(M) * kT;
}

void score_matrix_umma_launch(int sm_count, cudaStream_t st, const float* pooled_rows, const float* pooled_cols, float* scratch,
                              float* scores, long long ld, int R, int M, const PackedWeights& pw, const HeadParams& hp) {
    const int r16 = (R + umma::kTileI - 1) / umma::kTileI * umma::kTileI;
    const int m128 = (M + umma::kTileJ - 1) / umma::kTileJ * umma::kTileJ;
    float* proj_big = scratch;
    float* proj_small = proj_big + static_cast<size_t>(r16) * 512;
    float* cols_big = proj_small + static_cast<size_t>(r16) * 512;
    float* cols_small = cols_big + static_cast<size_t>(m128) * 32;
    float* rowblk = cols_small + static_cast<size_t>(m128) * 32;
    float* colblk = rowblk + static_cast<size_t>(R) * kT;
    const int cap = sm_count * 8;
    umma::sgpr_ntn_split_kernel<<<r16 < cap ? r16 : cap, kThreads, 0, st>>>(pooled_rows, R, r16, 0, proj_big, proj_small, rowblk, pw);
    umma::sgpr_ntn_split_kernel<<<m128 < cap ? m128 : cap, kThreads, 0, st>>>(pooled_cols, M, m128, 1, cols_big, cols_small, colblk, pw);
    umma::ScoreMatArgs a{};
    a.cols_big = cols_big; a.cols_small = cols_small; a.proj_big = proj_big; a.proj_small = proj_small;
    a.rowblk = rowblk; a.colblk = colblk;
    a.scores = scores; a.ld = ld; a.R = R; a.M = M;
    a.n_ib = r16 / umma::kTileI;
    a.n_tiles = (m128 / umma::kTileJ) * a.n_ib;
    const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;      // persistent: one CTA per SM, contiguous tile ranges
    umma::sgpr_score_matrix_umma_kernel<<<grid, umma::kThreadsUmma, umma::kSmemUmma, st>>>(a, hp);
}

}  // namespace sgpr
