// One object per NPL (-DSGPR_INST_NPL=1|2|4): the row-tiled training kernels and their launch wrappers.
#include "../../include/sgpr_b200_train.h"
#define SGPR_TRAIN_INST_ONLY
#include "train_kernels.cuh"
#include "launchers.hpp"

#ifndef SGPR_INST_NPL
#error "compile with -DSGPR_INST_NPL=1, 2 or 4"
#endif

namespace sgpr {
namespace train {

template <>
cudaError_t train_optin<SGPR_INST_NPL>(int optin_bytes) {
#ifdef SGPR_EMU
    (void)optin_bytes;
    return cudaSuccess;
#else
    constexpr int N = SGPR_INST_NPL;
    cudaError_t e = cudaFuncSetAttribute(sgpr_train_edge_fwd<N, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgpr_train_edge_fwd<N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgpr_train_edge_bwd<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgpr_train_end_fwd<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgpr_train_end_bwd<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin_bytes);
    return e;
#endif
}

template <>
void launch_edge_fwd<SGPR_INST_NPL>(int ties, int grid, size_t smem, cudaStream_t st, const TrainWs& W, int l) {
    if (ties == SGPR_TIES_CPU) {
        const auto kern = &sgpr_train_edge_fwd<SGPR_INST_NPL, 1>;
        SGPR_LAUNCH(kern, grid, kThreads, smem, st, W, l);
    } else {
        const auto kern = &sgpr_train_edge_fwd<SGPR_INST_NPL, 0>;
        SGPR_LAUNCH(kern, grid, kThreads, smem, st, W, l);
    }
}

template <>
void launch_end_fwd<SGPR_INST_NPL>(int grid, size_t smem, cudaStream_t st, const TrainWs& W) {
    SGPR_LAUNCH(sgpr_train_end_fwd<SGPR_INST_NPL>, grid, kThreads, smem, st, W);
}

template <>
void launch_end_bwd<SGPR_INST_NPL>(int grid, size_t smem, cudaStream_t st, const TrainWs& W, float* part) {
    SGPR_LAUNCH(sgpr_train_end_bwd<SGPR_INST_NPL>, grid, kThreads, smem, st, W, part);
}

template <>
void launch_edge_bwd<SGPR_INST_NPL>(int grid, size_t smem, cudaStream_t st, const TrainWs& W, int l, float* part0, float* part1) {
    SGPR_LAUNCH(sgpr_train_edge_bwd<SGPR_INST_NPL>, grid, kThreads, smem, st, W, l, part0, part1);
}

}  // namespace train
}  // namespace sgpr
