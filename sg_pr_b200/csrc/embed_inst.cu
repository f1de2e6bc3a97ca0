// One object per NPL (-DSGPR_INST_NPL=1|2|4): the instantiations of the fused eval kernel and their launch wrappers.
#include "../../include/sgpr_b200.h"
#include "embed_kernel.cuh"
#if SGPR_INST_NPL == 2
#include "embed_tc_kernel.cuh"
#endif
#include "launchers.hpp"

#ifndef SGPR_INST_NPL
#error "compile with -DSGPR_INST_NPL=1, 2 or 4"
#endif

namespace sgpr {

template <>
cudaError_t embed_optin<SGPR_INST_NPL>(int optin_bytes) {
#ifdef SGPR_EMU
    (void)optin_bytes;
    return cudaSuccess;
#else
    // the dynamic limit excludes each kernel's static __shared__ bytes
    const void* fns[6] = {reinterpret_cast<const void*>(&sgpr_embed_kernel<SGPR_INST_NPL, 0, 0>),
                          reinterpret_cast<const void*>(&sgpr_embed_kernel<SGPR_INST_NPL, 1, 0>),
                          reinterpret_cast<const void*>(&sgpr_embed_kernel<SGPR_INST_NPL, 0, 1>),
                          reinterpret_cast<const void*>(&sgpr_embed_kernel<SGPR_INST_NPL, 1, 1>),
                          reinterpret_cast<const void*>(&sgpr_embed_kernel<SGPR_INST_NPL, 0, 1, 1>),
                          reinterpret_cast<const void*>(&sgpr_embed_kernel<SGPR_INST_NPL, 1, 1, 1>)};
    for (const void* fn : fns) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, fn);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin_bytes - static_cast<int>(fa.sharedSizeBytes));
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
#endif
}

template <>
void embed_launch<SGPR_INST_NPL>(int ties, int grid, int smem, cudaStream_t st, const EmbedArgs& a, const PackedWeights& pw,
                                 const HeadParams& hp) {
    // a split launch whose units all get an SM of their own (`sole`) runs the one-CTA-per-SM compilation
    const bool sole = a.split && 2 * a.G <= a.sm_count;
    const auto kern = (ties == SGPR_TIES_CPU)
        ? (sole ? &sgpr_embed_kernel<SGPR_INST_NPL, 1, 1, 1> : a.split ? &sgpr_embed_kernel<SGPR_INST_NPL, 1, 1> : &sgpr_embed_kernel<SGPR_INST_NPL, 1, 0>)
        : (sole ? &sgpr_embed_kernel<SGPR_INST_NPL, 0, 1, 1> : a.split ? &sgpr_embed_kernel<SGPR_INST_NPL, 0, 1> : &sgpr_embed_kernel<SGPR_INST_NPL, 0, 0>);
    SGPR_LAUNCH(kern, grid, kThreads, smem, st, a, pw, hp);
}

#if SGPR_INST_NPL == 2
// the tensor-core variant of the fused kernel (N <= 64, default tie rule, no debug taps)
cudaError_t embed_tc_optin(int optin_bytes) {
#ifdef SGPR_EMU
    (void)optin_bytes;
    return cudaSuccess;
#else
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, sgpr_embed_tc_kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(sgpr_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                optin_bytes - static_cast<int>(fa.sharedSizeBytes));
#endif
}

int embed_tc_smem(int ks) { return make_tc_layout(ks).total; }

void embed_tc_launch(int grid, cudaStream_t st, const EmbedArgs& a, const PackedWeights& pw, const HeadParams& hp) {
    SGPR_LAUNCH(sgpr_embed_tc_kernel, grid, kThreads, make_tc_layout(a.KS).total, st, a, pw, hp);
}

#ifdef SGPR_TIMELINE
// debug builds: the stamp arrays are per translation unit; the N <= 64 kernels of THIS object are the ones profiled
int debug_read_timeline(long long* out) {
    return cudaMemcpyFromSymbol(out, g_timeline, sizeof(long long) * kWarps * 128) == cudaSuccess ? 0 : -2;
}
int debug_read_ctas(int* smid1024, long long* t2048, int* g1024) {
    if (smid1024 && cudaMemcpyFromSymbol(smid1024, g_smid, sizeof(int) * 1024) != cudaSuccess) return -2;
    if (t2048 && cudaMemcpyFromSymbol(t2048, g_cta_t, sizeof(long long) * 2048) != cudaSuccess) return -2;
    if (g1024 && cudaMemcpyFromSymbol(g1024, g_cta_g, sizeof(int) * 1024) != cudaSuccess) return -2;
    return 0;
}
#endif
#endif

}  // namespace sgpr
