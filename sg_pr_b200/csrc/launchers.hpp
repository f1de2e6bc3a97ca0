// Per-NPL launch wrappers.  The row-tiled kernels are templates on NPL (nodes per lane: N <= 32 / 64 / 128); each NPL is
// instantiated in its own object file (embed_inst.cu / train_inst.cu compiled with -DSGPR_INST_NPL=1|2|4) so that the
// build runs in parallel and a kernel edit recompiles only what it touches.  api.cu / train.cu see declarations only.
#pragma once
#include "common.cuh"

namespace sgpr {

struct EmbedArgs;

// sets the dynamic shared-memory opt-in of every instantiation (both k-NN tie rules) of sgpr_embed_kernel<NPL, *>
template <int NPL> cudaError_t embed_optin(int optin_bytes);
// ties: SGPR_TIES_CUDA / SGPR_TIES_CPU
template <int NPL> void embed_launch(int ties, int grid, int smem, cudaStream_t st, const EmbedArgs& a, const PackedWeights& pw,
                                     const HeadParams& hp);

#define SGPR_DECL_EMBED(NPL)                                                                                            \
    template <> cudaError_t embed_optin<NPL>(int optin_bytes);                                                           \
    template <> void embed_launch<NPL>(int ties, int grid, int smem, cudaStream_t st, const EmbedArgs& a,                \
                                       const PackedWeights& pw, const HeadParams& hp);
SGPR_DECL_EMBED(1)
SGPR_DECL_EMBED(2)
SGPR_DECL_EMBED(4)
#undef SGPR_DECL_EMBED

// embed_inst.cu (NPL = 2 object): the tensor-core variant of the fused kernel (embed_tc_kernel.cuh)
cudaError_t embed_tc_optin(int optin_bytes);
int embed_tc_smem(int ks);
void embed_tc_launch(int grid, cudaStream_t st, const EmbedArgs& a, const PackedWeights& pw, const HeadParams& hp);

#ifdef SGPR_TIMELINE
// debug builds (embed_inst.cu, NPL = 2 object): clock stamps written by the N <= 64 fused kernel
int debug_read_timeline(long long* out);
int debug_read_ctas(int* smid1024, long long* t2048, int* g1024);
#endif

// scoremat_umma.cu: the tcgen05 score-matrix kernel (not part of the tests/emu build: inline tcgen05 PTX)
cudaError_t score_matrix_umma_optin();
size_t score_matrix_umma_scratch_floats(int R, int M);     // operand planes + V-block terms
void score_matrix_fc1_planes(const float* fc1_w, float* planes_1024);      // host: B operand of the FC1 UMMA
void score_matrix_umma_launch(int sm_count, cudaStream_t st, const float* pooled_rows, const float* pooled_cols, float* scratch,
                              float* scores, long long ld, int R, int M, const PackedWeights& pw, const HeadParams& hp,
                              const float* fc1_planes_dev, int version, float* const* outs, int n_out);

namespace train {

struct TrainWs;

template <int NPL> cudaError_t train_optin(int optin_bytes);
template <int NPL> void launch_edge_fwd(int ties, int grid, size_t smem, cudaStream_t st, const TrainWs& W, int l);
template <int NPL> void launch_end_fwd(int grid, size_t smem, cudaStream_t st, const TrainWs& W);
template <int NPL> void launch_end_bwd(int grid, size_t smem, cudaStream_t st, const TrainWs& W, float* part);
template <int NPL> void launch_edge_bwd(int grid, size_t smem, cudaStream_t st, const TrainWs& W, int l, float* part0, float* part1);

#define SGPR_DECL_TRAIN(NPL)                                                                                            \
    template <> cudaError_t train_optin<NPL>(int optin_bytes);                                                           \
    template <> void launch_edge_fwd<NPL>(int ties, int grid, size_t smem, cudaStream_t st, const TrainWs& W, int l);    \
    template <> void launch_end_fwd<NPL>(int grid, size_t smem, cudaStream_t st, const TrainWs& W);                      \
    template <> void launch_end_bwd<NPL>(int grid, size_t smem, cudaStream_t st, const TrainWs& W, float* part);         \
    template <> void launch_edge_bwd<NPL>(int grid, size_t smem, cudaStream_t st, const TrainWs& W, int l, float* part0, \
                                          float* part1);
SGPR_DECL_TRAIN(1)
SGPR_DECL_TRAIN(2)
SGPR_DECL_TRAIN(4)
#undef SGPR_DECL_TRAIN

}  // namespace train
}  // namespace sgpr
