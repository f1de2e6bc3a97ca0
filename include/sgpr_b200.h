/*
 * sgpr_b200.h — C ABI of the B200-native SG_PR pairwise graph-similarity hot path.
 *
 * The reference (kxhit/SG_PR) has no native layer: its hot path is the Python method
 *     SG.forward(data) -> (score[B], att_1[B,N,1], att_2[B,N,1])          /root/reference/sg_net.py:112-138
 * executed as ~200 stock PyTorch launches.  This library is what sits UNDER that method in the drop-in
 * (`sg_pr_b200.sg_net.SG.forward`): plain pointers and sizes, no torch types, one fused sm_100a kernel per
 * batch.  Every entry point names the reference code it replaces.
 *
 * Conventions
 *   - All tensors are fp32, contiguous.  A graph is a channel-major block [15][N]: rows 0-2 node-centre xyz,
 *     rows 3-14 the 12 semantic-label channels — exactly what SGTrainer.transfer_to_torch emits
 *     (sg_net.py:296-299) and SG.forward consumes (sg_net.py:119-124).
 *   - `*_dev` pointers are device pointers on the context's device, `*_host` pointers are host memory
 *     (pinned or pageable).  Device entry points enqueue on `stream` (a cudaStream_t passed as void*, NULL =
 *     default stream) and never synchronise; the caller owns every buffer.
 *   - Return value: SGPR_OK (0) or a negative SGPR_E_* code; sgpr_last_error() gives the text for the calling
 *     thread.  The reference signals the same conditions with Python exceptions (e.g. topk raising when k > N,
 *     dgcnn.py:19; load_state_dict(strict) raising on a shape mismatch, sg_net.py:174).
 *   - A context is not re-entrant: it owns device workspace (arrival counters, the branch hand-over buffer, the
 *     ordering pre-pass) that consecutive launches share, so calls on ONE context must be issued from one thread at
 *     a time and onto one stream (or onto streams the caller orders against each other).  The reference has the same
 *     shape: a single Python thread, one CUDA stream.  Use one context per concurrent stream.
 *   - There is no CPU implementation behind this ABI: without a CUDA device sgpr_create fails.
 */
#ifndef SGPR_B200_H
#define SGPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGPR_ABI_VERSION 1

#define SGPR_OK            0
#define SGPR_E_INVALID   (-1)   /* bad argument / unsupported shape (N > 128, k > N, k < 1, ...)        */
#define SGPR_E_CUDA      (-2)   /* a CUDA runtime call failed                                          */
#define SGPR_E_NOWEIGHTS (-3)   /* forward called before sgpr_set_weights                              */
#define SGPR_E_ARCH      (-4)   /* layer sizes other than the reference architecture (config.yml:10-14) */

#define SGPR_NUM_LABELS    12   /* sg_net.py:200-202  global_labels = {0..11}                          */
#define SGPR_IN_CHANNELS   15   /* 3 + SGPR_NUM_LABELS                                                 */
#define SGPR_FILTERS_1     64   /* config/config.yml:11                                                */
#define SGPR_FILTERS_2     64   /* config/config.yml:12                                                */
#define SGPR_FILTERS_3     32   /* config/config.yml:13                                                */
#define SGPR_TENSOR_NEURONS 16  /* config/config.yml:14                                                */
#define SGPR_BOTTLENECK    16   /* config/config.yml:15                                                */
#define SGPR_MAX_NODES    128   /* largest node_num the fused kernel tiles in shared memory            */

/* One BatchNorm in eval mode (running statistics): nn.BatchNorm2d/1d at sg_net.py:52,56,60,64,68,72,76. */
typedef struct sgpr_bn {
    const float* weight;        /* gamma [C]        */
    const float* bias;          /* beta  [C]        */
    const float* running_mean;  /* [C]              */
    const float* running_var;   /* [C]              */
} sgpr_bn;

/*
 * Host pointers to the reference state_dict tensors (checkpoint keys after stripping "module.",
 * sg_net.py:164-174; shapes for the reference architecture).  Conv weights are the row-major
 * [C_out][2*C_in] matrices of the 1x1 convolutions (columns 0..C_in-1 multiply (x_j - x_i), columns
 * C_in..2C_in-1 multiply x_i — dgcnn.py:47).
 */
typedef struct sgpr_weights {
    const float* s_conv_w[3];   /* dgcnn_s_conv{1,2,3}.0.weight  [64][6], [64][128], [32][128]   sg_net.py:50-69 */
    sgpr_bn      s_bn[3];       /* dgcnn_s_conv{1,2,3}.1.*                                                       */
    const float* f_conv_w[3];   /* dgcnn_f_conv{1,2,3}.0.weight  [64][24], [64][128], [32][128]  sg_net.py:54-73 */
    sgpr_bn      f_bn[3];       /* dgcnn_f_conv{1,2,3}.1.*                                                       */
    const float* end_conv_w;    /* dgcnn_conv_end.0.weight       [32][64]                        sg_net.py:74-76 */
    sgpr_bn      end_bn;        /* dgcnn_conv_end.1.*                                                            */
    const float* att_w;         /* attention.weight_matrix       [32][32]               layers_batch.py:16-18    */
    const float* ntn_w;         /* tensor_network.weight_matrix  [32][32][16]           layers_batch.py:58       */
    const float* ntn_v;         /* tensor_network.weight_matrix_block [16][64]          layers_batch.py:59       */
    const float* ntn_b;         /* tensor_network.bias           [16]                   layers_batch.py:60       */
    const float* fc1_w;         /* fully_connected_first.weight  [16][16]               sg_net.py:47             */
    const float* fc1_b;         /* fully_connected_first.bias    [16]                                            */
    const float* fc2_w;         /* scoring_layer.weight          [1][16]                sg_net.py:48             */
    const float* fc2_b;         /* scoring_layer.bias            [1]                                             */
    float        bn_eps;        /* 1e-5 (nn.BatchNorm default)                                                   */
    int32_t      filters[3];    /* must be {64, 64, 32}; anything else -> SGPR_E_ARCH                            */
    int32_t      tensor_neurons;/* must be 16                                                                    */
    int32_t      bottleneck;    /* must be 16                                                                    */
} sgpr_weights;

typedef struct sgpr_ctx sgpr_ctx;   /* opaque: device id, packed weights, pair counters, staging buffers */

int         sgpr_abi_version(void);
const char* sgpr_last_error(void);

/* Replaces SG.__init__/setup_layers + `.cuda(gpu)` (sg_net.py:24-76, 175-176): binds a context to a device. */
int sgpr_create(sgpr_ctx** out, int device);
int sgpr_destroy(sgpr_ctx* ctx);

/*
 * Replaces load_state_dict + model.eval() (sg_net.py:164-174, eval_pair.py:11): packs the eval-mode weights
 * for the kernels (BN folded to scale/shift, sign-normalised so max-before-BN is exact, conv matrices split
 * and transposed) and uploads them.  Call again whenever the parameters change.
 */
int sgpr_set_weights(sgpr_ctx* ctx, const sgpr_weights* host_weights);

/*
 * THE HOT PATH.  Replaces SG.forward(data) in eval mode (sg_net.py:112-138): 2x dgcnn_conv_pass
 * (sg_net.py:79-110; dgcnn.py:14-49) -> AttentionModule x2 (layers_batch.py:28-39) -> TenorNetworkModule
 * (layers_batch.py:70-83) -> FC + sigmoid (sg_net.py:131-136), as ONE kernel launch.
 *   f1_dev, f2_dev : [B][15][N]           data["features_1"/"features_2"] after .cuda()   sg_net.py:119-120
 *   score_dev      : [B]                  return value 1
 *   att1_dev/att2_dev : [B][N] or NULL    return values 2,3 ([B,N,1] in the reference)
 */
int sgpr_forward_pairs(sgpr_ctx* ctx, const float* f1_dev, const float* f2_dev, int B, int N, int k,
                       float* score_dev, float* att1_dev, float* att2_dev, void* stream);

/*
 * Same call with HOST buffers — the boundary the reference actually exposes: SG.forward takes CPU tensors
 * (sg_net.py:517-521) and its callers read the result back with .cpu() (sg_net.py:523).  H2D copy, kernel,
 * D2H copy and one stream synchronise happen inside.  att*_host may be NULL.
 */
int sgpr_forward_pairs_host(sgpr_ctx* ctx, const float* f1_host, const float* f2_host, int B, int N, int k,
                            float* score_host, float* att1_host, float* att2_host);

/*
 * Compact input: the [15][N] block the reference builds per graph (sg_net.py:270-299) is 3 coordinates
 * plus a ONE-HOT label per node — 60 bytes a node for 13 bytes of information.  A compact graph record is
 *     float   xyz[3][N];     rows 0-2 of the block, unchanged
 *     uint8_t label[N];      0..11 = the one-hot row that holds the 1; any other value = no label (zero pad)
 * padded to sgpr_compact_stride(N) = ceil16(13 N) bytes; records are contiguous and 16-byte aligned.  The kernel pulls
 * one record per graph (device memory, or pinned host memory read in place over PCIe) and expands it in shared memory
 * into the block the one-hot entry points read — results are bit-identical to sgpr_forward_pairs / sgpr_embed on the
 * expanded input.  4.6x fewer input bytes per pair (1,664 instead of 7,680 at N = 64).
 */
size_t sgpr_compact_stride(int N);
/* Host-only: G one-hot blocks [15][N] fp32 -> G compact records (records_host: G * sgpr_compact_stride(N) bytes, e.g. a
 * pinned staging buffer).  What a caller holding the reference's own input tensors (sg_net.py:296-299, 517-519) runs to
 * hand them over in 13 instead of 60 bytes a node.  Returns 0, or 1 when some label value is neither +0.0f nor 1.0f or a
 * node carries more than one 1 — such a batch has no compact form (the records are then unspecified); negative on bad
 * arguments. */
int sgpr_compact_from_blocks(const float* blocks_host, int G, int N, void* records_host);
int sgpr_forward_pairs_compact(sgpr_ctx* ctx, const void* graphs1, const void* graphs2, int B, int N, int k,
                               float* score_dev, float* att1_dev, float* att2_dev, void* stream);
int sgpr_embed_compact(sgpr_ctx* ctx, const void* graphs, int M, int N, int k, float* pooled_dev, float* att_dev,
                       float* emb_dev, void* stream);

/*
 * Per-graph half of the path (sg_net.py:123,126 for one side): node embeddings + attention pooling.
 *   graphs_dev : [M][15][N]
 *   pooled_dev : [M][32]        AttentionModule `representation`            layers_batch.py:38
 *   att_dev    : [M][N] | NULL  AttentionModule `sigmoid_scores`            layers_batch.py:37
 *   emb_dev    : [M][N][32] | NULL  dgcnn_conv_pass output                  sg_net.py:109-110
 * Every graph's result is independent of its batch, which is what makes the all-pairs scan
 * an embed-once + score-matrix job.
 */
int sgpr_embed(sgpr_ctx* ctx, const float* graphs_dev, int M, int N, int k,
               float* pooled_dev, float* att_dev, float* emb_dev, void* stream);

/*
 * Pair head only (layers_batch.py:70-83 + sg_net.py:131-136) on pooled vectors:
 *   score[p] = head(pooled[pair_idx[2p]], pooled[pair_idx[2p+1]])   — ordered, the score is not symmetric.
 */
int sgpr_score_pairs(sgpr_ctx* ctx, const float* pooled_dev, const int32_t* pair_idx_dev, int P,
                     float* score_dev, void* stream);

/*
 * All ordered pairs of a row block against all columns — the N x N sequence scan of BASELINE config 4
 * (the reference only has the offline loop gen_sem_kitti_graph_pairs.py:43-52):
 *   scores[r][c] = head(pooled_rows[r], pooled_cols[c]),  scores row stride = ld_scores floats (>= M).
 */
int sgpr_score_matrix(sgpr_ctx* ctx, const float* pooled_rows_dev, int R, const float* pooled_cols_dev, int M,
                      float* scores_dev, int64_t ld_scores, void* stream);

/*
 * The same with the exchange step of the multi-GPU scan fused into the kernel: every score is stored into ALL n_out
 * (1..8) destination matrices — this GPU's result and its peers' results mapped into this process (CUDA IPC) and
 * reachable over NVLink after sgpr_enable_peer_access — instead of one local store followed by an all-gather.
 * scores_dev_list[p] points at row 0 OF THIS ROW BLOCK inside destination p; same ld for all.  The stores are posted
 * writes that overlap the tensor-core tiles; the caller orders them against the peers' reads with a stream-ordered
 * collective afterwards (sg_pr_b200/scan.py uses a 1-element all-reduce).  tcgen05 kernel only.
 */
int sgpr_score_matrix_multi(sgpr_ctx* ctx, const float* pooled_rows_dev, int R, const float* pooled_cols_dev, int M,
                            float* const* scores_dev_list, int n_out, int64_t ld_scores, void* stream);
int sgpr_enable_peer_access(sgpr_ctx* ctx, int peer_device);
/*
 * Peer-visible buffers for it, one process per GPU: sgpr_peer_alloc = cudaMalloc on the context's device + a 64-byte
 * CUDA IPC handle to hand to the other processes; sgpr_peer_open maps a peer's handle into this process WITH THIS
 * CONTEXT'S DEVICE CURRENT (cudaIpcMemLazyEnablePeerAccess), which is what makes the returned pointer usable by this
 * GPU's kernels over NVLink.  sgpr_peer_close / sgpr_peer_free undo them.
 */
int sgpr_peer_alloc(sgpr_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char* handle64);
int sgpr_peer_open(sgpr_ctx* ctx, const unsigned char* handle64, void** dev_ptr);
int sgpr_peer_close(sgpr_ctx* ctx, void* dev_ptr);
int sgpr_peer_free(sgpr_ctx* ctx, void* dev_ptr);

/* Debug/parity taps: k-NN index lists and per-layer EdgeConv outputs of sgpr_embed for the parity tests.
 *   knn_dev       : [M][6][N][k] uint8 | NULL   (layer order: xyz 1-3, sem 1-3 — sg_net.py:84-102)
 *   layer_out_dev : [M][6][N][64] float | NULL  (32-channel layers use the first 32 columns)            */
int sgpr_embed_trace(sgpr_ctx* ctx, const float* graphs_dev, int M, int N, int k,
                     float* pooled_dev, float* att_dev, float* emb_dev,
                     uint8_t* knn_dev, float* layer_out_dev, void* stream);

/*
 * Host-only view of the weight packing (no device needed; used by the CPU tests of the packing rules).
 * sgpr_packed_size() = floats in the packed blob; sgpr_pack_weights_host fills `blob` (that many floats) and the
 * 289-float FC head {fc1_w[256], fc1_b[16], fc2_w[16], fc2_b}; `offsets` receives the 17 section offsets in the
 * order of struct PackedWeights (csrc/common.cuh).
 */
size_t sgpr_packed_size(void);
int sgpr_pack_weights_host(const sgpr_weights* host_weights, float* blob, float* head289, size_t* offsets17);

/*
 * k-NN tie rule of `dgcnn.knn`'s topk (dgcnn.py:19).  Distances on the one-hot semantic branch (sg_net.py:94) tie at the
 * k-th position in most rows; when a graph has at least k zero pads every rule picks equivalent nodes, otherwise the
 * rule decides the score (by up to ~0.2).  ATen breaks such ties differently on its two devices, so "the reference"
 * means one of:
 *   SGPR_TIES_CUDA (default)  ties at the k-th value go to the lowest column index — ATen's CUDA radix-select topk,
 *                             i.e. the reference on its native device (`.cuda(gpu)`, sg_net.py:119-120, 176);
 *   SGPR_TIES_CPU             the order std::nth_element leaves behind in ATen's CPU topk (TopKImpl.h), restated step
 *                             for step in csrc/topk_nth.cuh — bit-for-bit the index sets of the reference run on a CPU
 *                             (the build container's oracle and golden vectors).  One lane per row, serial: a parity
 *                             mode, several times slower in the selection stage.
 * Applies to every later launch of the context.  SGPR_KNN_TIES=cpu in the environment sets the initial mode.
 */
#define SGPR_TIES_CUDA 0
#define SGPR_TIES_CPU  1
int sgpr_set_knn_ties(sgpr_ctx* ctx, int mode);
int sgpr_get_knn_ties(const sgpr_ctx* ctx);

/*
 * Host-only view of the SGPR_TIES_CPU selection (no device needed; used by the CPU tests that pin it against libstdc++
 * and torch's CPU topk): rows [num_rows][n] -> idx_out [num_rows][k] in nth_element's order.  depth_limit < 0 runs
 * std::nth_element's own budget 2*floor(log2 n); >= 0 forces it (exercises the heap-select fallback).
 */
int sgpr_topk_cpu_rule_host(const float* rows, int num_rows, int n, int k, int depth_limit, int32_t* idx_out);

/* Number of kernel launches this context has enqueued so far (bench.py's `gpu_launches`). */
int64_t sgpr_launch_count(const sgpr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SGPR_B200_H */
