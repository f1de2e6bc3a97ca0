// C-ABI implementation of include/sgpr_b200_train.h: device state, workspace, the launch chain of one training step.
// Links only cudart.  (tests/emu builds this same file with the host compiler against tests/emu/cuda_emu.h — test
// infrastructure for debugging without a GPU; the product library is the nvcc build.)
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/sgpr_b200_train.h"
#include "train_kernels.cuh"
#include "launchers.hpp"

using namespace sgpr;
using namespace sgpr::train;

int sgpr_fail(int code, const char* fmt, ...);      // api.cu: sets the text sgpr_last_error() returns

namespace {

#define TRY_CUDA(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return sgpr_fail(SGPR_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

struct Guard {
    int prev = -1;
    explicit Guard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~Guard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

const char* const kNames[] = {
    "dgcnn_s_conv1.0.weight", "dgcnn_s_conv2.0.weight", "dgcnn_s_conv3.0.weight",
    "dgcnn_f_conv1.0.weight", "dgcnn_f_conv2.0.weight", "dgcnn_f_conv3.0.weight", "dgcnn_conv_end.0.weight",
    "dgcnn_s_conv1.1.weight", "dgcnn_s_conv1.1.bias", "dgcnn_s_conv2.1.weight", "dgcnn_s_conv2.1.bias",
    "dgcnn_s_conv3.1.weight", "dgcnn_s_conv3.1.bias", "dgcnn_f_conv1.1.weight", "dgcnn_f_conv1.1.bias",
    "dgcnn_f_conv2.1.weight", "dgcnn_f_conv2.1.bias", "dgcnn_f_conv3.1.weight", "dgcnn_f_conv3.1.bias",
    "dgcnn_conv_end.1.weight", "dgcnn_conv_end.1.bias",
    "attention.weight_matrix", "tensor_network.weight_matrix", "tensor_network.weight_matrix_block",
    "tensor_network.bias", "fully_connected_first.weight", "fully_connected_first.bias", "scoring_layer.weight",
    "scoring_layer.bias",
    "dgcnn_s_conv1.1.running_mean", "dgcnn_s_conv1.1.running_var", "dgcnn_s_conv2.1.running_mean",
    "dgcnn_s_conv2.1.running_var", "dgcnn_s_conv3.1.running_mean", "dgcnn_s_conv3.1.running_var",
    "dgcnn_f_conv1.1.running_mean", "dgcnn_f_conv1.1.running_var", "dgcnn_f_conv2.1.running_mean",
    "dgcnn_f_conv2.1.running_var", "dgcnn_f_conv3.1.running_mean", "dgcnn_f_conv3.1.running_var",
    "dgcnn_conv_end.1.running_mean", "dgcnn_conv_end.1.running_var"};
constexpr int kNumTensors = sizeof(kNames) / sizeof(kNames[0]);
constexpr int kNumParamTensors = 29;
int64_t g_offsets[kNumTensors], g_sizes[kNumTensors];
bool g_layout_ready = false;

void build_layout() {
    if (g_layout_ready) return;
    int i = 0;
    for (int L = 0; L < 7; ++L, ++i) { g_offsets[i] = conv_off(L); g_sizes[i] = conv_size(L); }
    for (int L = 0; L < 7; ++L) {
        const int C = layer_cout(L);
        g_offsets[i] = P_BN + bn_off(L); g_sizes[i++] = C;
        g_offsets[i] = P_BN + bn_off(L) + C; g_sizes[i++] = C;
    }
    const int head_off[8] = {P_ATT, P_NTNW, P_NTNV, P_NTNB, P_FC1W, P_FC1B, P_FC2W, P_FC2B};
    const int head_size[8] = {1024, 16384, 1024, 16, 256, 16, 16, 1};
    for (int h = 0; h < 8; ++h) { g_offsets[i] = head_off[h]; g_sizes[i++] = head_size[h]; }
    for (int L = 0; L < 7; ++L) {
        const int C = layer_cout(L);
        g_offsets[i] = R_OFF + bn_off(L); g_sizes[i++] = C;
        g_offsets[i] = R_OFF + bn_off(L) + C; g_sizes[i++] = C;
    }
    g_layout_ready = true;
}

}  // namespace

struct Plan {
    bool valid = false, backward_ready = false;
    TrainWs W{};
    int npl = 2, grid4 = 0, grid2 = 0, head_grid = 0;
    size_t fwd_smem = 0, bwd_smem = 0, end_fwd_smem = 0, end_bwd_smem = 0, pool_smem = 0;
    float* part_head = nullptr;
    float* part_att = nullptr;
    float* part_end = nullptr;
    float* part_conv[6] = {};
};

struct sgpr_train {
    int device = 0;
    int sm_count = 0;
    bool has_state = false;
    bool wpk_valid = false;        // the packed GEMM copies match d_state (kept current by the optimiser kernel)
    float* d_state = nullptr;      // [STATE_TOTAL]
    float* d_adam = nullptr;       // m | v  [2][P_TOTAL]
    float* d_grads = nullptr;      // [P_TOTAL]
    float* d_wpk = nullptr;        // [WPK_ALL]
    double* d_sums = nullptr;      // stats | bsum  [2][2][7][2][64]
    float* d_misc = nullptr;       // loss[1] | losspart[kMaxHeadGrid]
    unsigned char* d_ws = nullptr; size_t ws_cap = 0;       // per-batch workspace
    float* d_part = nullptr;       size_t part_cap = 0;     // gradient partials
    float lr = 1e-3f, wd = 0.0f, b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    long long steps = 0;
    long long launches = 0;
    long long forward_gen = 0;     // counts train-mode forwards; sgpr_train_backward names the one it differentiates
    int knn_ties = SGPR_TIES_CUDA; // k-NN tie rule (sgpr_train_set_knn_ties)
    TrainWs last{};                // pointers of the last step (debug taps)
    bool has_last = false;
    Plan plan;
};

namespace {

constexpr int kMaxHeadGrid = 1024;
constexpr size_t kSumDoubles = 2 * 2 * 7 * 128;

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

template <typename T>
T* carve(unsigned char*& p, size_t count) {
    T* out = reinterpret_cast<T*>(p);
    p += align256(count * sizeof(T));
    return out;
}

}  // namespace

extern "C" {

int sgpr_train_layout(int* count, int* n_params, const char* const** names, const int64_t** offsets, const int64_t** sizes) {
    build_layout();
    if (count) *count = kNumTensors;
    if (n_params) *n_params = kNumParamTensors;
    if (names) *names = kNames;
    if (offsets) *offsets = g_offsets;
    if (sizes) *sizes = g_sizes;
    return SGPR_OK;
}
int64_t sgpr_train_param_count(void) { return P_TOTAL; }
int64_t sgpr_train_state_count(void) { return STATE_TOTAL; }

int sgpr_train_create(sgpr_train** out, int device) {
    if (!out) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return sgpr_fail(SGPR_E_CUDA, "sgpr_train_create: no CUDA device; this library has no CPU path");
    if (device < 0 || device >= count) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_create: device %d out of range", device);
    Guard guard(device);
    cudaDeviceProp prop;
    TRY_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return sgpr_fail(SGPR_E_CUDA, "sgpr_train_create: device is sm_%d%d; built for sm_100a only", prop.major, prop.minor);
    sgpr_train* t = new (std::nothrow) sgpr_train();
    if (!t) return sgpr_fail(SGPR_E_CUDA, "sgpr_train_create: out of host memory");
    t->device = device;
    t->sm_count = prop.multiProcessorCount;
    const int optin = static_cast<int>(prop.sharedMemPerBlockOptin) - 1024;
    if (const char* kt = getenv("SGPR_KNN_TIES")) t->knn_ties = (strcmp(kt, "cpu") == 0) ? SGPR_TIES_CPU : SGPR_TIES_CUDA;
    e = train_optin<1>(optin);
    if (e == cudaSuccess) e = train_optin<2>(optin);
    if (e == cudaSuccess) e = train_optin<4>(optin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgpr_train_pool_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_state), STATE_TOTAL * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_adam), 2 * P_TOTAL * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_grads), P_TOTAL * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_wpk), WPK_ALL * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_sums), kSumDoubles * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->d_misc), (1 + kMaxHeadGrid) * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(t->d_adam, 0, 2 * P_TOTAL * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(t->d_grads, 0, P_TOTAL * sizeof(float));
    if (e != cudaSuccess) {
        int rc = sgpr_fail(SGPR_E_CUDA, "sgpr_train_create: %s", cudaGetErrorString(e));
        sgpr_train_destroy(t);
        return rc;
    }
    *out = t;
    return SGPR_OK;
}

int sgpr_train_destroy(sgpr_train* t) {
    if (!t) return SGPR_OK;
    Guard guard(t->device);
    cudaFree(t->d_state);
    cudaFree(t->d_adam);
    cudaFree(t->d_grads);
    cudaFree(t->d_wpk);
    cudaFree(t->d_sums);
    cudaFree(t->d_misc);
    cudaFree(t->d_ws);
    cudaFree(t->d_part);
    delete t;
    return SGPR_OK;
}

int sgpr_train_set_state(sgpr_train* t, const float* state_host, int reset_optimizer) {
    if (!t || !state_host) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_set_state: NULL argument");
    Guard guard(t->device);
    TRY_CUDA(cudaDeviceSynchronize());
    TRY_CUDA(cudaMemcpy(t->d_state, state_host, STATE_TOTAL * sizeof(float), cudaMemcpyHostToDevice));
    if (reset_optimizer) {
        TRY_CUDA(cudaMemset(t->d_adam, 0, 2 * P_TOTAL * sizeof(float)));
        t->steps = 0;
    }
    t->has_state = true;
    t->wpk_valid = false;
    return SGPR_OK;
}

int sgpr_train_get_state(sgpr_train* t, float* state_host) {
    if (!t || !state_host) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_get_state: NULL argument");
    if (!t->has_state) return sgpr_fail(SGPR_E_NOWEIGHTS, "sgpr_train_get_state: call sgpr_train_set_state first");
    Guard guard(t->device);
    TRY_CUDA(cudaDeviceSynchronize());
    TRY_CUDA(cudaMemcpy(state_host, t->d_state, STATE_TOTAL * sizeof(float), cudaMemcpyDeviceToHost));
    return SGPR_OK;
}

int sgpr_train_set_optimizer(sgpr_train* t, float lr, float weight_decay, float beta1, float beta2, float eps) {
    if (!t) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_set_optimizer: NULL context");
    if (!(lr >= 0.0f) || !(weight_decay >= 0.0f) || !(beta1 >= 0.0f && beta1 < 1.0f) || !(beta2 >= 0.0f && beta2 < 1.0f) || !(eps >= 0.0f))
        return sgpr_fail(SGPR_E_INVALID, "sgpr_train_set_optimizer: invalid hyper-parameter (torch.optim.Adam raises ValueError)");
    t->lr = lr; t->wd = weight_decay; t->b1 = beta1; t->b2 = beta2; t->eps = eps;
    return SGPR_OK;
}

}  // extern "C"

namespace {

// Everything one batch needs: workspace carve-out, grids, partial buffers.  Kept in the context after a forward so that a
// separate backward call (sgpr_train_backward) finds the same tensors.
int make_plan(sgpr_train* t, const float* f1_dev, const float* f2_dev, const float* target_dev, int B, int N, int k,
              int mirrored, const char* who) {
    if (!t->has_state) return sgpr_fail(SGPR_E_NOWEIGHTS, "%s: call sgpr_train_set_state first", who);
    if (B < 1) return sgpr_fail(SGPR_E_INVALID, "%s: batch %d < 1", who, B);
    if (N < 2 || N > SGPR_MAX_NODES) return sgpr_fail(SGPR_E_INVALID, "%s: node_num %d outside [2,%d]", who, N, SGPR_MAX_NODES);
    if (k < 1 || k > N)
        return sgpr_fail(SGPR_E_INVALID, "%s: k=%d must satisfy 1 <= k <= node_num=%d (topk raises in the reference, dgcnn.py:19)", who, k, N);
    if (!f1_dev || (!f2_dev && !mirrored)) return sgpr_fail(SGPR_E_INVALID, "%s: NULL feature pointer", who);
    if (mirrored && (B & 1)) return sgpr_fail(SGPR_E_INVALID, "%s: a mirrored batch holds every pair in both orders, B=%d is odd", who, B);
    const int G = B, S = mirrored ? 1 : 2, SG = S * B;
    Plan& P = t->plan;
    P.valid = false;
    P.npl = (N <= 32) ? 1 : (N <= 64) ? 2 : 4;
    const int nmax = 32 * P.npl;
    const int KS = (k + 3) & ~3;

    size_t need = 0;
    for (int L = 0; L < 6; ++L) {
        const size_t per = static_cast<size_t>(SG) * N * layer_cout(L);
        need += 5 * align256(per * 4) + align256(per) + align256(static_cast<size_t>(SG) * N * k);
    }
    need += 2 * align256(static_cast<size_t>(SG) * N * 32 * 4) + align256(static_cast<size_t>(SG) * 32 * 4) +
            align256(static_cast<size_t>(2) * G * 32 * 4) +
            align256(static_cast<size_t>(SG) * N * 4) + align256(static_cast<size_t>(G) * 4);
    if (need > t->ws_cap) {
        if (t->d_ws) cudaFree(t->d_ws);
        t->d_ws = nullptr; t->ws_cap = 0;
        const size_t want = need + need / 8;
        TRY_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->d_ws), want));
        t->ws_cap = want;
    }
    const int per_sm = (P.npl <= 2) ? 2 : 1;
    const int cap = t->sm_count * per_sm;
    P.grid4 = 2 * S * G < cap ? 2 * S * G : (cap / (2 * S)) * (2 * S);        // (branch, side) x graphs
    if (P.grid4 < 2 * S) P.grid4 = 2 * S;
    P.grid2 = S * G < cap ? S * G : (cap / S) * S;                            // side x graphs
    if (P.grid2 < S) P.grid2 = S;
    const int groups = mirrored ? G / 2 : G;                  // pool_head: one group = the two graphs of a pair
    P.head_grid = groups < 2 * t->sm_count ? groups : 2 * t->sm_count;
    if (P.head_grid > kMaxHeadGrid) P.head_grid = kMaxHeadGrid;
    P.pool_smem = static_cast<size_t>(4) * N * 33 * sizeof(float);
    const int nb = P.grid4 / 2;                                // partial rows per branch in the EdgeConv backward
    size_t part_need = static_cast<size_t>(P.head_grid) * kHeadFloats + static_cast<size_t>(P.head_grid) * 1024 +
                       static_cast<size_t>(P.grid2) * 2048;
    for (int L = 0; L < 6; ++L) part_need += static_cast<size_t>(nb) * conv_size(L);
    if (part_need > t->part_cap) {
        if (t->d_part) cudaFree(t->d_part);
        t->d_part = nullptr; t->part_cap = 0;
        TRY_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->d_part), part_need * sizeof(float)));
        t->part_cap = part_need;
    }

    TrainWs W{};
    W.G = G; W.S = S; W.mirrored = mirrored; W.N = N; W.k = k; W.KS = KS; W.eps = 1e-5f;
    W.f[0] = f1_dev; W.f[1] = f2_dev; W.target = target_dev; W.dpred = nullptr;
    W.state = t->d_state; W.wpk = t->d_wpk;
    unsigned char* p = t->d_ws;
    for (int L = 0; L < 6; ++L) {
        const size_t per = static_cast<size_t>(SG) * N * layer_cout(L);
        W.yext[L] = carve<float>(p, per);
        W.a[L] = carve<float>(p, per);
        W.d[L] = carve<float>(p, per);
        W.sumy[L] = carve<float>(p, per);
        W.gz[L] = carve<float>(p, per);
        W.enode[L] = carve<uint8_t>(p, per);
        W.idx[L] = carve<uint8_t>(p, static_cast<size_t>(SG) * N * k);
    }
    W.yend = carve<float>(p, static_cast<size_t>(SG) * N * 32);
    W.gzend = carve<float>(p, static_cast<size_t>(SG) * N * 32);
    W.pooled = carve<float>(p, static_cast<size_t>(SG) * 32);
    W.dpooled = carve<float>(p, static_cast<size_t>(2) * G * 32);        // d/d e1 of pair p | d/d e2 of pair p
    W.att = carve<float>(p, static_cast<size_t>(SG) * N);
    W.pred = carve<float>(p, static_cast<size_t>(G));
    W.stats = t->d_sums;
    W.bsum = t->d_sums + kSumDoubles / 2;
    W.loss = t->d_misc;
    W.losspart = t->d_misc + 1;
    W.grads = t->d_grads;
    W.adam_m = t->d_adam;
    W.adam_v = t->d_adam + P_TOTAL;
    W.head_grid = P.head_grid;

    // gradient partials and the segment table the optimiser sums them by
    float* pp = t->d_part;
    P.part_head = pp; pp += static_cast<size_t>(P.head_grid) * kHeadFloats;
    P.part_att = pp;  pp += static_cast<size_t>(P.head_grid) * 1024;
    P.part_end = pp;  pp += static_cast<size_t>(P.grid2) * 2048;
    for (int L = 0; L < 6; ++L) { P.part_conv[L] = pp; pp += static_cast<size_t>(nb) * conv_size(L); }
    int ns = 0;
    for (int L = 0; L < 6; ++L) W.seg[ns++] = Segment{conv_off(L), conv_size(L), nb, conv_size(L), P.part_conv[L]};
    W.seg[ns++] = Segment{P_ENDW, 2048, P.grid2, 2048, P.part_end};
    W.seg[ns++] = Segment{P_ATT, 1024, P.head_grid, 1024, P.part_att};
    W.seg[ns++] = Segment{P_NTNW, kHeadFloats, P.head_grid, kHeadFloats, P.part_head};
    W.nseg = ns;
    P.W = W;
    P.fwd_smem = fwd_layout(nmax, KS).total;
    P.bwd_smem = bwd_layout(nmax, KS).total;
    P.end_fwd_smem = (64 * 32 + 2 * static_cast<size_t>(nmax) * XS + 256) * 4 + kWarps * 128 * 8;
    P.end_bwd_smem = (32 * 64 + static_cast<size_t>(nmax) * XS + static_cast<size_t>(nmax) * 36 + 256 + 192) * 4 + 4 * 128 * 8;
    P.valid = true;
    return SGPR_OK;
}

#define BY_NPL(FN, ...)                                                   \
    do {                                                                  \
        if (P.npl == 1) FN<1>(__VA_ARGS__);                               \
        else if (P.npl == 2) FN<2>(__VA_ARGS__);                          \
        else FN<4>(__VA_ARGS__);                                          \
        t->launches += 1;                                                 \
    } while (0)

// pack | EdgeConv fwd x3 | conv_end fwd  (statistics zeroed first)
int launch_forward(sgpr_train* t, cudaStream_t st) {
    Plan& P = t->plan;
    const TrainWs& W = P.W;
    TRY_CUDA(cudaMemsetAsync(t->d_sums, 0, kSumDoubles / 2 * sizeof(double), st));
    if (!t->wpk_valid) {
        SGPR_LAUNCH(sgpr_train_pack_kernel, 32, kThreads, 0, st, W);
        t->launches += 1;
        t->wpk_valid = true;
    }
    for (int l = 0; l < 3; ++l) BY_NPL(launch_edge_fwd, t->knn_ties, P.grid4, P.fwd_smem, st, W, l);
    BY_NPL(launch_end_fwd, P.grid2, P.end_fwd_smem, st, W);
    return SGPR_OK;
}

// conv_end BN + attention + pair head (+ loss, head bwd, attention bwd); mode: kHeadFused / kHeadForward / kHeadBackward
void launch_head(sgpr_train* t, cudaStream_t st, int mode) {
    Plan& P = t->plan;
    SGPR_LAUNCH(sgpr_train_pool_head_kernel, P.head_grid, kThreads, P.pool_smem, st, P.W, P.part_head, P.part_att, mode);
    t->launches += 1;
}

// conv_end bwd | EdgeConv bwd x3   (the caller zeroes the dbeta / dgamma sums before launch_head)
int launch_backward(sgpr_train* t, cudaStream_t st) {
    Plan& P = t->plan;
    const TrainWs& W = P.W;
    BY_NPL(launch_end_bwd, P.grid2, P.end_bwd_smem, st, W, P.part_end);
    for (int l = 2; l >= 0; --l) BY_NPL(launch_edge_bwd, P.grid4, P.bwd_smem, st, W, l, P.part_conv[l], P.part_conv[3 + l]);
    return SGPR_OK;
}
#undef BY_NPL

void launch_adam(sgpr_train* t, cudaStream_t st, int apply) {
    AdamArgs A{};
    A.lr = t->lr; A.wd = t->wd; A.b1 = t->b1; A.b2 = t->b2; A.eps = t->eps;
    const long long step = t->steps + 1;
    A.bc1 = static_cast<float>(1.0 - std::pow(static_cast<double>(t->b1), static_cast<double>(step)));
    A.bc2_sqrt = static_cast<float>(std::sqrt(1.0 - std::pow(static_cast<double>(t->b2), static_cast<double>(step))));
    A.apply = apply;
    SGPR_LAUNCH(sgpr_train_adam_kernel, (STATE_TOTAL + 31) / 32, kThreads, 0, st, t->plan.W, A);
    t->launches += 1;
    if (apply == kApplyAll) t->steps = step;
}

}  // namespace

extern "C" {

int sgpr_train_step(sgpr_train* t, const float* f1_dev, const float* f2_dev, const float* target_dev, int B, int N, int k,
                    float* loss_dev, float* pred_dev, int flags, void* stream) {
    if (!t) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_step: NULL context");
    const int apply = (flags & SGPR_TRAIN_APPLY) ? 1 : 0, mirrored = (flags & SGPR_TRAIN_MIRRORED) ? 1 : 0;
    if (!target_dev) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_step: NULL target pointer");
    Guard guard(t->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = make_plan(t, f1_dev, f2_dev, target_dev, B, N, k, mirrored, "sgpr_train_step");
    if (rc) return rc;
    TRY_CUDA(cudaMemsetAsync(t->d_sums + kSumDoubles / 2, 0, kSumDoubles / 2 * sizeof(double), st));
    rc = launch_forward(t, st);
    if (rc) return rc;
    launch_head(t, st, kHeadFused);
    rc = launch_backward(t, st);
    if (rc) return rc;
    launch_adam(t, st, apply ? kApplyAll : kApplyNone);
    const TrainWs& W = t->plan.W;
    if (loss_dev) TRY_CUDA(cudaMemcpyAsync(loss_dev, W.loss, sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (pred_dev) TRY_CUDA(cudaMemcpyAsync(pred_dev, W.pred, static_cast<size_t>(B) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRY_CUDA(cudaGetLastError());
    t->last = W;
    t->has_last = true;
    t->plan.backward_ready = false;
    return SGPR_OK;
}

int sgpr_train_forward(sgpr_train* t, const float* f1_dev, const float* f2_dev, int B, int N, int k, float* pred_dev,
                       float* att1_dev, float* att2_dev, int flags, void* stream) {
    if (!t) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_forward: NULL context");
    const int apply = (flags & SGPR_TRAIN_APPLY) ? 1 : 0, mirrored = (flags & SGPR_TRAIN_MIRRORED) ? 1 : 0;
    Guard guard(t->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = make_plan(t, f1_dev, f2_dev, nullptr, B, N, k, mirrored, "sgpr_train_forward");
    if (rc) return rc;
    rc = launch_forward(t, st);
    if (rc) return rc;
    launch_head(t, st, kHeadForward);
    if (apply) launch_adam(t, st, kApplyRunning);             // nn.BatchNorm updates its running statistics in forward
    const TrainWs& W = t->plan.W;
    if (pred_dev) TRY_CUDA(cudaMemcpyAsync(pred_dev, W.pred, static_cast<size_t>(B) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const size_t att_bytes = static_cast<size_t>(B) * N * sizeof(float);
    if (att1_dev) TRY_CUDA(cudaMemcpyAsync(att1_dev, W.att, att_bytes, cudaMemcpyDeviceToDevice, st));
    if (att2_dev && !mirrored) TRY_CUDA(cudaMemcpyAsync(att2_dev, W.att + static_cast<size_t>(B) * N, att_bytes, cudaMemcpyDeviceToDevice, st));
    TRY_CUDA(cudaGetLastError());
    t->last = W;
    t->has_last = true;
    t->plan.backward_ready = true;
    t->forward_gen += 1;
    return SGPR_OK;
}

int sgpr_train_backward(sgpr_train* t, const float* dpred_dev, float* grads_dev, void* stream) {
    if (!t || !dpred_dev) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_backward: NULL argument");
    if (!t->plan.valid || !t->plan.backward_ready)
        return sgpr_fail(SGPR_E_INVALID, "sgpr_train_backward: no sgpr_train_forward to differentiate (or another call used the workspace since)");
    Guard guard(t->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    t->plan.W.dpred = dpred_dev;
    TRY_CUDA(cudaMemsetAsync(t->d_sums + kSumDoubles / 2, 0, kSumDoubles / 2 * sizeof(double), st));
    launch_head(t, st, kHeadBackward);
    int rc = launch_backward(t, st);
    if (rc) return rc;
    launch_adam(t, st, kApplyNone);                            // sums the partials into the flat gradient vector
    if (grads_dev) TRY_CUDA(cudaMemcpyAsync(grads_dev, t->d_grads, P_TOTAL * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRY_CUDA(cudaGetLastError());
    t->last = t->plan.W;
    return SGPR_OK;
}

int sgpr_train_set_state_dev(sgpr_train* t, const float* state_dev, void* stream) {
    if (!t || !state_dev) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_set_state_dev: NULL argument");
    Guard guard(t->device);
    TRY_CUDA(cudaMemcpyAsync(t->d_state, state_dev, STATE_TOTAL * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    t->has_state = true;
    t->wpk_valid = false;
    return SGPR_OK;
}

int sgpr_train_get_state_dev(sgpr_train* t, float* state_dev, void* stream) {
    if (!t || !state_dev) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_get_state_dev: NULL argument");
    if (!t->has_state) return sgpr_fail(SGPR_E_NOWEIGHTS, "sgpr_train_get_state_dev: no state has been set");
    Guard guard(t->device);
    TRY_CUDA(cudaMemcpyAsync(state_dev, t->d_state, STATE_TOTAL * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return SGPR_OK;
}

int sgpr_train_assemble(sgpr_train* t, const float* graphs_dev, int M, int N, const int32_t* pair_idx_dev, int P,
                        uint64_t seed, uint64_t step, float* out_f1_dev, float* draws_dev, float* jitter_dev, void* stream) {
    if (!t) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_assemble: NULL context");
    if (P < 0 || M < 1) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_assemble: bad sizes (M=%d, P=%d)", M, P);
    if (N < 2 || N > SGPR_MAX_NODES) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_assemble: node_num %d outside [2,%d]", N, SGPR_MAX_NODES);
    if (P == 0) return SGPR_OK;
    if (!graphs_dev || !pair_idx_dev || !out_f1_dev) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_assemble: NULL pointer");
    Guard guard(t->device);
    AssembleArgs A{};
    A.graphs = graphs_dev; A.pair_idx = pair_idx_dev; A.out = out_f1_dev; A.draws = draws_dev; A.jitter = jitter_dev;
    A.M = M; A.N = N; A.P = P;
    A.seed_lo = static_cast<uint32_t>(seed); A.seed_hi = static_cast<uint32_t>(seed >> 32);
    A.step_lo = static_cast<uint32_t>(step); A.step_hi = static_cast<uint32_t>(step >> 32);
    const int grid = 2 * P < 8 * t->sm_count ? 2 * P : 8 * t->sm_count;
    SGPR_LAUNCH(sgpr_train_assemble_kernel, grid, kThreads, 0, static_cast<cudaStream_t>(stream), A);
    t->launches += 1;
    TRY_CUDA(cudaGetLastError());
    return SGPR_OK;
}

int sgpr_train_get_grads(sgpr_train* t, float* grads_host) {
    if (!t || !grads_host) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_get_grads: NULL argument");
    Guard guard(t->device);
    TRY_CUDA(cudaDeviceSynchronize());
    TRY_CUDA(cudaMemcpy(grads_host, t->d_grads, P_TOTAL * sizeof(float), cudaMemcpyDeviceToHost));
    return SGPR_OK;
}

int64_t sgpr_train_forward_generation(const sgpr_train* t) { return t ? t->forward_gen : 0; }

int sgpr_train_set_knn_ties(sgpr_train* t, int mode) {
    if (!t) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_set_knn_ties: NULL context");
    if (mode != SGPR_TIES_CUDA && mode != SGPR_TIES_CPU)
        return sgpr_fail(SGPR_E_INVALID, "sgpr_train_set_knn_ties: mode %d is neither SGPR_TIES_CUDA (0) nor SGPR_TIES_CPU (1)", mode);
    t->knn_ties = mode;
    return SGPR_OK;
}
int sgpr_train_get_knn_ties(const sgpr_train* t) { return t ? t->knn_ties : SGPR_E_INVALID; }

int64_t sgpr_train_step_count(const sgpr_train* t) { return t ? t->steps : 0; }
int64_t sgpr_train_launch_count(const sgpr_train* t) { return t ? t->launches : 0; }

int64_t sgpr_train_debug_read(sgpr_train* t, const char* what, int layer, void* host, int64_t cap_bytes) {
    if (!t || !what || !host) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_debug_read: NULL argument");
    if (!t->has_last) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_debug_read: no step has run yet");
    const TrainWs& W = t->last;
    const size_t SG = static_cast<size_t>(W.S) * W.G;
    const void* src = nullptr;
    size_t bytes = 0;
    const bool per_layer = !strcmp(what, "yext") || !strcmp(what, "a") || !strcmp(what, "d") || !strcmp(what, "sumy") ||
                           !strcmp(what, "gz") || !strcmp(what, "enode") || !strcmp(what, "idx");
    if (per_layer) {
        if (layer < 0 || layer > 5) return sgpr_fail(SGPR_E_INVALID, "sgpr_train_debug_read: layer %d outside [0,5]", layer);
        const size_t per = SG * W.N * layer_cout(layer);
        if (!strcmp(what, "yext")) { src = W.yext[layer]; bytes = per * 4; }
        else if (!strcmp(what, "a")) { src = W.a[layer]; bytes = per * 4; }
        else if (!strcmp(what, "d")) { src = W.d[layer]; bytes = per * 4; }
        else if (!strcmp(what, "sumy")) { src = W.sumy[layer]; bytes = per * 4; }
        else if (!strcmp(what, "gz")) { src = W.gz[layer]; bytes = per * 4; }
        else if (!strcmp(what, "enode")) { src = W.enode[layer]; bytes = per; }
        else { src = W.idx[layer]; bytes = SG * W.N * W.k; }
    } else if (!strcmp(what, "yend")) { src = W.yend; bytes = SG * W.N * 32 * 4; }
    else if (!strcmp(what, "gzend")) { src = W.gzend; bytes = SG * W.N * 32 * 4; }
    else if (!strcmp(what, "pooled")) { src = W.pooled; bytes = SG * 32 * 4; }
    else if (!strcmp(what, "dpooled")) { src = W.dpooled; bytes = 2 * static_cast<size_t>(W.G) * 32 * 4; }
    else if (!strcmp(what, "att")) { src = W.att; bytes = SG * W.N * 4; }
    else if (!strcmp(what, "stats")) { src = W.stats; bytes = kSumDoubles / 2 * 8; }
    else if (!strcmp(what, "bsum")) { src = W.bsum; bytes = kSumDoubles / 2 * 8; }
    else return sgpr_fail(SGPR_E_INVALID, "sgpr_train_debug_read: unknown tensor '%s'", what);
    if (static_cast<int64_t>(bytes) > cap_bytes)
        return sgpr_fail(SGPR_E_INVALID, "sgpr_train_debug_read: '%s' needs %zu bytes, buffer has %lld", what, bytes, (long long)cap_bytes);
    Guard guard(t->device);
    TRY_CUDA(cudaDeviceSynchronize());
    TRY_CUDA(cudaMemcpy(host, src, bytes, cudaMemcpyDeviceToHost));
    return static_cast<int64_t>(bytes);
}

}  // extern "C"
