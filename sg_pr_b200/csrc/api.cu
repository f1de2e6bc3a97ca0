// C-ABI implementation (include/sgpr_b200.h): context, weight packing/upload, kernel launches.
// No torch types, links only cudart.  There is deliberately no CPU path: every entry point needs the device.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/sgpr_b200.h"
#include "common.cuh"
#include "embed_kernel.cuh"
#include "head_kernels.cuh"
#include "launchers.hpp"
#include "pack.hpp"

using namespace sgpr;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void set_error_text(const char* text) { snprintf(g_err, sizeof(g_err), "%s", text); }

#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) return fail(SGPR_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

template <typename T>
int ensure(T*& ptr, size_t& cap, size_t need) {
    if (need <= cap) return SGPR_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ptr), want * sizeof(T));
    if (e != cudaSuccess) return fail(SGPR_E_CUDA, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
    cap = want;
    return SGPR_OK;
}

}  // namespace

// error reporting for the other translation units of the library (train.cu)
int sgpr_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    set_error_text(buf);
    return code;
}

struct sgpr_ctx {
    int device = 0;
    int sm_count = 0;
    bool has_weights = false;
    float* d_blob = nullptr;
    PackOffsets off{};
    PackedWeights pw{};
    HeadParams hp{};
    int* d_counters = nullptr;   size_t counters_cap = 0;
    float* d_pooled = nullptr;   size_t pooled_cap = 0;
    float* d_in = nullptr;       size_t in_cap = 0;      // host-path staging: f1 | f2
    float* d_out = nullptr;      size_t out_cap = 0;     // host-path staging: score | att1 | att2
    float* d_proj = nullptr;     size_t proj_cap = 0;
    float* d_blk = nullptr;      size_t blk_cap = 0;     // rowblk | colblk
    cudaStream_t stream = nullptr;                      // host-path stream
    long long launches = 0;
    int* d_order = nullptr;      size_t order_cap = 0;   // rows[G] | order[G]
    float* d_halves = nullptr;   size_t halves_cap = 0;  // branch-split launches: [G][2][N][32]
    int* d_gctr = nullptr;       size_t gctr_cap = 0;    //   and their per-graph arrival counters (zero between launches)
    int split_max = -1;          // largest G launched branch-split (-1: two thirds of the resident CTA slots)
    int split = 1;               // branch-split launches while the graphs fit the resident CTAs; SGPR_NO_SPLIT=1 disables it
    int* d_ctrs = nullptr;                               // {done counter, work counter}
    int zerocopy = 1;            // host entry point: read pinned buffers in place; SGPR_NO_ZEROCOPY=1 forces staged copies
    int balance = 1;             // order graphs by active rows before the fused kernel; SGPR_NO_BALANCE=1 disables it
    int dedup = 1;               // collapse trailing all-zero nodes (exact); SGPR_NO_DEDUP=1 disables it for experiments
    float* d_fc1_planes = nullptr;   // [2][16][32] FC1 weights as UMMA B operand (tcgen05 score matrix, version 2)
    int scoremat_version = 1;    // tcgen05 score-matrix kernel: 1 = FFMA2 epilogue (default, faster), 2 = FC1 on the tensor cores
                                 // too, A operand in TMEM (SGPR_SCOREMAT_V2=1; see scoremat_umma.cuh for the measurements)
    int embed_tc = 0;            // N <= 64: the tensor-core variant of the fused kernel (SGPR_EMBED_TC=1)
    int scoremat_ffma = 0;       // score matrix on fp32 FMA instead of tcgen05 (SGPR_SCOREMAT_FFMA=1; always in tests/emu)
    int knn_ties = SGPR_TIES_CUDA;   // k-NN tie rule (sgpr_set_knn_ties); SGPR_KNN_TIES=cpu|cuda sets the initial value
};

extern "C" {

int sgpr_abi_version(void) { return SGPR_ABI_VERSION; }

const char* sgpr_last_error(void) { return g_err; }

int sgpr_create(sgpr_ctx** out, int device) {
    if (!out) return fail(SGPR_E_INVALID, "sgpr_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SGPR_E_CUDA, "sgpr_create: no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(SGPR_E_INVALID, "sgpr_create: device %d out of range [0,%d)", device, count);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(SGPR_E_CUDA, "sgpr_create: cudaSetDevice(%d) failed", device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SGPR_E_CUDA, "sgpr_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    sgpr_ctx* ctx = new (std::nothrow) sgpr_ctx();
    if (!ctx) return fail(SGPR_E_CUDA, "sgpr_create: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->off = make_offsets();
    if (const char* nd = getenv("SGPR_NO_DEDUP")) ctx->dedup = (nd[0] == '1') ? 0 : 1;
    if (const char* nz = getenv("SGPR_NO_ZEROCOPY")) ctx->zerocopy = (nz[0] == '1') ? 0 : 1;
    if (const char* nb = getenv("SGPR_NO_BALANCE")) ctx->balance = (nb[0] == '1') ? 0 : 1;
    if (const char* ns = getenv("SGPR_NO_SPLIT")) ctx->split = (ns[0] == '1') ? 0 : 1;
    if (const char* sm = getenv("SGPR_SPLIT_MAX")) ctx->split_max = atoi(sm);
    if (const char* v2 = getenv("SGPR_SCOREMAT_V2")) ctx->scoremat_version = (v2[0] == '1') ? 2 : 1;
    if (const char* et = getenv("SGPR_EMBED_TC")) ctx->embed_tc = (et[0] == '1') ? 1 : 0;
    if (const char* sf = getenv("SGPR_SCOREMAT_FFMA")) ctx->scoremat_ffma = (sf[0] == '1') ? 1 : 0;
    if (const char* kt = getenv("SGPR_KNN_TIES")) ctx->knn_ties = (strcmp(kt, "cpu") == 0) ? SGPR_TIES_CPU : SGPR_TIES_CUDA;
    // opt in to the full shared-memory carve-out (the per-NPL objects subtract each kernel's static __shared__ bytes)
    const int optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    e = embed_optin<1>(optin);
    if (e == cudaSuccess) e = embed_optin<2>(optin);
    if (e == cudaSuccess) e = embed_optin<4>(optin);
    if (e == cudaSuccess) e = embed_tc_optin(optin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgpr_score_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
#ifndef SGPR_EMU
    if (e == cudaSuccess) e = score_matrix_umma_optin();
#endif
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_blob), ctx->off.total * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_fc1_planes), 1024 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_ctrs), 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_ctrs, 0, 2 * sizeof(int));
    if (e != cudaSuccess) {
        int rc = fail(SGPR_E_CUDA, "sgpr_create: %s", cudaGetErrorString(e));
        sgpr_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SGPR_OK;
}

int sgpr_destroy(sgpr_ctx* ctx) {
    if (!ctx) return SGPR_OK;
    DeviceGuard guard(ctx->device);
    cudaFree(ctx->d_blob);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_pooled);
    cudaFree(ctx->d_in);
    cudaFree(ctx->d_out);
    cudaFree(ctx->d_proj);
    cudaFree(ctx->d_blk);
    cudaFree(ctx->d_order);
    cudaFree(ctx->d_halves);
    cudaFree(ctx->d_gctr);
    cudaFree(ctx->d_ctrs);
    cudaFree(ctx->d_fc1_planes);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SGPR_OK;
}

int sgpr_set_weights(sgpr_ctx* ctx, const sgpr_weights* w) {
    if (!ctx || !w) return fail(SGPR_E_INVALID, "sgpr_set_weights: NULL argument");
    if (w->filters[0] != SGPR_FILTERS_1 || w->filters[1] != SGPR_FILTERS_2 || w->filters[2] != SGPR_FILTERS_3 ||
        w->tensor_neurons != SGPR_TENSOR_NEURONS || w->bottleneck != SGPR_BOTTLENECK)
        return fail(SGPR_E_ARCH,
                    "sgpr_set_weights: kernels are built for filters 64/64/32, tensor_neurons 16, bottle_neck 16 "
                    "(got %d/%d/%d, %d, %d)",
                    w->filters[0], w->filters[1], w->filters[2], w->tensor_neurons, w->bottleneck);
    const void* must[] = {w->s_conv_w[0], w->s_conv_w[1], w->s_conv_w[2], w->f_conv_w[0], w->f_conv_w[1], w->f_conv_w[2],
                          w->end_conv_w, w->att_w, w->ntn_w, w->ntn_v, w->ntn_b, w->fc1_w, w->fc1_b, w->fc2_w, w->fc2_b,
                          w->end_bn.weight, w->end_bn.bias, w->end_bn.running_mean, w->end_bn.running_var};
    for (const void* p : must)
        if (!p) return fail(SGPR_E_INVALID, "sgpr_set_weights: a weight pointer is NULL");
    for (int l = 0; l < 3; ++l) {
        const sgpr_bn* bns[2] = {&w->s_bn[l], &w->f_bn[l]};
        for (const sgpr_bn* b : bns)
            if (!b->weight || !b->bias || !b->running_mean || !b->running_var)
                return fail(SGPR_E_INVALID, "sgpr_set_weights: a BatchNorm pointer is NULL");
    }
    DeviceGuard guard(ctx->device);
    std::vector<float> blob;
    pack_weights(*w, blob, ctx->hp, ctx->off);
    // synchronous copy: the host vector dies at return, and weights must be visible to every stream afterwards
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(ctx->d_blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
    const float* b = ctx->d_blob;
    const PackOffsets& o = ctx->off;
    ctx->pw = PackedWeights{b + o.s1,    b + o.w_s2,  b + o.w_s3,  b + o.w_f1,  b + o.w_f2,  b + o.w_f3,
                            b + o.w_end, b + o.ab_s2, b + o.ab_s3, b + o.ab_f1, b + o.ab_f2, b + o.ab_f3,
                            b + o.ab_end, b + o.att_w, b + o.ntn_w, b + o.ntn_v, b + o.ntn_b,
                            b + o.wtc_s2, b + o.wtc_s3, b + o.wtc_f2, b + o.wtc_f3};
#ifndef SGPR_EMU
    {
        float planes[1024];
        score_matrix_fc1_planes(w->fc1_w, planes);
        CUDA_TRY(cudaMemcpy(ctx->d_fc1_planes, planes, sizeof(planes), cudaMemcpyHostToDevice));
    }
#endif
    ctx->has_weights = true;
    return SGPR_OK;
}

}  // extern "C"

namespace {

int check_shape(const char* who, int count, int N, int k) {
    if (count < 0) return fail(SGPR_E_INVALID, "%s: negative batch %d", who, count);
    if (N < 1 || N > SGPR_MAX_NODES) return fail(SGPR_E_INVALID, "%s: node_num %d outside [1,%d]", who, N, SGPR_MAX_NODES);
    if (k < 1 || k > N)
        return fail(SGPR_E_INVALID, "%s: k=%d must satisfy 1 <= k <= node_num=%d (topk raises in the reference, dgcnn.py:19)",
                    who, k, N);
    return SGPR_OK;
}

// Pinned (page-locked) host memory is mapped into the device address space under UVA: the kernel can pull each
// graph's 60*N-byte block straight over PCIe with its bulk-TMA load.  Returns the device-visible alias or nullptr.
const void* device_alias_of_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeHost && at.devicePointer) return at.devicePointer;
    return nullptr;
}

bool is_device_memory(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int launch_embed(sgpr_ctx* ctx, EmbedArgs a, cudaStream_t st) {
    const int N = a.N;
    a.KS = (a.k + 3) & ~3;
    a.dedup = ctx->dedup;
    const int npl = (N <= 32) ? 1 : (N <= 64) ? 2 : 4;
    const SmemLayout L = make_layout(32 * npl, a.KS);
    const int per_sm = (npl <= 2) ? 2 : 1;
    const int capacity = ctx->sm_count * per_sm;
    int grid = a.G < capacity ? a.G : capacity;
    if (grid < 1) return SGPR_OK;
    a.order = nullptr;
    a.work_ctr = nullptr;
    a.split = 0;
    a.sm_count = ctx->sm_count;
    a.halves = nullptr;
    a.gctr = nullptr;
    // Small launches run one BRANCH of a graph per work unit: the xyz and the semantic EdgeConv stacks are independent
    // until conv_end, so a graph's critical path halves (embed_kernel.cuh, EmbedArgs::split).  It pays while the 2G units
    // stay within ~4/3 of the resident CTA slots (B200, N = 64: 37 vs 57 us at 32 graphs, 79 vs 86 at 192, 94 vs 86 at
    // 224 — profiles/experiments/r02_embed_branch_split.txt).  Bit-identical; the debug taps keep the whole-graph form.
    const int split_max = ctx->split_max >= 0 ? ctx->split_max : (2 * capacity) / 3;
    if (ctx->split && a.G <= split_max && !a.trace_knn && !a.trace_layers && !(ctx->embed_tc && npl == 2)) {
        int rc = ensure(ctx->d_halves, ctx->halves_cap, static_cast<size_t>(a.G) * 2 * N * kF3);
        if (rc) return rc;
        if (static_cast<size_t>(a.G) > ctx->gctr_cap) {
            rc = ensure(ctx->d_gctr, ctx->gctr_cap, static_cast<size_t>(a.G));
            if (rc) return rc;
            cudaError_t e = cudaMemsetAsync(ctx->d_gctr, 0, ctx->gctr_cap * sizeof(int), st);
            if (e != cudaSuccess) return fail(SGPR_E_CUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(e));
        }
        a.split = 1;
        a.halves = ctx->d_halves;
        a.gctr = ctx->d_gctr;
        grid = 2 * a.G < capacity ? 2 * a.G : capacity;
    }
    // persistent launches (more graphs than resident CTAs): pop the graphs heaviest-first (LPT) so that the tail of the
    // launch is short; results are unchanged.  Not worth its pre-pass while every graph has its own resident CTA: the
    // hardware's CTA->SM placement cannot be steered (profiles/r01_variants_timeline.txt), and claiming graphs by the SM a
    // CTA lands on so that heavy graphs share an SM with light ones balances the row sums but not the launch time
    // (profiles/experiments/r02_embed_sm_pairing.txt).  Skipped when the graphs sit in pinned host memory (the pre-pass
    // would pull them over PCIe a second time).
    if (ctx->balance && !a.compact && a.G > capacity && is_device_memory(a.g0)) {
        int rc = ensure(ctx->d_order, ctx->order_cap, static_cast<size_t>(2) * a.G);
        if (rc) return rc;
        const int resident = (a.G <= capacity) ? 1 : 0;
        int ogrid = (a.G + kWarps - 1) / kWarps;
        if (ogrid > 4 * ctx->sm_count) ogrid = 4 * ctx->sm_count;
        SGPR_LAUNCH(sgpr_order_kernel, ogrid, kThreads, 0, st, a.g0, a.g1, a.pairs, a.G, a.N, a.dedup, ctx->sm_count, resident,
                    ctx->d_order, ctx->d_order + a.G, ctx->d_ctrs, ctx->d_ctrs + 1);
        ctx->launches += 1;
        a.order = ctx->d_order + a.G;
        if (!resident) a.work_ctr = ctx->d_ctrs + 1;
    }
    if (ctx->embed_tc && npl == 2 && ctx->knn_ties == SGPR_TIES_CUDA && !a.trace_knn && !a.trace_layers) {
        embed_tc_launch(grid, st, a, ctx->pw, ctx->hp);
        ctx->launches += 1;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(SGPR_E_CUDA, "embed (tensor-core) kernel launch failed: %s", cudaGetErrorString(e));
        return SGPR_OK;
    }
    switch (npl) {
        case 1: embed_launch<1>(ctx->knn_ties, grid, L.total, st, a, ctx->pw, ctx->hp); break;
        case 2: embed_launch<2>(ctx->knn_ties, grid, L.total, st, a, ctx->pw, ctx->hp); break;
        default: embed_launch<4>(ctx->knn_ties, grid, L.total, st, a, ctx->pw, ctx->hp); break;
    }
    ctx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SGPR_E_CUDA, "embed kernel launch failed: %s", cudaGetErrorString(e));
    return SGPR_OK;
}

int forward_pairs_impl(sgpr_ctx* ctx, const float* f1, const float* f2, int B, int N, int k, float* score, float* att1,
                       float* att2, cudaStream_t st, int compact = 0) {
    int rc = ensure(ctx->d_pooled, ctx->pooled_cap, static_cast<size_t>(2) * B * kF3);
    if (rc) return rc;
    if (static_cast<size_t>(B) > ctx->counters_cap) {
        rc = ensure(ctx->d_counters, ctx->counters_cap, static_cast<size_t>(B));
        if (rc) return rc;
        CUDA_TRY(cudaMemsetAsync(ctx->d_counters, 0, ctx->counters_cap * sizeof(int), st));
    }
    EmbedArgs a{};
    a.g0 = f1; a.g1 = f2; a.G = 2 * B; a.N = N; a.k = k; a.pairs = 1; a.compact = compact;
    a.pooled = ctx->d_pooled; a.att0 = att1; a.att1 = att2; a.emb = nullptr;
    a.score = score; a.counters = ctx->d_counters; a.trace_knn = nullptr; a.trace_layers = nullptr;
    return launch_embed(ctx, a, st);
}

}  // namespace

extern "C" {

int sgpr_forward_pairs(sgpr_ctx* ctx, const float* f1_dev, const float* f2_dev, int B, int N, int k, float* score_dev,
                       float* att1_dev, float* att2_dev, void* stream) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_forward_pairs: ctx is NULL");
    if (!ctx->has_weights) return fail(SGPR_E_NOWEIGHTS, "sgpr_forward_pairs: call sgpr_set_weights first");
    int rc = check_shape("sgpr_forward_pairs", B, N, k);
    if (rc) return rc;
    if (B == 0) return SGPR_OK;
    if (!f1_dev || !f2_dev || !score_dev) return fail(SGPR_E_INVALID, "sgpr_forward_pairs: NULL feature/score pointer");
    DeviceGuard guard(ctx->device);
    return forward_pairs_impl(ctx, f1_dev, f2_dev, B, N, k, score_dev, att1_dev, att2_dev, static_cast<cudaStream_t>(stream));
}

size_t sgpr_compact_stride(int N) { return (static_cast<size_t>(N) * 13 + 15) & ~static_cast<size_t>(15); }

int sgpr_compact_from_blocks(const float* blocks_host, int G, int N, void* records_host) {
    if (G < 0 || N < 1 || N > SGPR_MAX_NODES) return fail(SGPR_E_INVALID, "sgpr_compact_from_blocks: G=%d N=%d out of range", G, N);
    if (G == 0) return SGPR_OK;
    if (!blocks_host || !records_host) return fail(SGPR_E_INVALID, "sgpr_compact_from_blocks: NULL pointer");
    const size_t stride = sgpr_compact_stride(N);
    unsigned char* out = static_cast<unsigned char*>(records_host);
    const uint32_t kOne = 0x3f800000u;                        // 1.0f
    uint32_t bad = 0;
    for (int g = 0; g < G; ++g) {
        const float* blk = blocks_host + static_cast<size_t>(g) * kInCh * N;
        unsigned char* rec = out + static_cast<size_t>(g) * stride;
        std::memcpy(rec, blk, static_cast<size_t>(12) * N);                      // xyz rows, bit for bit
        // plain integer loops over a row (the compiler vectorises them): per node the number of 1.0f, the label they
        // name, and any bits that are neither +0.0f nor 1.0f
        uint32_t label[SGPR_MAX_NODES], ones[SGPR_MAX_NODES], other[SGPR_MAX_NODES];
        for (int n = 0; n < N; ++n) { label[n] = 0u; ones[n] = 0u; other[n] = 0u; }
        for (int c = 0; c < kLabels; ++c) {
            uint32_t row[SGPR_MAX_NODES];
            std::memcpy(row, blk + static_cast<size_t>(3 + c) * N, static_cast<size_t>(4) * N);
            const uint32_t cc = static_cast<uint32_t>(c);
            for (int n = 0; n < N; ++n) {
                const uint32_t w = row[n];
                const uint32_t one = (w == kOne) ? 1u : 0u;
                ones[n] += one;
                label[n] += one * cc;
                other[n] |= w ^ ((0u - one) & kOne);
            }
        }
        unsigned char* lab = rec + static_cast<size_t>(12) * N;
        for (int n = 0; n < N; ++n) {
            bad |= other[n] | ((ones[n] > 1u) ? 1u : 0u);
            lab[n] = static_cast<unsigned char>(ones[n] ? label[n] : 255u);
        }
        std::memset(rec + static_cast<size_t>(13) * N, 0, stride - static_cast<size_t>(13) * N);
    }
    return bad ? 1 : SGPR_OK;
}

int sgpr_forward_pairs_compact(sgpr_ctx* ctx, const void* g1, const void* g2, int B, int N, int k, float* score_dev,
                               float* att1_dev, float* att2_dev, void* stream) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_forward_pairs_compact: ctx is NULL");
    if (!ctx->has_weights) return fail(SGPR_E_NOWEIGHTS, "sgpr_forward_pairs_compact: call sgpr_set_weights first");
    int rc = check_shape("sgpr_forward_pairs_compact", B, N, k);
    if (rc) return rc;
    if (B == 0) return SGPR_OK;
    if (!g1 || !g2 || !score_dev) return fail(SGPR_E_INVALID, "sgpr_forward_pairs_compact: NULL graph/score pointer");
    if ((reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(g2)) & 15u)
        return fail(SGPR_E_INVALID, "sgpr_forward_pairs_compact: graph records must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    return forward_pairs_impl(ctx, static_cast<const float*>(g1), static_cast<const float*>(g2), B, N, k, score_dev, att1_dev,
                              att2_dev, static_cast<cudaStream_t>(stream), static_cast<int>(sgpr_compact_stride(N)));
}

int sgpr_embed_compact(sgpr_ctx* ctx, const void* graphs, int M, int N, int k, float* pooled_dev, float* att_dev,
                       float* emb_dev, void* stream) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_embed_compact: ctx is NULL");
    if (!ctx->has_weights) return fail(SGPR_E_NOWEIGHTS, "sgpr_embed_compact: call sgpr_set_weights first");
    int rc = check_shape("sgpr_embed_compact", M, N, k);
    if (rc) return rc;
    if (M == 0) return SGPR_OK;
    if (!graphs || !pooled_dev) return fail(SGPR_E_INVALID, "sgpr_embed_compact: NULL graphs/pooled pointer");
    if (reinterpret_cast<uintptr_t>(graphs) & 15u)
        return fail(SGPR_E_INVALID, "sgpr_embed_compact: graph records must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    EmbedArgs a{};
    a.g0 = static_cast<const float*>(graphs); a.g1 = nullptr; a.G = M; a.N = N; a.k = k; a.pairs = 0;
    a.compact = static_cast<int>(sgpr_compact_stride(N));
    a.pooled = pooled_dev; a.att0 = att_dev; a.att1 = nullptr; a.emb = emb_dev;
    return launch_embed(ctx, a, static_cast<cudaStream_t>(stream));
}

int sgpr_forward_pairs_host(sgpr_ctx* ctx, const float* f1_host, const float* f2_host, int B, int N, int k,
                            float* score_host, float* att1_host, float* att2_host) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_forward_pairs_host: ctx is NULL");
    if (!ctx->has_weights) return fail(SGPR_E_NOWEIGHTS, "sgpr_forward_pairs_host: call sgpr_set_weights first");
    int rc = check_shape("sgpr_forward_pairs_host", B, N, k);
    if (rc) return rc;
    if (B == 0) return SGPR_OK;
    if (!f1_host || !f2_host || !score_host) return fail(SGPR_E_INVALID, "sgpr_forward_pairs_host: NULL feature/score pointer");
    DeviceGuard guard(ctx->device);
    const size_t side = static_cast<size_t>(B) * kInCh * N;
    const size_t att = static_cast<size_t>(B) * N;
    cudaStream_t st = ctx->stream;
    // ---- zero-copy: every buffer pinned -> the kernel reads the inputs and writes the results in place ----
    if (ctx->zerocopy) {
        const float* d1 = static_cast<const float*>(device_alias_of_pinned(f1_host));
        const float* d2 = static_cast<const float*>(device_alias_of_pinned(f2_host));
        float* ds = static_cast<float*>(const_cast<void*>(device_alias_of_pinned(score_host)));
        float* da1 = att1_host ? static_cast<float*>(const_cast<void*>(device_alias_of_pinned(att1_host))) : nullptr;
        float* da2 = att2_host ? static_cast<float*>(const_cast<void*>(device_alias_of_pinned(att2_host))) : nullptr;
        if (d1 && d2 && ds && (!att1_host || da1) && (!att2_host || da2)) {
            rc = forward_pairs_impl(ctx, d1, d2, B, N, k, ds, da1, da2, st);
            if (rc) return rc;
            CUDA_TRY(cudaStreamSynchronize(st));
            return SGPR_OK;
        }
    }
    // ---- staged: H2D copies, kernel, D2H copies ----
    rc = ensure(ctx->d_in, ctx->in_cap, 2 * side);
    if (rc) return rc;
    rc = ensure(ctx->d_out, ctx->out_cap, static_cast<size_t>(B) + 2 * att);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_in, f1_host, side * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_in + side, f2_host, side * sizeof(float), cudaMemcpyHostToDevice, st));
    float* d_score = ctx->d_out;
    float* d_att1 = ctx->d_out + B;
    float* d_att2 = d_att1 + att;
    rc = forward_pairs_impl(ctx, ctx->d_in, ctx->d_in + side, B, N, k, d_score, att1_host ? d_att1 : nullptr,
                            att2_host ? d_att2 : nullptr, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(score_host, d_score, static_cast<size_t>(B) * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (att1_host) CUDA_TRY(cudaMemcpyAsync(att1_host, d_att1, att * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (att2_host) CUDA_TRY(cudaMemcpyAsync(att2_host, d_att2, att * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SGPR_OK;
}

int sgpr_embed_trace(sgpr_ctx* ctx, const float* graphs_dev, int M, int N, int k, float* pooled_dev, float* att_dev,
                     float* emb_dev, uint8_t* knn_dev, float* layer_out_dev, void* stream) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_embed: ctx is NULL");
    if (!ctx->has_weights) return fail(SGPR_E_NOWEIGHTS, "sgpr_embed: call sgpr_set_weights first");
    int rc = check_shape("sgpr_embed", M, N, k);
    if (rc) return rc;
    if (M == 0) return SGPR_OK;
    if (!graphs_dev || !pooled_dev) return fail(SGPR_E_INVALID, "sgpr_embed: NULL graphs/pooled pointer");
    DeviceGuard guard(ctx->device);
    EmbedArgs a{};
    a.g0 = graphs_dev; a.g1 = nullptr; a.G = M; a.N = N; a.k = k; a.pairs = 0;
    a.pooled = pooled_dev; a.att0 = att_dev; a.att1 = nullptr; a.emb = emb_dev;
    a.score = nullptr; a.counters = nullptr; a.trace_knn = knn_dev; a.trace_layers = layer_out_dev;
    return launch_embed(ctx, a, static_cast<cudaStream_t>(stream));
}

int sgpr_embed(sgpr_ctx* ctx, const float* graphs_dev, int M, int N, int k, float* pooled_dev, float* att_dev,
               float* emb_dev, void* stream) {
    return sgpr_embed_trace(ctx, graphs_dev, M, N, k, pooled_dev, att_dev, emb_dev, nullptr, nullptr, stream);
}

int sgpr_score_pairs(sgpr_ctx* ctx, const float* pooled_dev, const int32_t* pair_idx_dev, int P, float* score_dev,
                     void* stream) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_score_pairs: ctx is NULL");
    if (!ctx->has_weights) return fail(SGPR_E_NOWEIGHTS, "sgpr_score_pairs: call sgpr_set_weights first");
    if (P < 0) return fail(SGPR_E_INVALID, "sgpr_score_pairs: negative pair count");
    if (P == 0) return SGPR_OK;
    if (!pooled_dev || !pair_idx_dev || !score_dev) return fail(SGPR_E_INVALID, "sgpr_score_pairs: NULL pointer");
    DeviceGuard guard(ctx->device);
    const int grid = P < ctx->sm_count * 8 ? P : ctx->sm_count * 8;
    SGPR_LAUNCH(sgpr_score_pairs_kernel, grid, kThreads, 0, static_cast<cudaStream_t>(stream), pooled_dev, pair_idx_dev, P,
                score_dev, ctx->pw, ctx->hp);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return SGPR_OK;
}

static int score_matrix_impl(sgpr_ctx* ctx, const float* pooled_rows_dev, int R, const float* pooled_cols_dev, int M,
                             float* scores_dev, int64_t ld_scores, void* stream, float* const* outs, int n_out) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_score_matrix: ctx is NULL");
    if (!ctx->has_weights) return fail(SGPR_E_NOWEIGHTS, "sgpr_score_matrix: call sgpr_set_weights first");
    if (R < 0 || M < 0) return fail(SGPR_E_INVALID, "sgpr_score_matrix: negative size");
    if (R == 0 || M == 0) return SGPR_OK;
    if (ld_scores < M) return fail(SGPR_E_INVALID, "sgpr_score_matrix: ld_scores %lld < M %d", (long long)ld_scores, M);
    if (!pooled_rows_dev || !pooled_cols_dev || (!scores_dev && n_out == 0)) return fail(SGPR_E_INVALID, "sgpr_score_matrix: NULL pointer");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#ifdef SGPR_EMU
    const bool ffma = true;
#else
    const bool ffma = ctx->scoremat_ffma != 0;
#endif
    if (ffma && n_out > 0)
        return fail(SGPR_E_INVALID, "sgpr_score_matrix_multi: the fused peer store exists in the tcgen05 kernel only (unset SGPR_SCOREMAT_FFMA)");
    if (ffma) {
        int rc = ensure(ctx->d_proj, ctx->proj_cap, static_cast<size_t>(R) * 512);
        if (rc) return rc;
        rc = ensure(ctx->d_blk, ctx->blk_cap, (static_cast<size_t>(R) + M) * kT);
        if (rc) return rc;
        float* rowblk = ctx->d_blk;
        float* colblk = ctx->d_blk + static_cast<size_t>(R) * kT;
        const int cap = ctx->sm_count * 8;
        float* const no_out = nullptr;
        SGPR_LAUNCH(sgpr_ntn_prep_kernel, R < cap ? R : cap, kThreads, 0, st, pooled_rows_dev, R, ctx->d_proj, rowblk, no_out, ctx->pw);
        SGPR_LAUNCH(sgpr_ntn_prep_kernel, M < cap ? M : cap, kThreads, 0, st, pooled_cols_dev, M, no_out, no_out, colblk, ctx->pw);
        const dim3 grid((M + kSmTJ - 1) / kSmTJ, (R + kSmTI - 1) / kSmTI);
        const size_t smem = (kSmTI * 512 + kSmTI * kT) * sizeof(float);
        SGPR_LAUNCH(sgpr_score_matrix_kernel, grid, kThreads, smem, st, ctx->d_proj, rowblk, pooled_cols_dev, colblk, R, M, scores_dev,
                    static_cast<long long>(ld_scores), ctx->pw.ntn_b, ctx->hp);
    }
#ifndef SGPR_EMU
    else {
        // tcgen05 path: split/swizzled operand planes (2 launches), then the persistent UMMA kernel
        int rc = ensure(ctx->d_proj, ctx->proj_cap, score_matrix_umma_scratch_floats(R, M));
        if (rc) return rc;
        score_matrix_umma_launch(ctx->sm_count, st, pooled_rows_dev, pooled_cols_dev, ctx->d_proj, scores_dev,
                                 static_cast<long long>(ld_scores), R, M, ctx->pw, ctx->hp, ctx->d_fc1_planes, n_out > 0 ? 1 : ctx->scoremat_version, outs, n_out);
    }
#endif
    ctx->launches += ffma ? 3 : 2;
    CUDA_TRY(cudaGetLastError());
    return SGPR_OK;
}

int sgpr_score_matrix(sgpr_ctx* ctx, const float* pooled_rows_dev, int R, const float* pooled_cols_dev, int M,
                      float* scores_dev, int64_t ld_scores, void* stream) {
    return score_matrix_impl(ctx, pooled_rows_dev, R, pooled_cols_dev, M, scores_dev, ld_scores, stream, nullptr, 0);
}

int sgpr_score_matrix_multi(sgpr_ctx* ctx, const float* pooled_rows_dev, int R, const float* pooled_cols_dev, int M,
                            float* const* scores_dev_list, int n_out, int64_t ld_scores, void* stream) {
    if (n_out < 1 || n_out > 8 || !scores_dev_list) return fail(SGPR_E_INVALID, "sgpr_score_matrix_multi: need 1..8 destination matrices");
    for (int p = 0; p < n_out; ++p)
        if (!scores_dev_list[p]) return fail(SGPR_E_INVALID, "sgpr_score_matrix_multi: destination %d is NULL", p);
    return score_matrix_impl(ctx, pooled_rows_dev, R, pooled_cols_dev, M, scores_dev_list[0], ld_scores, stream, scores_dev_list, n_out);
}

// ---- peer-visible result buffers (CUDA IPC): cudaMalloc'ed here so that the handle names a whole allocation ----------
int sgpr_peer_alloc(sgpr_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char* handle64) {
    if (!ctx || !dev_ptr || !handle64 || bytes == 0) return fail(SGPR_E_INVALID, "sgpr_peer_alloc: bad argument");
#ifdef SGPR_EMU
    return fail(SGPR_E_INVALID, "sgpr_peer_alloc: no peers in the emulator build");
#else
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    DeviceGuard guard(ctx->device);
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(SGPR_E_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    std::memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return SGPR_OK;
#endif
}

int sgpr_peer_open(sgpr_ctx* ctx, const unsigned char* handle64, void** dev_ptr) {
    if (!ctx || !dev_ptr || !handle64) return fail(SGPR_E_INVALID, "sgpr_peer_open: bad argument");
#ifdef SGPR_EMU
    return fail(SGPR_E_INVALID, "sgpr_peer_open: no peers in the emulator build");
#else
    DeviceGuard guard(ctx->device);        // opened with THIS context's device current: the mapping is for its kernels
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SGPR_OK;
#endif
}

int sgpr_peer_close(sgpr_ctx* ctx, void* dev_ptr) {
    if (!ctx || !dev_ptr) return SGPR_OK;
#ifndef SGPR_EMU
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
#endif
    return SGPR_OK;
}

int sgpr_peer_free(sgpr_ctx* ctx, void* dev_ptr) {
    if (!ctx || !dev_ptr) return SGPR_OK;
#ifndef SGPR_EMU
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaFree(dev_ptr));
#endif
    return SGPR_OK;
}

int sgpr_enable_peer_access(sgpr_ctx* ctx, int peer_device) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_enable_peer_access: ctx is NULL");
    if (peer_device == ctx->device) return SGPR_OK;
#ifdef SGPR_EMU
    return fail(SGPR_E_INVALID, "sgpr_enable_peer_access: no peers in the emulator build");
#else
    DeviceGuard guard(ctx->device);
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return fail(SGPR_E_CUDA, "sgpr_enable_peer_access: device %d cannot access device %d", ctx->device, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    if (e != cudaSuccess) return fail(SGPR_E_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", peer_device, cudaGetErrorString(e));
    return SGPR_OK;
#endif
}

int64_t sgpr_launch_count(const sgpr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sgpr_set_knn_ties(sgpr_ctx* ctx, int mode) {
    if (!ctx) return fail(SGPR_E_INVALID, "sgpr_set_knn_ties: ctx is NULL");
    if (mode != SGPR_TIES_CUDA && mode != SGPR_TIES_CPU)
        return fail(SGPR_E_INVALID, "sgpr_set_knn_ties: mode %d is neither SGPR_TIES_CUDA (0) nor SGPR_TIES_CPU (1)", mode);
    ctx->knn_ties = mode;
    return SGPR_OK;
}

int sgpr_get_knn_ties(const sgpr_ctx* ctx) { return ctx ? ctx->knn_ties : SGPR_E_INVALID; }

int sgpr_topk_cpu_rule_host(const float* rows, int num_rows, int n, int k, int depth_limit, int32_t* idx_out) {
    if (!rows || !idx_out) return fail(SGPR_E_INVALID, "sgpr_topk_cpu_rule_host: NULL argument");
    if (num_rows < 0 || n < 1 || n > SGPR_MAX_NODES || k < 1 || k > n)
        return fail(SGPR_E_INVALID, "sgpr_topk_cpu_rule_host: need 1 <= k <= n <= %d", SGPR_MAX_NODES);
    float v[SGPR_MAX_NODES];
    uint8_t ix[SGPR_MAX_NODES];
    for (int r = 0; r < num_rows; ++r) {
        for (int j = 0; j < n; ++j) { v[j] = rows[static_cast<size_t>(r) * n + j]; ix[j] = static_cast<uint8_t>(j); }
        if (depth_limit < 0) nth::topk_cpu_rule<uint8_t>(v, ix, n, k);
        else nth::introselect(nth::Seq<uint8_t>{v, ix}, 0, k - 1, n, depth_limit);
        for (int j = 0; j < k; ++j) idx_out[static_cast<size_t>(r) * k + j] = ix[j];
    }
    return SGPR_OK;
}

#ifdef SGPR_TIMELINE
// debug builds only (not declared in the public header): clock stamps of the N <= 64 fused kernel (embed_inst.cu)
int sgpr_debug_timeline(long long* out) { return sgpr::debug_read_timeline(out); }
int sgpr_debug_ctas(int* smid1024, long long* t2048) { return sgpr::debug_read_ctas(smid1024, t2048, nullptr); }
int sgpr_debug_cta_graphs(int* g1024) { return sgpr::debug_read_ctas(nullptr, nullptr, g1024); }
#endif

size_t sgpr_packed_size(void) { return make_offsets().total; }

int sgpr_pack_weights_host(const sgpr_weights* w, float* blob, float* head289, size_t* offsets17) {
    if (!w || !blob || !head289 || !offsets17) return fail(SGPR_E_INVALID, "sgpr_pack_weights_host: NULL argument");
    if (w->filters[0] != SGPR_FILTERS_1 || w->filters[1] != SGPR_FILTERS_2 || w->filters[2] != SGPR_FILTERS_3 ||
        w->tensor_neurons != SGPR_TENSOR_NEURONS || w->bottleneck != SGPR_BOTTLENECK)
        return fail(SGPR_E_ARCH, "sgpr_pack_weights_host: unsupported layer sizes");
    const PackOffsets o = make_offsets();
    std::vector<float> tmp;
    HeadParams hp;
    pack_weights(*w, tmp, hp, o);
    std::memcpy(blob, tmp.data(), tmp.size() * sizeof(float));
    std::memcpy(head289, hp.fc1_w, 256 * sizeof(float));
    std::memcpy(head289 + 256, hp.fc1_b, 16 * sizeof(float));
    std::memcpy(head289 + 272, hp.fc2_w, 16 * sizeof(float));
    head289[288] = hp.fc2_b;
    const size_t offs[17] = {o.s1, o.w_s2, o.w_s3, o.w_f1, o.w_f2, o.w_f3, o.w_end, o.ab_s2, o.ab_s3, o.ab_f1,
                             o.ab_f2, o.ab_f3, o.ab_end, o.att_w, o.ntn_w, o.ntn_v, o.ntn_b};
    std::memcpy(offsets17, offs, sizeof(offs));
    return SGPR_OK;
}

}  // extern "C"
