/*
 * sgpr_b200_train.h — C ABI of the B200-native SG_PR TRAINING step (BASELINE config 3).
 *
 * The reference has no native layer; what this replaces is the device work of
 *     SGTrainer.process_batch(batch, training=True)                       /root/reference/sg_net.py:312-345
 * from the point where the batch tensors exist:
 *     optimizer.zero_grad(); prediction = model(data)   (SG.forward in train mode, sg_net.py:112-138: the seven
 *                                                        BatchNorm layers of sg_net.py:50-76 use batch statistics,
 *                                                        one BatchNorm batch per side, sg_net.py:123-124)
 *     losses = mean(binary_cross_entropy(prediction, target))             sg_net.py:335
 *     losses.backward(); optimizer.step()               (Adam, lr, weight_decay — sg_net.py:337-338, 351-352)
 *
 * State lives on the device as ONE flat fp32 vector: the trainable parameters in their state_dict shapes followed by
 * the BatchNorm running statistics (sgpr_train_layout lists name / offset / size of every tensor, names are the
 * checkpoint keys of sg_net.py:164-174 without the "module." prefix).  The int64 `num_batches_tracked` counters are
 * not part of it: each BatchNorm runs twice per step (once per side), the host adds 2 per step.
 *
 * Conventions are those of sgpr_b200.h: fp32, contiguous, device pointers, work enqueued on `stream` without
 * synchronising, 0 / negative SGPR_E_* return codes with sgpr_last_error() text.  No CPU implementation exists behind
 * this ABI.
 */
#ifndef SGPR_B200_TRAIN_H
#define SGPR_B200_TRAIN_H

#include <stddef.h>
#include <stdint.h>

#include "sgpr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgpr_train sgpr_train;   /* opaque: device state vector, Adam moments, workspace */

/* Flat layout: *count tensors; names[i], offsets[i] (floats), sizes[i] (floats).  The first *n_params tensors are
 * trainable (sgpr_train_param_count floats in total), the rest are running_mean / running_var buffers. */
int     sgpr_train_layout(int* count, int* n_params, const char* const** names, const int64_t** offsets,
                          const int64_t** sizes);
int64_t sgpr_train_param_count(void);   /* 47,985 */
int64_t sgpr_train_state_count(void);   /* 48,689 = parameters + 704 running statistics */

/* Replaces SG(...).cuda() + torch.optim.Adam(model.parameters(), lr, weight_decay)   sg_net.py:158-176, 351-352 */
int sgpr_train_create(sgpr_train** out, int device);
int sgpr_train_destroy(sgpr_train* t);

/* Upload / download the flat state (host pointers).  reset_optimizer != 0 zeroes Adam's moments and step count. */
int sgpr_train_set_state(sgpr_train* t, const float* state_host, int reset_optimizer);
int sgpr_train_get_state(sgpr_train* t, float* state_host);

/* torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8; the reference passes lr and weight_decay (config.yml). */
int sgpr_train_set_optimizer(sgpr_train* t, float lr, float weight_decay, float beta1, float beta2, float eps);

/*
 * ONE optimiser step on a batch of B ordered pairs (the reference feeds every listed pair in both orders, so its B is
 * twice the listed batch — sg_net.py:324-331):
 *   f1_dev, f2_dev : [B][15][N]   data["features_1"/"features_2"]
 *   target_dev     : [B]          data["target"] (0 / 1)
 *   loss_dev       : [1]          mean BCE of this batch (before the update), may be NULL
 *   pred_dev       : [B]          train-mode predictions (before the update), may be NULL
 *   flags          : SGPR_TRAIN_APPLY     full step (Adam update + running statistics); without it forward + backward
 *                                         only — gradients stay readable through sgpr_train_get_grads, nothing is updated
 *                    SGPR_TRAIN_MIRRORED  the caller guarantees features_2[p] == features_1[p ^ 1] for every p (exactly
 *                                         what process_batch builds: each listed pair (a, b) appears as (a, b) and
 *                                         (b, a), sg_net.py:324-331).  Both sides then hold the same graphs, so the
 *                                         second dgcnn_conv_pass (sg_net.py:124) has the same batch statistics and
 *                                         activations as the first: the kernels embed every graph ONCE and back-
 *                                         propagate the sum of its two roles' gradients.  f2_dev is not read (may be
 *                                         NULL); B must be even.  Same result as the unmirrored call up to rounding.
 * B >= 1, 2 <= N <= SGPR_MAX_NODES, 1 <= k <= N.
 */
#define SGPR_TRAIN_APPLY    1
#define SGPR_TRAIN_MIRRORED 2
int sgpr_train_step(sgpr_train* t, const float* f1_dev, const float* f2_dev, const float* target_dev, int B, int N, int k,
                    float* loss_dev, float* pred_dev, int flags, void* stream);

/*
 * The same step split at the loss, for callers that drive autograd themselves (model(data) in train mode, their own
 * loss, loss.backward(), their own optimiser — the general use of SG.forward, sg_net.py:112-138):
 *   sgpr_train_forward : train-mode forward only.  pred_dev [B]; att1_dev / att2_dev [B][N] attention scores or NULL
 *                        (att2 is not written for a mirrored batch: it is att1 with adjacent rows swapped).
 *                        With SGPR_TRAIN_APPLY the BatchNorm running statistics are updated, as nn.BatchNorm does in
 *                        forward; parameters are never touched.  Everything the backward needs stays in the context,
 *                        except the input blocks themselves: f1_dev / f2_dev must stay valid until sgpr_train_backward.
 *   sgpr_train_backward: backward of the LAST sgpr_train_forward from dpred_dev [B] = d loss / d prediction; the flat
 *                        gradient vector (sgpr_train_param_count floats) is left readable through sgpr_train_get_grads
 *                        and, if grads_dev != NULL, copied there (device pointer).
 *   sgpr_train_set_state_dev / _get_state_dev : the flat state vector from / to DEVICE memory (sgpr_train_state_count
 *                        floats), for modules whose parameters already live on the GPU.
 */
int sgpr_train_forward(sgpr_train* t, const float* f1_dev, const float* f2_dev, int B, int N, int k, float* pred_dev,
                       float* att1_dev, float* att2_dev, int flags, void* stream);
int sgpr_train_backward(sgpr_train* t, const float* dpred_dev, float* grads_dev, void* stream);
int sgpr_train_set_state_dev(sgpr_train* t, const float* state_dev, void* stream);
int sgpr_train_get_state_dev(sgpr_train* t, float* state_dev, void* stream);

/*
 * Training-batch assembly + augmentation on the device, for graphs already resident in HBM.  Replaces the host loop of
 * process_batch (sg_net.py:316-331) with transfer_to_torch's training branch (sg_net.py:286-295) and the augmentations
 * of utils.py:91-178 (reference defaults; applied to all node_num rows, zero pads included; one x-flip decision per
 * listed pair):
 *   graphs_dev   : [M][15][N]  un-augmented padded blocks, exactly what transfer_to_torch(training=False) builds
 *   pair_idx_dev : [P][2] int32  listed pairs (graph a, graph b), indices into graphs_dev
 *   out_f1_dev   : [2P][15][N]  row 2p = augmented a, row 2p+1 = augmented b  == features_1 of the mirrored batch
 *   draws_dev    : NULL or [2P][12]  the random draws of every slot (flip u, angle u, scale, 3 raw perturbation normals,
 *                  3 shifts) and jitter_dev : NULL or [2P][N][3] raw jitter normals — parity taps: the tests feed them
 *                  to the reference-shaped numpy functions and compare.
 * Random numbers are Philox4x32-10 keyed by `seed` with counter (step, slot, node): reproducible from (seed, step), but
 * not numpy's global stream (which the reference never seeds) — parity with the reference is distributional.
 */
int sgpr_train_assemble(sgpr_train* t, const float* graphs_dev, int M, int N, const int32_t* pair_idx_dev, int P,
                        uint64_t seed, uint64_t step, float* out_f1_dev, float* draws_dev, float* jitter_dev, void* stream);

/*
 * Every sgpr_train_forward overwrites the context's single activation workspace and advances this counter; a caller that
 * keeps several forwards alive (autograd graphs) records the value after its forward and must find it unchanged when it
 * calls sgpr_train_backward — otherwise the backward would differentiate a LATER forward's activations.
 */
int64_t sgpr_train_forward_generation(const sgpr_train* t);

/* k-NN tie rule of the train-mode forward: SGPR_TIES_CUDA (default) / SGPR_TIES_CPU, see sgpr_set_knn_ties (sgpr_b200.h). */
int sgpr_train_set_knn_ties(sgpr_train* t, int mode);
int sgpr_train_get_knn_ties(const sgpr_train* t);

/* Gradients of the last step w.r.t. the trainable parameters (before weight decay), flat layout, host pointer. */
int sgpr_train_get_grads(sgpr_train* t, float* grads_host);

/* Number of optimiser steps applied so far / kernel launches enqueued so far. */
int64_t sgpr_train_step_count(const sgpr_train* t);
int64_t sgpr_train_launch_count(const sgpr_train* t);

/*
 * Debug/parity tap: copy an internal tensor of the LAST step to the host.  what: "yext","a","d","sumy","gz","enode",
 * "idx" (layer = 0..5: xyz 1-3, sem 1-3), "yend","gzend","pooled","att","dpooled","stats","bsum" (layer ignored).
 * Returns the number of bytes copied (<= cap_bytes) or a negative error.
 */
int64_t sgpr_train_debug_read(sgpr_train* t, const char* what, int layer, void* host, int64_t cap_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SGPR_B200_TRAIN_H */
