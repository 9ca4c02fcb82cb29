/* prego_b200 -- C ABI of the B200-native MiniROAD online step-recognition path.
 *
 * The reference (aleflabo/PREGO) is pure Python and has no FFI of its own; every entry
 * point below names the reference Python interface it stands in for.  All pointers are
 * raw CUDA device pointers unless marked "host"; torch (or any other caller) owns every
 * buffer, the library only owns the packed copy of the weights inside prego_model_t.
 * All functions return 0 (PREGO_OK) or a PREGO_ERR_* code; the message of the last
 * failing call on the calling thread is available from prego_last_error().  Kernels are
 * enqueued on the given stream (a cudaStream_t passed as void*); nothing synchronises.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef PREGO_B200_H
#define PREGO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PREGO_ABI_VERSION 4

#define PREGO_OK 0
#define PREGO_ERR_INVALID 1     /* bad argument / unsupported shape */
#define PREGO_ERR_CUDA 2        /* a CUDA runtime / driver call failed */
#define PREGO_ERR_WORKSPACE 3   /* workspace too small */
#define PREGO_ERR_STATE 4       /* weights not loaded, or not packed in the operand format this call reads */

/* Compute precision of the projection / recurrence / head GEMMs. */
#define PREGO_PREC_BF16 0       /* tcgen05 kind::f16, bf16 operands, fp32 accumulate (throughput path) */
#define PREGO_PREC_FP32 1       /* exact fp32 FFMA path (1e-4 parity mode) */
#define PREGO_PREC_TF32 3       /* training only: fp32 storage, tcgen05 kind::tf32 operands, fp32 accumulate */
#define PREGO_PREC_TF32X3 5     /* training only: fp32-class products on the tensor cores -- every operand split as a = hi + lo (TF32 each),
                                 * hi hi + lo hi + hi lo contracted in one kind::tf32 product over 3K (error ~2^-21 per term); the
                                 * persistent recurrence kernels (B <= 64) are exact fp32 in every mode */
#define PREGO_PREC_F16X3 4      /* fp32-class accuracy ON the tensor cores: every fp32 operand travels as fp16 hi + fp16 lo (22 bits),
                                   x.w = x_hi w_hi + x_lo w_hi + x_hi w_lo in one tcgen05 kind::f16 GEMM over 3 K, fp32 accumulate;
                                   LayerNorm, gates, state, softmax in fp32.  Same 1e-4 logit bound as PREGO_PREC_FP32 at a
                                   fraction of its cost; fp32 features only */
#define PREGO_PREC_F16 2        /* tcgen05 kind::f16, fp16 operands (10-bit mantissa = TF32 accuracy at the bf16 rate),
                                   fp32 accumulate; inputs saturate at +-65504 (default throughput path) */

/* Storage format of the rgb / flow feature tensors handed to prego_forward (feature ingest).
 * The reference loader produces fp32 (datasets/dataset.py:129-131).  PREGO_FEAT_16 declares that the caller already
 * keeps the features in the 16-bit operand format of `precision` (fp16 for PREGO_PREC_F16, bf16 for PREGO_PREC_BF16):
 * results are bit-identical to passing the same values as fp32 (the fp32 path rounds to that format first), the
 * staging pass disappears (the projection GEMM's TMA reads the caller's tensors in place) and host->device traffic
 * halves. */
#define PREGO_FEAT_F32 0
#define PREGO_FEAT_16 1

typedef struct prego_model prego_model_t;

/* Shapes read from the reference config dict by MROAD.__init__
 * (step_recognition/model/rnn/rnn.py:21-47): d_rgb / d_flow are FEATURE_SIZES[...] or 0
 * when cfg['no_rgb'] / cfg['no_flow'] is set. */
typedef struct prego_dims {
    int32_t d_rgb;
    int32_t d_flow;
    int32_t embed_dim;    /* cfg['embedding_dim'] = 2048 */
    int32_t hidden_dim;   /* cfg['hidden_dim']    = 1024 */
    int32_t num_classes;  /* cfg['num_classes']   = 86 | 12 */
} prego_dims_t;

/* The reference state_dict (rnn.py:38-47), fp32, torch layouts, device pointers. */
typedef struct prego_weights {
    const float* layer1_0_weight;          /* [E, d_rgb + d_flow] */
    const float* layer1_0_bias;            /* [E] */
    const float* layer1_1_weight;          /* [E]  LayerNorm gamma */
    const float* layer1_1_bias;            /* [E]  LayerNorm beta  */
    const float* gru_weight_ih_l0;         /* [3H, E]  rows r | z | n */
    const float* gru_weight_hh_l0;         /* [3H, H] */
    const float* gru_bias_ih_l0;           /* [3H] */
    const float* gru_bias_hh_l0;           /* [3H] */
    const float* f_classification_0_weight;/* [K, H] */
    const float* f_classification_0_bias;  /* [K] */
} prego_weights_t;

typedef struct prego_forward_args {
    const void* rgb;         /* [B, T, d_rgb]  contiguous, fp32 or 16-bit per feature_dtype (ignored when d_rgb == 0)  */
    const void* flow;        /* [B, T, d_flow] contiguous, same type (ignored when d_flow == 0 or flow_is_zero)        */
    int64_t B;
    int64_t T;
    float* h_state;          /* [B, H] GRU state, read before / written after; NULL = zeros in, dropped out
                                (rnn.py:49,60 always starts from h0 = 0) */
    float* probs;            /* [B, T, K] softmax probabilities (eval-mode out['logits'], rnn.py:69) or NULL */
    float* logits;           /* [B, T, K] raw logits (train-mode out['logits'], rnn.py:67) or NULL */
    int32_t* labels;         /* [B, T] argmax labels (trainer/eval.py:53) or NULL */
    void* workspace;         /* >= prego_workspace_bytes(model, B, chunk_T, precision) bytes, 1024-aligned */
    size_t workspace_bytes;
    int32_t precision;       /* PREGO_PREC_* */
    int32_t chunk_T;         /* frames per stream processed per pass (time chunk with carried h); 0 = T */
    int32_t feature_dtype;   /* PREGO_FEAT_* */
    int32_t flow_is_zero;    /* 1: the caller asserts flow == 0 everywhere, as both shipped configs feed it
                                (datasets/dataset.py:63-69); flow may be NULL and its half of the projection is skipped.
                                Results are bit-identical to passing an all-zero flow tensor. */
} prego_forward_args_t;

int prego_abi_version(void);
const char* prego_last_error(void);

/* Replaces MROAD.__init__ (rnn.py:21-49) + build_model(...).to(device) (model_builder.py:7-9). */
int prego_model_create(const prego_dims_t* dims, int32_t device, prego_model_t** out);
int prego_model_destroy(prego_model_t* model);

/* Replaces model.load_state_dict(torch.load(path)) (main.py:48): packs the ten fp32 tensors into the
 * library-owned operand formats (fp16 / bf16 copies, split-fp16 copies with a per-matrix power-of-two scale, gate-interleaved GRU
 * rows, padded head).  Synchronises `stream` once (the scale search reads three maxima back): not capturable into a CUDA graph. */
int prego_model_load_weights(prego_model_t* model, const prego_weights_t* w, void* stream);
/* The same for a subset of the operand formats (a training loop re-packs after every optimizer step and reads only the
 * fp32 set: PREGO_PACK_F32 costs no synchronisation and ~0.1 ms instead of ~0.5 ms).  PREGO_PACK_F32 is always included.
 * Formats not named become stale: prego_forward / prego_online_open in a precision whose format is stale return
 * PREGO_ERR_STATE until it is packed again.  prego_model_load_weights = PREGO_PACK_ALL. */
#define PREGO_PACK_F32 1u /* fp32 copies, gate-interleaved GRU rows: PREGO_PREC_FP32 inference and the training entry points */
#define PREGO_PACK_16 2u  /* fp16 and bf16 operand copies (PREGO_PREC_F16 / BF16, online sessions, anticipation in 16 bits) */
#define PREGO_PACK_X3 4u  /* split-fp16 copies (PREGO_PREC_F16X3); the scale search synchronises `stream` once */
#define PREGO_PACK_ALL 7u
int prego_model_load_weights_ex(prego_model_t* model, const prego_weights_t* w, uint32_t formats, void* stream);

size_t prego_workspace_bytes(const prego_model_t* model, int64_t B, int64_t chunk_T, int32_t precision);

/* Replaces MROAD.forward (rnn.py:51-71) plus the label extraction of Evaluate.eval (trainer/eval.py:53). */
int prego_forward(prego_model_t* model, const prego_forward_args_t* args, void* stream);

/* ---- Per-frame online sessions (BASELINE configs[1]: one frame per call, carried GRU state; rnn.py:51-71 called
 * with T = 1) for 1..8 concurrent streams.  The session owns a CUDA graph of the per-frame kernels; a step costs
 * one graph launch.  h_state [B, H] (in/out), probs [B, K] / logits [B, K] / labels [B] (each may be NULL) are
 * caller-owned device buffers fixed for the lifetime of the session; rgb / flow point at the frame's B x d_rgb /
 * B x d_flow fp32 features (device) and may change every step.  Weights are captured at open time: re-open after
 * prego_model_load_weights. */
typedef struct prego_online prego_online_t;
int prego_online_open(prego_model_t* model, int32_t num_streams, int32_t precision, float* h_state, float* probs,
                      float* logits, int32_t* labels, prego_online_t** out);
int prego_online_step(prego_online_t* session, const float* rgb, const float* flow, void* stream);
/* Host-side completion of the last prego_online_step without a stream synchronize: spins on pinned doorbell words the
 * frame's kernel writes, one 8-byte store {frame number, label} per stream (label and flag in the same store: no
 * system-scope fence on the device), and copies the labels into the session's label buffer.  Meant for sessions whose
 * label buffer lives in pinned host memory: step + wait is the whole per-frame round trip (no D2H copy).  Only sessions
 * on the one-kernel-per-frame path opened with `labels` in pinned host memory support it (PREGO_ERR_STATE otherwise):
 * device-resident sessions do not pay the per-frame PCIe store.  probs / logits in host memory are fenced first. */
int prego_online_wait(prego_online_t* session);
/* prego_online_step + prego_online_wait in one call (one foreign-function round trip per frame for bindings where that costs ~1 us). */
int prego_online_step_wait(prego_online_t* session, const float* rgb, const float* flow, void* stream);
int prego_online_close(prego_online_t* session);
/* Diagnostics: SM-clock stamps of the phase boundaries of the last frame, out[ctas][16] (host memory), for sessions
 * opened with the environment variable PREGO_ONLINE_TRACE=1; *num_ctas = rows written (<= max_ctas). */
int prego_online_trace(prego_online_t* session, int64_t* out, int32_t max_ctas, int32_t* num_ctas);

/* Device-side watchdog: the persistent recurrence kernels bound every inter-CTA spin; if a peer never shows up
 * they set a flag instead of hanging the GPU.  Reads (and clears) it; synchronises the device.  0 = healthy. */
int prego_device_error(prego_model_t* model, int32_t* out);
/* How many times the batched recurrence had to be launched WITHOUT the cooperative attribute (driver refused it for
 * the cluster launch; taken only when the occupancy calculator confirms every CTA pair is co-resident, otherwise
 * prego_forward fails).  0 on a healthy B200 stack; no reference counterpart. */
int prego_recurrence_fallbacks(prego_model_t* model, int64_t* out);

/* ---- Training step (reference: rnn.py:51-71 in train mode, trainer/train.py:20-24) -------------------------
 * prego_train_forward computes the raw logits [B, T, K] (train-mode out['logits'], rnn.py:67) with dropout active
 * (rnn.py:43; own counter-based mask from `seed`) and keeps the activations in the workspace;
 * prego_train_backward takes dL/dlogits and writes the gradients of the ten state_dict tensors (overwriting
 * `grads`, torch layouts): full BPTT through all T steps.  What loss.backward() does for the reference module.
 * Exact fp32.  The criterion (criterions/loss.py:15-34), the optimizer step and the gradient all-reduce stay
 * with the caller (torch / torch.distributed NCCL). */
typedef struct prego_grads {
    float* layer1_0_weight;
    float* layer1_0_bias;
    float* layer1_1_weight;
    float* layer1_1_bias;
    float* gru_weight_ih_l0;
    float* gru_weight_hh_l0;
    float* gru_bias_ih_l0;
    float* gru_bias_hh_l0;
    float* f_classification_0_weight;
    float* f_classification_0_bias;
} prego_grads_t;

typedef struct prego_train_args {
    const float* rgb;        /* [B, T, d_rgb]  */
    const float* flow;       /* [B, T, d_flow] */
    int64_t B;
    int64_t T;
    float* logits;           /* forward:  [B, T, K] out */
    const float* dlogits;    /* backward: [B, T, K] in  */
    const prego_grads_t* grads; /* backward: ten gradient buffers, overwritten */
    void* workspace;         /* >= prego_train_workspace_bytes(model, B, T); the same buffer for forward and backward */
    size_t workspace_bytes;
    float dropout_p;         /* cfg['dropout']; 0 disables */
    uint64_t seed;           /* dropout mask stream */
    int32_t precision;       /* PREGO_PREC_FP32 (0 is read as FP32 too): every GEMM exact fp32 on CUDA cores (the parity
                                mode); PREGO_PREC_TF32: the large projections and their weight / input gradients run on
                                tcgen05 kind::tf32 (fp32 storage, 10-bit-mantissa operands, fp32 accumulate) -- the
                                reference's own GPU practice (cuDNN RNN allows TF32; main.py --amp trains in fp16).
                                The recurrence, LayerNorm, loss and the classifier stay exact fp32 either way. */
    int32_t reserved;
    void* gru_grads_event;   /* prego_train_backward only, may be NULL: a cudaEvent_t recorded on `stream` as soon as the gradients of
                                the gru.* and f_classification.* tensors are final, i.e. before the layer1 backward (LayerNorm backward,
                                dW1 = the largest GEMM of the step) is enqueued.  A data-parallel caller starts the all-reduce of that
                                bucket on this event so it overlaps the rest of the backward. */
} prego_train_args_t;

size_t prego_train_workspace_bytes(const prego_model_t* model, int64_t B, int64_t T);
/* The same for a given training precision (PREGO_PREC_TF32X3 adds the split-operand scratch). */
size_t prego_train_workspace_bytes_ex(const prego_model_t* model, int64_t B, int64_t T, int32_t precision);
int prego_train_forward(prego_model_t* model, const prego_train_args_t* args, void* stream);
int prego_train_backward(prego_model_t* model, const prego_train_args_t* args, void* stream);

/* Phase timing of prego_forward, measured with CUDA events on the launching stream (what bench.py's
 * roofline uses).  prego_profile_begin arms it; prego_profile_end waits for the last recorded event and
 * returns the summed device time (ms) and kernel-launch count of each phase since begin. */
#define PREGO_PHASE_STAGE 0       /* feature concat + operand rounding */
#define PREGO_PHASE_GEMM1 1       /* Linear D_in -> E (rnn.py:40) */
#define PREGO_PHASE_LAYERNORM 2   /* LayerNorm + ReLU (rnn.py:41-42) */
#define PREGO_PHASE_GEMM2 3       /* GRU input gates (rnn.py:61) */
#define PREGO_PHASE_RECURRENCE 4  /* GRU time steps (rnn.py:61) */
#define PREGO_PHASE_HEAD 5        /* classifier + softmax + argmax (rnn.py:62-69, eval.py:53) */
#define PREGO_NUM_PHASES 6
int prego_profile_begin(prego_model_t* model);
int prego_profile_end(prego_model_t* model, double* phase_ms /*[PREGO_NUM_PHASES]*/,
                      int64_t* phase_launches /*[PREGO_NUM_PHASES]*/);

/* Replaces the window vote of aggregate() (utils/aggregate.py:55,65-72) for a ragged batch of label
 * sequences.  labels: concatenated int32; offsets[B+1] (frames) and win_offsets[B+1] (windows) int64 device
 * arrays; modes[win_offsets[B]] out.  err_flag (device int, caller-zeroed) becomes 1 if a label is outside
 * [0, num_labels). */
int prego_window_mode(const int32_t* labels, const int64_t* offsets, const int64_t* win_offsets, int32_t B,
                      int64_t total_windows, int32_t window, int32_t num_labels, int32_t* modes, int32_t* err_flag,
                      void* stream);

/* Replaces find_changes + eliminate_consecutive_duplicates (utils/aggregate.py:7-43) for a ragged batch.
 * Sequence b = seq[seg_offsets[b] .. seg_offsets[b+1]).  out_vals / out_changes use the same offsets (capacity =
 * sequence length); change indices are element index * scale, the last one is final_len[b].
 * counts[b] = number of runs, -1 for an empty sequence (the reference raises IndexError). */
int prego_rle(const int32_t* seq, const int64_t* seg_offsets, const int64_t* final_len, int32_t B, int64_t scale,
              int32_t* out_vals, int64_t* out_changes, int32_t* counts, void* stream);

/* ---- MiniROADA anticipation head (reference: MROADA, rnn.py:73-137, registered "MiniROADA").
 * Same trunk as MROAD (rnn.py:117-124), plus per frame
 *     ant_h[a]      = anticipation_layer.0(relu(h_t))[a*H:(a+1)*H]            (rnn.py:108-110,125)
 *     ant_logits[a] = f_classification(relu(ant_h[a])),  a = 0..A-1             (rnn.py:126, the SAME classifier)
 * and, in eval mode, softmax over the classes (rnn.py:133).  prego_model_load_anticipation packs the two extra
 * state_dict tensors (A = cfg['anticipation_length']); prego_forward_anticipation is prego_forward plus the
 * anticipation outputs.  The [rows, A*H] activation is produced in row slabs sized by `workspace` (never the whole
 * [B*T, A*H] tensor): any size >= prego_anticipation_workspace_bytes(model, 256, precision) works, larger slabs are
 * faster.  The one-kernel-per-frame online path is not used by this entry point.  Inference only. */
typedef struct prego_anticipation_args {
    float* probs;            /* [B, T, A, K] softmax (eval-mode out['anticipation_logits'], rnn.py:133-135) or NULL */
    float* logits;           /* [B, T, A, K] raw (train-mode out['anticipation_logits'], rnn.py:130) or NULL */
    int32_t* labels;         /* [B, T, A] first-max argmax or NULL */
    void* workspace;         /* slab buffer, 1024-aligned */
    size_t workspace_bytes;
} prego_anticipation_args_t;

int prego_model_load_anticipation(prego_model_t* model, int32_t anticipation_length,
                                  const float* anticipation_layer_0_weight /* [A*H, H] */,
                                  const float* anticipation_layer_0_bias /* [A*H] */, void* stream);
size_t prego_anticipation_workspace_bytes(const prego_model_t* model, int64_t slab_rows, int32_t precision);
int prego_forward_anticipation(prego_model_t* model, const prego_forward_args_t* args,
                               const prego_anticipation_args_t* ant, void* stream);

/* ---- Per-frame average precision on the device (reference: perframe_average_precision,
 * utils/metrics.py:25-62 with metrics == 'AP', i.e. sklearn.metrics.average_precision_score per class).
 * scores: [N, K] fp32 probabilities in [0, 1] (the concatenated eval-mode outputs, trainer/eval.py:46-49);
 * ground truth either `targets` [N, K] fp32 (non-zero = positive; the reference's one-hot target rows) or, when
 * targets == NULL, `target_labels` [N] int32 (one-hot implied).  For every class k:
 *     order frames by score descending; at each DISTINCT score threshold n: P_n = tp_n / (tp_n + fp_n),
 *     R_n = tp_n / positives;  ap[k] = sum_n (R_n - R_{n-1}) * P_n                      (float64)
 * num_pos[k] = positives of class k (ap[k] is NaN when 0: the reference skips such classes, metrics.py:54).
 * Class 0 (background) is computed too; the caller drops it (metrics.py:47,53).  Every class is cut into slices so
 * that the whole machine works on it: four stable 8-bit LSD radix passes over the 31-bit keys (score bits << 1 |
 * positive) in `workspace`, then a sliced scan over the sorted keys (16 launches in all, on `stream`).
 * err_flag (device int, caller-zeroed) becomes 1 if a score is outside [0, 1] or NaN. */
size_t prego_ap_workspace_bytes(int64_t N, int32_t K);
int prego_perframe_ap(const float* scores, const float* targets, const int32_t* target_labels, int64_t N, int32_t K,
                      double* ap /*[K]*/, int64_t* num_pos /*[K]*/, void* workspace, size_t workspace_bytes,
                      int32_t* err_flag, void* stream);

/* ---- Host half of the feature ingest (reference: datasets/dataset.py:120-132 hands the model fp32
 * HOST tensors).  Rounds n fp32 values (host memory) to the 16-bit operand format of `precision` (PREGO_PREC_F16 /
 * PREGO_PREC_BF16) with exactly the rounding the device applies in its staging pass (fp16: clamp to +-65504, NaN ->
 * -65504, round to nearest even; bf16: round to nearest even), on `num_threads` host threads (AVX-512 / AVX2+F16C /
 * scalar, picked at run time; prego_host_round_impl() = 2 / 1 / 0).  The result is what PREGO_FEAT_16 expects: the
 * host->device link carries half the bytes and prego_forward returns bit-identical results.  Data-format staging only:
 * nothing of the model is computed on the CPU.  Blocking; thread-safe. */
int prego_host_round_features(const float* src, void* dst, int64_t n, int32_t precision, int32_t num_threads);
int prego_host_round_impl(void);
/* 1 when all n values are +-0.0 -- the reference loader's all-zero flow dummy (datasets/dataset.py:63-69), which the
 * caller may then declare with flow_is_zero instead of copying and multiplying it -- 0 otherwise, -1 on a bad argument. */
int prego_host_all_zero(const float* src, int64_t n, int32_t num_threads);
/* Ring stager: rounds a large fp32 HOST tensor (same rule) and copies it to `dst_device` (16-bit, device) on `stream`
 * through a small pinned ring (`ring_slots` x `slot_bytes`, a few MiB each: the copy engine loses ~30 % on sub-MiB
 * copies): the rounded values are still in the CPU's last-level cache when the DMA engine reads them, so the host's
 * DRAM -- the end-to-end bottleneck -- only sees the fp32 read.  A persistent pool of `num_threads` threads (the caller
 * included) claims 128 KiB chunks from one atomic counter, no barrier (spins during a run, sleeps between runs).  The
 * stager belongs to the device that is current at create time.  prego_host_stager_run returns when the last copy has
 * been ENQUEUED on `stream` (the rounding itself is done).  One run at a time per stager. */
typedef struct prego_host_stager prego_host_stager_t;
int prego_host_stager_create(int32_t num_threads, int32_t ring_slots, int64_t slot_bytes, prego_host_stager_t** out);
int prego_host_stager_destroy(prego_host_stager_t* stager);
int prego_host_stager_run(prego_host_stager_t* stager, const float* src, void* dst_device, int64_t n, int32_t precision,
                          void* stream);

/* Building blocks exposed for parity tests and micro-benchmarks. */
/* C[M,N] (fp32, ldc = N) = A[M,K] * W[N,K]^T + bias[N] with 16-bit operands (precision = PREGO_PREC_F16 or
 * PREGO_PREC_BF16) on the tcgen05 path; N % tile_n == 0, tile_n in {96, 128, 192, 256}, K % 64 == 0. */
int prego_gemm16_nt(const void* A, const void* W, const float* bias, float* C, int64_t M, int64_t N, int64_t K,
                    int32_t tile_n, int32_t precision, void* stream);
/* Fused LayerNorm path (rnn.py:39-44 without a separate LayerNorm pass), building blocks:
 * prego_gemm16_stats_nt: Y[M,N] fp16 = A W^T + bias (16-bit operands, CTA pairs) and, from the same epilogue, the LayerNorm
 *   statistics of the rounded rows: stats [N/256][M] (sum, sum of squares) partials and rowstat [M] = (rstd, -mean * rstd)
 *   (biased variance, eps inside the root).  N % 256 == 0, K % 64 == 0.
 * prego_gemm16_ln_nt: C[M,N] fp32 = relu((Y * a_r + b_r) * gamma + beta) W^T + bias with Y fp16 [M,K] normalised on its way
 *   into the tensor core (transform warps rewrite the TMA-staged A tile in place); rowstat = (a_r, b_r) per row.
 *   rowstat == NULL: identity transform (Y in the operand format passes through the extra pipeline hop unchanged). */
int prego_gemm16_stats_nt(const void* A, const void* W, const float* bias, void* Y, float* stats, float* rowstat, int64_t M,
                          int64_t N, int64_t K, int32_t precision, float eps, void* stream);
int prego_gemm16_ln_nt(const void* Y, const float* rowstat, const float* gamma, const float* beta, const void* W,
                       const float* bias, float* C, int64_t M, int64_t N, int64_t K, int32_t precision, void* stream);
/* Same contract with fp32 storage and TF32 tensor-core operands (CTA pairs); N % 256 == 0, K % 32 == 0;
 * accumulate bit 0 adds into C, bit 8 (0x100) selects 64-column tiles (many small tiles: the per-time-step products of the
 * training recurrence).  Used by the PREGO_PREC_TF32 training step. */
int prego_gemm_tf32_nt(const float* A, const float* W, const float* bias, float* C, int64_t M, int64_t N, int64_t K,
                       int32_t accumulate, void* stream);
/* Same contract in exact fp32 on CUDA cores (K % 16 == 0). */
int prego_gemm_f32_nt(const float* A, const float* W, const float* bias, float* C, int64_t M, int64_t N, int64_t K,
                      void* stream);

/* ---- Fused AdamW over a list of fp32 tensors (one launch for all of them): the optimizer step of main.py:62-67
 * (torch.optim.AdamW, amsgrad = False, maximize = False), same update order as torch's single-tensor rule:
 *   p *= 1 - lr * wd;  m = lerp(m, g, 1 - b1);  v = b2 * v + (1 - b2) * g * g;
 *   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
 * grad_scale multiplies every gradient first (1 / world_size after a sum all-reduce; 1 otherwise). */
#define PREGO_ADAMW_MAX_TENSORS 16
typedef struct prego_adamw_args {
    float* params[PREGO_ADAMW_MAX_TENSORS];
    const float* grads[PREGO_ADAMW_MAX_TENSORS];
    float* exp_avg[PREGO_ADAMW_MAX_TENSORS];
    float* exp_avg_sq[PREGO_ADAMW_MAX_TENSORS];
    int64_t numel[PREGO_ADAMW_MAX_TENSORS];
    int32_t num_tensors;
    int32_t step;            /* t >= 1, after the increment */
    float lr, beta1, beta2, eps, weight_decay, grad_scale;
} prego_adamw_args_t;
int prego_adamw_step(const prego_adamw_args_t* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PREGO_B200_H */
