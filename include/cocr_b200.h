/*
 * cocr_b200.h — C ABI of the B200-native conformer_ocr hot path.
 *
 * The reference (mittagessen/conformer_ocr) is pure Python/torch and has no FFI of its
 * own; its seam for this path is three Python call signatures.  Each entry
 * point below names the reference interface it replaces (paths relative to
 * /root/reference/conformer_ocr/).  No torch types cross this boundary: plain pointers,
 * sizes, a `cudaStream_t` passed as `void*`, `int` status codes.  The ctypes binding that a
 * maintainer of the reference would add is shown in INTEGRATION.md and lives in
 * conformer_ocr_b200/_lib.py.
 *
 * Conventions
 *  - every function returns COCR_OK (0) or a negative cocr_status; the message of the
 *    last failure is available from cocr_last_error().  No exception crosses the ABI.
 *  - "dev" pointers are device memory on the handle's device, "host" pointers are host
 *    memory.  The library never frees caller memory and the caller never frees library
 *    memory; packed weights and the activation workspace belong to the handle.
 *  - all work is enqueued on the caller's stream; the only hidden synchronisation is
 *    in the *_host entry points (which return host results) and when the handle has to
 *    grow its workspace or rebuild the positional table for a new frame count.
 *  - one handle per (process, device); calls on one handle are not re-entrant.
 */
#ifndef COCR_B200_H_
#define COCR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cocr_handle cocr_handle;

typedef enum cocr_status {
    COCR_OK = 0,
    COCR_ERR_BAD_ARG = -1,        /* -> ValueError in the Python shim                     */
    COCR_ERR_UNSUPPORTED = -2,    /* shape/hyper-parameter the kernels do not implement    */
    COCR_ERR_OOM = -3,            /* -> RuntimeError('CUDA out of memory. ...')             */
    COCR_ERR_CUDA = -4,           /* any other CUDA runtime/driver failure                 */
    COCR_ERR_STATE = -5           /* call order violated (weights missing/not finalised)   */
} cocr_status;

typedef enum cocr_precision {
    COCR_FP32 = 0,   /* CUDA-core fp32 everywhere; logits within 1e-4 (max-norm rel.) of the reference */
    COCR_BF16 = 1    /* tcgen05 bf16 operands, fp32 accumulate / residual stream / statistics         */
} cocr_precision;

typedef enum cocr_mask_mode {
    COCR_MASK_REFERENCE_UNMASKED = 0, /* what the reference computes: no mask (encoder.py:67-103)  */
    COCR_MASK_RAGGED = 1              /* keys beyond a line's own length are excluded             */
} cocr_mask_mode;

/* Hyper-parameters: the ctor arguments of ConformerEncoder (conformer/encoder.py:135-152)
 * as PytorchRecognitionModel passes them (pred.py:51-89), plus num_classes of the decoder
 * nn.Linear (pred.py:90). */
typedef struct cocr_hparams {
    int32_t height;                     /* input_dim                                   */
    int32_t encoder_dim;
    int32_t num_layers;
    int32_t num_heads;
    int32_t ff_expansion;               /* feed_forward_expansion_factor               */
    int32_t conv_expansion;             /* conv_expansion_factor (must be 2)           */
    int32_t conv_kernel_size;           /* odd                                          */
    int32_t half_step_residual;         /* bool                                         */
    int32_t subsampling_conv_channels;
    int32_t subsampling_factor;         /* only 4 is implemented                        */
    int32_t num_classes;                /* 0 = encoder only                             */
    int32_t precision;                  /* cocr_precision                               */
    int32_t mask_mode;                  /* cocr_mask_mode                               */
} cocr_hparams;

/* Library version / build information (static string). */
const char *cocr_version(void);

/* Message of the last failure on this handle (or of the last failed cocr_create when
 * `h` is NULL).  Never NULL. */
const char *cocr_last_error(const cocr_handle *h);

/* Replaces ConformerEncoder.__init__ + nn.Linear decoder construction (pred.py:75-92,
 * model.py:99-116).  Validates the hyper-parameters with the reference's own rules
 * (attention.py:53, convolution.py:132-133,180-181). */
int cocr_create(const cocr_hparams *hp, int device, cocr_handle **out);
int cocr_destroy(cocr_handle *h);

/* Replaces nn.Module.load_state_dict for `PytorchRecognitionModel.nn` (pred.py:194,209).
 * `key` is the state-dict name below the ModuleDict: "encoder...." / "decoder.weight" /
 * "decoder.bias" (SURVEY.md §8a "Weights").  `data` is host fp32, C-contiguous, `shape`
 * has `ndim` entries.  Unknown keys -> COCR_ERR_BAD_ARG; "num_batches_tracked" is ignored. */
int cocr_set_weight(cocr_handle *h, const char *key, const float *data_host, int ndim, const int64_t *shape);
/* Folds BatchNorm (eval), packs/casts the tensors for the kernels, uploads them. */
int cocr_finalize_weights(cocr_handle *h);

/* calc_length (conformer/convolution.py:240-247) for subsampling factor 4. */
int32_t cocr_out_frames(int32_t width);

/* ConformerEncoder.forward (conformer/encoder.py:183-204).
 *   img_dev     fp32 image batch addressed as img[b*stride_b + w*stride_w + y*stride_h]
 *               (element strides), b < B, w < W (time axis), y < height; covers both the
 *               (N,1,H,W) collate layout (pred.py:119 takes a transposed view of it) and a
 *               contiguous (B,W,H) tensor.
 *   widths_dev  int32 (B) true widths (`input_lengths`); may be NULL = all W.
 *   enc_out_dev fp32 (B, T, encoder_dim), T = cocr_out_frames(W).
 *   out_lens_dev int32 (B) = calc_length(widths); may be NULL.                        */
int cocr_encode(cocr_handle *h, const float *img_dev, int64_t stride_b, int64_t stride_w, int64_t stride_h,
                const int32_t *widths_dev, int32_t B, int32_t W, void *stream,
                float *enc_out_dev, int32_t *out_lens_dev);

/* PytorchRecognitionModel.forward / RecognitionModel.forward (pred.py:118-122,
 * model.py:125-127): encoder + decoder nn.Linear -> raw logits ("probits") fp32 (B,T,C). */
int cocr_logits(cocr_handle *h, const float *img_dev, int64_t stride_b, int64_t stride_w, int64_t stride_h,
                const int32_t *widths_dev, int32_t B, int32_t W, void *stream,
                float *logits_dev, int32_t *out_lens_dev);

/* kraken.lib.ctc_decoder.greedy_decoder applied per line to logits[b,:len_b,:]
 * (call sites pred.py:138-143,158-163,173-177; model.py:163-167): argmax over classes
 * (lowest index wins ties), merge runs, drop blank 0.
 *   logits_dev fp32 (B,T,C); lens_dev int32 (B) or NULL (= T)
 *   labels/starts/ends int32 (B,T), conf fp32 (B,T): first counts[b] entries per line are
 *   valid; counts int32 (B).  ends are inclusive frame indices; conf = max winning logit. */
int cocr_greedy_decode(cocr_handle *h, const float *logits_dev, const int32_t *lens_dev,
                       int32_t B, int32_t T, int32_t C, void *stream,
                       int32_t *labels_dev, int32_t *starts_dev, int32_t *ends_dev, float *conf_dev,
                       int32_t *counts_dev);

/* predict_labels (pred.py:165-178) without the (B,C,T) logits round trip through the
 * host: encoder -> projection -> argmax -> collapse, all on the device. */
int cocr_forward_decode(cocr_handle *h, const float *img_dev, int64_t stride_b, int64_t stride_w, int64_t stride_h,
                        const int32_t *widths_dev, int32_t B, int32_t W, void *stream,
                        int32_t *labels_dev, int32_t *starts_dev, int32_t *ends_dev, float *conf_dev,
                        int32_t *counts_dev, int32_t *out_lens_dev);

/* The same call a `cocr test` batch makes (cli/test.py:185-191), with HOST buffers:
 * image batch (B,1,H,W) fp32 contiguous + int32 widths on the host (pinned or pageable),
 * H2D copy, forward, decode, D2H of the packed result.  Synchronises `stream` before
 * returning.  labels/starts/ends/conf are (B,T) host arrays, counts/out_lens (B). */
int cocr_recognize_host(cocr_handle *h, const float *img_host, const int32_t *widths_host,
                        int32_t B, int32_t W, void *stream,
                        int32_t *labels_host, int32_t *starts_host, int32_t *ends_host, float *conf_host,
                        int32_t *counts_host, int32_t *out_lens_host);

/* The same, split in two so that a caller can keep two batches in flight (slot 0 / 1): `begin` enqueues the H2D copy
 * of the batch on the handle's copy stream, the forward pass + decode on `stream`, the D2H of the records on a third stream
 * of the handle (so the other slot's forward pass, already queued on `stream`, starts at once), and returns without waiting;
 * `end` waits for that slot and unpacks the results.  While batch i computes, the H2D copy of batch i+1 (other slot)
 * proceeds: this is how a DataLoader loop (cli/test.py:185-191) hides PCIe behind the forward pass.  A slot must be
 * ended before it is begun again; img_host / widths_host must stay valid (and unchanged) until the slot is ended. */
int cocr_recognize_host_begin(cocr_handle *h, int32_t slot, const float *img_host, const int32_t *widths_host,
                              int32_t B, int32_t W, void *stream);
int cocr_recognize_host_end(cocr_handle *h, int32_t slot,
                            int32_t *labels_host, int32_t *starts_host, int32_t *ends_host, float *conf_host,
                            int32_t *counts_host, int32_t *out_lens_host);

/* uint8 input (extends dataset.py:158 / pred.py:119, whose batches are float tensors produced by
 * torchvision's ToTensor from 8-bit line images): the same calls on raw 8-bit pixels.  The front-end kernel scales every
 * pixel by the correctly rounded p / 255.0f, so the results are bit-identical to passing `img.float() / 255`; the
 * host-to-device copy and the image read are four times smaller.  bf16 handles only (COCR_ERR_UNSUPPORTED otherwise).
 * Strides are in elements, as for the float entry points. */
int cocr_forward_decode_u8(cocr_handle *h, const uint8_t *img, int64_t sb, int64_t sw, int64_t sh,
                           const int32_t *widths, int32_t B, int32_t W, void *stream,
                           int32_t *labels, int32_t *starts, int32_t *ends, float *conf,
                           int32_t *counts, int32_t *out_lens);
int cocr_recognize_host_u8_begin(cocr_handle *h, int32_t slot, const uint8_t *img_host, const int32_t *widths_host,
                                 int32_t B, int32_t W, void *stream);   /* collect with cocr_recognize_host_end */

/* Training-side tail of the path (RecognitionModel._step, model.py:135-142).
 * cocr_log_softmax: nn.functional.log_softmax(probits, dim=-1) (model.py:136) over `rows` rows of C classes, fp32.
 * cocr_ctc_loss:    nn.CTCLoss(reduction='sum', zero_infinity=True) (model.py:119,139-142; blank 0) applied to
 *                   log_softmax(logits): logits_dev (B,T,C) RAW decoder outputs (the log-softmax is fused: only its
 *                   per-frame normaliser is computed), targets_dev the 1-D concatenation of the label sequences,
 *                   target_lens_dev / input_lens_dev (B) int32 (input_lens = encoder_lens; NULL = T), max_target_len an
 *                   upper bound of target_lens.  Writes the summed loss (1 float), optionally the per-line losses (B;
 *                   infinite losses already zeroed) and optionally grad_logits_dev (B,T,C) = d loss_sum / d logits, which is
 *                   what autograd passes to the decoder's backward.  The gradient is summed in a fixed order
 *                   (bit-reproducible). */
int cocr_log_softmax(cocr_handle *h, const float *logits_dev, int64_t rows, int32_t C, void *stream, float *out_dev);
int cocr_ctc_loss(cocr_handle *h, const float *logits_dev, const int32_t *targets_dev, const int32_t *target_lens_dev,
                  const int32_t *input_lens_dev, int32_t B, int32_t T, int32_t C, int32_t max_target_len, void *stream,
                  float *loss_sum_dev, float *loss_per_line_dev, float *grad_logits_dev);

/* The decoder nn.Linear(encoder_dim, C) on its own (pred.py:90,121; model.py:114,135), fp32: logits = enc * W^T + b for
 * `rows` encoder frames (enc_dev (rows, encoder_dim) as cocr_encode returns it), and its backward pass given
 * grad_logits_dev (rows, C) (e.g. from cocr_ctc_loss): grad_weight_dev (C, encoder_dim), grad_bias_dev (C) and, when not
 * NULL, grad_enc_dev (rows, encoder_dim) = what flows back into the encoder.  Sums run in a fixed order (bit-reproducible).
 * (The whole training step, encoder backward included, is cocr_train_step below; these two entry points remain for callers
 * that keep the encoder frozen.) */
int cocr_project(cocr_handle *h, const float *enc_dev, int64_t rows, void *stream, float *logits_dev);
int cocr_decoder_backward(cocr_handle *h, const float *grad_logits_dev, const float *enc_dev, int64_t rows, void *stream,
                          float *grad_weight_dev, float *grad_bias_dev, float *grad_enc_dev);

/* Device copy of the records of the batch in `slot` (valid from cocr_recognize_host_begin until the slot is begun again;
 * complete once cocr_recognize_host_end has returned): one block [labels | starts | ends | conf] (4 x (B,T) 32-bit) followed
 * by [counts | out_lens] (2 x (B) int32).  This is what a multi-GPU caller hands to the collective that gathers the decoded
 * records of all shards without a second trip through the host.  The library keeps ownership. */
int cocr_slot_records_dev(cocr_handle *h, int32_t slot, void **records_dev, size_t *bytes, int32_t *B, int32_t *T);

/* Bytes of device workspace the handle holds after a call with this (B, W). */
int cocr_workspace_bytes(cocr_handle *h, int32_t B, int32_t W, size_t *bytes);

/* Which kernels the handle's hyper-parameters select (after cocr_finalize_weights): one "component: kernel" line per component,
 * NUL-terminated in buf; *fallbacks = number of components that run on a slower general kernel because the fast one is built
 * for the default geometry (encoder_dim 144, head dim 36, height 96, ...).  The Python shim warns once when it is non-zero. */
int cocr_kernel_plan(cocr_handle *h, char *buf, size_t cap, int32_t *fallbacks);

/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t cocr_launch_count(const cocr_handle *h);

/* Per-kernel-family timing with CUDA events on the launching stream (bench.py's roofline leg).
 * While enabled, every kernel launch of this handle is bracketed by an event pair.
 * cocr_profile_report synchronises the device and writes one line per family:
 *   "<name> <launches> <total_ms> <algorithmic_flops> <algorithmic_bytes>\n"
 * into buf (NUL-terminated, truncated to cap), then clears the records. */
int cocr_set_profiling(cocr_handle *h, int enabled);
int cocr_profile_report(cocr_handle *h, char *buf, size_t cap);

/* ---- One data-parallel TRAINING step (BASELINE configs[3]) ------------------------------------
 * Replaces, for one rank, RecognitionModel.training_step -> _step (model.py:129-158) + autograd's backward pass +
 * torch.optim.AdamW.step (model.py:272-274): the encoder in train() mode (BatchNorm1d batch statistics with the
 * running-statistics update, convolution.py:141; dropout where the reference has it: feed_forward.py:48-53,
 * attention.py:99,134, convolution.py:146,225), decoder nn.Linear, log_softmax, CTCLoss(reduction='sum',
 * zero_infinity=True), and hand-written backward kernels for every module on the path.  fp32.
 *
 * Parameters / gradients / optimiser moments are FLAT fp32 device buffers owned by the caller.  The layout is exported:
 * cocr_train_layout gives the buffer length (n_total floats), the length of the trainable prefix (n_trainable: what a
 * data-parallel caller all-reduces, what AdamW updates) and the number of entries; cocr_train_entry gives, per entry, the
 * state-dict key (below the ModuleDict: "encoder...." / "decoder....", SURVEY.md 8a "Weights"), offset and element count,
 * every tensor in the reference's own shape and element order.  The BatchNorm running statistics (not trainable) follow the
 * trainable prefix.  The handle needs hyper-parameters only (no cocr_set_weight / cocr_finalize_weights).
 *
 * cocr_train_step: params_dev is read (running statistics are UPDATED in place), grads_dev (n_total floats) is overwritten
 * with d loss / d param.  img/strides/widths as cocr_encode; targets_dev the 1-D int32 concatenation of the label sequences,
 * target_lens_dev (B) int32, max_target_len an upper bound; loss_sum_dev 1 float; loss_per_line_dev (B) or NULL.  The mask of
 * every dropout site is a pure function of (cfg->seed, site, element index) - cocr_dropout_mask returns the factors
 * (0 or 1/(1-p)) so a checker can replay them.  Attention is unmasked, as in the reference (encoder.py:67-103).
 * debug_taps_dev: NULL, or 2 * (1 + 5 * num_layers) * B * T * encoder_dim floats receiving the residual stream after the
 * front-end and after each module of each block, followed by d loss / d (those tensors). */
typedef struct cocr_train_config {
    float input_dropout_p, feed_forward_dropout_p, attention_dropout_p, conv_dropout_p;   /* encoder.py:135-152 */
    float bn_momentum;                                                                   /* nn.BatchNorm1d default 0.1 */
    uint32_t compute;                 /* 0: fp32 CUDA-core kernels (parity path); 1: bf16 operands on tcgen05, fp32 accumulation / master copy */
    uint64_t seed;
} cocr_train_config;

int cocr_train_layout(cocr_handle *h, int64_t *n_total, int64_t *n_trainable, int32_t *n_entries);
int cocr_train_entry(cocr_handle *h, int32_t index, char *key_buf, size_t key_cap, int64_t *offset, int64_t *numel, int32_t *trainable);
int cocr_train_workspace_bytes(cocr_handle *h, int32_t B, int32_t W, size_t *bytes);
int cocr_train_step(cocr_handle *h, float *params_dev, float *grads_dev, const float *img_dev, int64_t stride_b, int64_t stride_w,
                    int64_t stride_h, const int32_t *widths_dev, int32_t B, int32_t W, const int32_t *targets_dev,
                    const int32_t *target_lens_dev, int32_t max_target_len, const cocr_train_config *cfg, void *stream,
                    float *loss_sum_dev, float *loss_per_line_dev, float *debug_taps_dev);
/* torch.optim.AdamW single-tensor arithmetic over n floats; `step` counts from 1. */
int cocr_adamw_step(cocr_handle *h, float *params_dev, const float *grads_dev, float *exp_avg_dev, float *exp_avg_sq_dev, int64_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, void *stream);
int cocr_dropout_mask(cocr_handle *h, uint64_t seed, uint32_t site, float p, int64_t n, void *stream, float *out_dev);

/* Built-in self checks of individual kernels against a scalar device reference; used by
 * tests/ on the GPU box.  `which`: see csrc/selftest.cu.  Returns COCR_OK or an error and
 * writes the max abs error into *max_err (may be NULL). */
int cocr_selftest(int device, int which, int64_t p0, int64_t p1, int64_t p2, double *max_err);

#ifdef __cplusplus
}
#endif
#endif /* COCR_B200_H_ */
