// Shared declarations of the cocr_b200 library (internal; the public ABI is include/cocr_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/cocr_b200.h"

namespace cocr {

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : shape) n *= s;
        return n;
    }
};

// fp32 device pointers of one conformer block (encoder.py:67-100), BatchNorm already folded.
struct LayerF32 {
    const float *ffn_ln_g[2], *ffn_ln_b[2];      // sequential.{0,3}.module.sequential.0
    const float *ffn_w1[2], *ffn_b1[2];          // (ff, d), (ff)
    const float *ffn_w2[2], *ffn_b2[2];          // (d, ff), (d)
    const float *att_ln_g, *att_ln_b;
    const float *w_qkv, *b_qkv;                  // (3d, d) = [Wq; Wk; Wv], (3d)
    const float *w_pos;                          // (d, d), no bias
    const float *u_bias, *v_bias;                // (h, dh)
    const float *w_out, *b_out;                  // (d, d), (d)
    const float *conv_ln_g, *conv_ln_b;
    const float *w_pw1, *b_pw1;                  // (2d, d), (2d)
    const float *w_dw;                           // (d, K)
    const float *bn_scale, *bn_shift;            // folded BatchNorm1d (eval)
    const float *w_pw2, *b_pw2;                  // (d, d), (d)
    const float *final_ln_g, *final_ln_b;
};

struct Dims {
    int H, d, L, nh, dh, ff, K, cc, C;
    int F2, F, feat;                             // height after conv 1 / conv 2, cc*F
    float ffn_factor;
};

struct Workspace {
    char *base = nullptr;
    size_t bytes = 0;
};

struct Bf16State;   // defined in bf16_path.cu
struct TrainState;  // defined in train.cu

}  // namespace cocr

struct cocr_handle {
    cocr_hparams hp;
    cocr::Dims dm;
    int device = 0;
    int num_sms = 148;
    std::string err;
    std::map<std::string, cocr::HostTensor> host_w;
    std::map<std::string, std::vector<int64_t>> expect;   // key -> expected shape
    bool finalized = false;

    // fp32 weights on the device (one arena)
    float *arena = nullptr;
    size_t arena_floats = 0;
    const float *sub_w1 = nullptr, *sub_b1 = nullptr;     // conv.0 (cc,9),(cc)
    const float *sub_wdw = nullptr, *sub_bdw = nullptr;   // conv.2 (cc,9),(cc)
    const float *sub_wpw = nullptr, *sub_bpw = nullptr;   // conv.3 (cc,cc),(cc)
    const float *sub_wout = nullptr, *sub_bout = nullptr; // out.0 permuted to (d, F*cc) [f*cc+c]
    const float *dec_w = nullptr, *dec_b = nullptr;       // (C,d),(C)
    std::vector<cocr::LayerF32> layers;

    // activation workspace (grown on demand) and cached positional projections
    cocr::Workspace ws;
    float *pos_proj = nullptr;      // (L, 2T-1, d) fp32
    int pos_T = 0;

    // staging for the *_host entry points: two slots so that the H2D copy of one batch (copy stream) overlaps the
    // forward pass of the previous one
    struct HostSlot {
        void *img = nullptr;           // device image batch (fp32 or uint8 pixels)
        size_t img_bytes = 0;
        int32_t *widths = nullptr;
        size_t widths_n = 0;
        char *res = nullptr;           // device result block
        char *res_host = nullptr;      // pinned mirror
        size_t res_bytes = 0;
        cudaEvent_t h2d_done = nullptr, fwd_done = nullptr, done = nullptr;
        int B = 0, T = 0;
        bool busy = false;
    } slots[2];
    cudaStream_t copy_stream = nullptr;       // H2D of the batches
    cudaStream_t d2h_stream = nullptr;        // D2H of the result blocks (the next forward pass does not wait for it)

    std::vector<float> fe_weights;   // conv.0 / conv.2 weights + biases in kernel-parameter order (mem_kernels.cu: FrontendWeights)
    int64_t launches = 0;
    cocr::Bf16State *bf = nullptr;
    cocr::TrainState *train = nullptr;   // training step state (train.cu), created on first use

    // optional per-launch event timing (cocr_set_profiling)
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t a, b; double flops, bytes; };
    std::vector<ProfRec> prof;
};

namespace cocr {

int set_error(cocr_handle *h, int code, const char *fmt, ...);
int check_cuda(cocr_handle *h, cudaError_t e, const char *what);
void set_global_error(const char *msg);

#define COCR_CUDA(h, call)                                               \
    do {                                                                 \
        int _rc = cocr::check_cuda((h), (call), #call);                  \
        if (_rc != COCR_OK) return _rc;                                  \
    } while (0)

#define COCR_TRY(call)                   \
    do {                                 \
        int _rc = (call);                \
        if (_rc != COCR_OK) return _rc;  \
    } while (0)

// after-launch check that also counts the launch
#define COCR_LAUNCHED(h)                                                 \
    do {                                                                 \
        (h)->launches++;                                                 \
        int _rc = cocr::check_cuda((h), cudaGetLastError(), "kernel launch"); \
        if (_rc != COCR_OK) return _rc;                                  \
    } while (0)

// RAII event pair around one kernel launch; a no-op unless profiling is on.
struct ProfScope {
    cocr_handle *h;
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    ProfScope(cocr_handle *h_, const char *name, cudaStream_t st_, double flops, double bytes) : h(h_), st(st_) {
        if (!h->profiling) return;
        cudaEvent_t a;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
        h->prof.push_back({name, a, b, flops, bytes});
    }
    ~ProfScope() {
        if (b) cudaEventRecord(b, st);
    }
};

// Function attributes (cudaFuncSetAttribute) belong to a device: the "already set" memo of every launcher is kept per
// device so that handles on several GPUs of one process each get theirs.
template <typename T>
struct PerDevice {
    T v[64] = {};
    T &get() {
        int d = 0;
        cudaGetDevice(&d);
        return v[d & 63];
    }
};

inline int out_len1(int n) { return (n + 2 - 3) / 2 + 1; }   // one stride-2 3x3 pad-1 conv; n >= 1

// ---- fp32 path (fp32_path.cu) ------------------------------------------------------------
struct Fp32Plan {      // workspace carve-up for one (B, W)
    int B, W, T2, T, M;
    size_t off_a1, off_a2, off_feat, off_x, off_xn, off_h, off_g, off_best_val, off_best_idx, total;
};
void fp32_plan(const cocr_handle *h, int B, int W, Fp32Plan *p);
int fp32_encode(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths,
                int B, int W, cudaStream_t st, float *enc_out /* may be null: stays in ws.x */, int32_t *out_lens);
int fp32_logits(cocr_handle *h, const float *enc /* (M,d) */, int M, cudaStream_t st, float *logits);
int launch_out_lens(cocr_handle *h, const int32_t *widths, int B, int W, int32_t *out_lens, cudaStream_t st);
int launch_greedy_decode(cocr_handle *h, const float *logits, const int32_t *lens, int B, int T, int C,
                         cudaStream_t st, float *best_val, int32_t *best_idx, int32_t *labels, int32_t *starts,
                         int32_t *ends, float *conf, int32_t *counts);
int launch_collapse(cocr_handle *h, const float *best_val, const int32_t *best_idx, const int32_t *lens, int B, int T,
                    cudaStream_t st, int32_t *labels, int32_t *starts, int32_t *ends, float *conf, int32_t *counts);
int ensure_workspace(cocr_handle *h, size_t bytes);
int ensure_pos_proj(cocr_handle *h, int T, cudaStream_t st);

// generic fp32 GEMM: C[m,n] = epi(sum_k A[m,k] W[n,k] + bias[n]); act 0 none / 1 SiLU / 2 ReLU;
// res != null: C = res + alpha * value
int launch_sgemm(cocr_handle *h, const float *A, int lda, const float *Wt, int ldw, const float *bias, float *Cout,
                 int ldc, int M, int N, int K, int act, const float *res, int ldres, float alpha, cudaStream_t st,
                 int out_bf16 = 0);
int launch_frontend_convs(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, int B, int W, float *a1,
                          float *a2, cudaStream_t st);
int launch_attn_simt_bf16(cocr_handle *h, const __nv_bfloat16 *qkv, const float *pos, const float *ub, const float *vb,
                          const int32_t *lens, __nv_bfloat16 *ctx, int B, int T, cudaStream_t st);
int launch_glu_dw_bf16(cocr_handle *h, const __nv_bfloat16 *u, const LayerF32 &lw, __nv_bfloat16 *g, int B, int T,
                       cudaStream_t st);
int launch_layernorm(cocr_handle *h, const float *x, float *y, const float *g, const float *b, int M, int d,
                     cudaStream_t st);

// ---- training-side tail (ctc_loss.cu)
int launch_log_softmax(cocr_handle *h, const float *logits, int64_t rows, int C, cudaStream_t st, float *out);
int launch_ctc_loss(cocr_handle *h, const float *logits, const int32_t *targets, const int32_t *target_lens, const int32_t *input_lens,
                    int B, int T, int C, int max_target_len, cudaStream_t st, float *loss_sum, float *loss_per_line, float *grad);

int launch_decoder_backward(cocr_handle *h, const float *grad_logits, const float *enc, int M, cudaStream_t st, float *grad_w, float *grad_b,
                            float *grad_enc);

// ---- batched tcgen05 GEMM (bgemm_tc.cu): D[z] (M x N) = alpha * A[z] (M x K) * B[z] (N x K)^T (+ bias); z = (z1, z2)
struct BgOperand {
    const void *ptr;               // bf16
    int mn;                        // 0: stored [m][k] (contraction index contiguous); 1: stored [k][m]
    long long ld, z1_stride, z2_stride;   // elements; a stride of 0 = shared by the batch
};
struct BgProblem {
    int M = 0, N = 0, K = 0, nz1 = 1, nz2 = 1;
    BgOperand a = {}, b = {};
    void *out = nullptr;
    long long ldc = 0, out_z1 = 0, out_z2 = 0;
    long long out_ks = 0;          // out_mode 0 with a split contraction: element stride between the partial results of the k-slices
    int out_mode = 0;              // 0 fp32 store, 1 bf16 store, 2 fp32 atomic add, 3 fp32 read-add-store
    float alpha = 1.0f;
    const float *bias = nullptr;   // (N)
    int ksplit = 1;                // split of the contraction over CTAs (needs out_mode 2)
    int relu = 0;
};
int launch_bgemm(cocr_handle *h, const BgProblem &g, cudaStream_t st, const char *tag);
int selftest_bgemm(int device, int M, int N, int K, int flags, double *max_err);

// ---- training step (train.cu): flat parameter / gradient buffers owned by the caller
void train_destroy(cocr_handle *h);
int train_layout(cocr_handle *h, int64_t *n_total, int64_t *n_trainable, int32_t *n_entries);
int train_entry(cocr_handle *h, int32_t i, char *key, size_t cap, int64_t *off, int64_t *numel, int32_t *trainable);
int train_workspace_bytes(cocr_handle *h, int B, int W, size_t *bytes);
int train_step(cocr_handle *h, float *params, float *grads, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int B, int W,
               const int32_t *targets, const int32_t *target_lens, int max_target_len, const cocr_train_config *cfg, cudaStream_t st, float *loss_out,
               float *loss_per_line, float *taps);
int adamw_step(cocr_handle *h, float *p, const float *g, float *m, float *v, int64_t n, float lr, float b1, float b2, float eps, float wd, int step,
               cudaStream_t st);
int dropout_mask(cocr_handle *h, uint64_t seed, uint32_t site, float prob, int64_t n, cudaStream_t st, float *out);

// ---- optimised memory-bound kernels (mem_kernels.cu); return COCR_ERR_UNSUPPORTED when the geometry is not covered
int launch_glu_dw_fast(cocr_handle *h, const __nv_bfloat16 *u, const LayerF32 &lw, __nv_bfloat16 *g, int B, int T,
                       cudaStream_t st);
int launch_dw_fast(cocr_handle *h, const __nv_bfloat16 *x, const float *taps_dev, const float *hshift_dev, __nv_bfloat16 *g, int B, int T,
                   cudaStream_t st);
int launch_frontend_convs2(cocr_handle *h, const void *img, bool u8, int64_t sb, int64_t sw, int64_t sh, int B, int W, __nv_bfloat16 *a2,
                           cudaStream_t st);
// tensor-core front-end (frontend_tc.cu): image -> feature rows in one kernel; wpw2 = (32, 64) bf16 [Wpw | Wpw]
int launch_frontend_tc(cocr_handle *h, const void *img, bool u8, int64_t sb, int64_t sw, int64_t sh, int B, int W, const __nv_bfloat16 *wpw2,
                       __nv_bfloat16 *feat, cudaStream_t st);
int launch_frontend_fused(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, int B, int W,
                          __nv_bfloat16 *feat, cudaStream_t st);

// ---- bf16 path (bf16_path.cu) ------------------------------------------------------------
int bf16_init(cocr_handle *h);       // pack weights; called from cocr_finalize_weights when precision == bf16
void bf16_destroy(cocr_handle *h);
int bf16_kernel_plan(const cocr_handle *h, char *buf, size_t cap, int *fallbacks);
int bf16_encode(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths,
                int B, int W, cudaStream_t st, float *enc_out, int32_t *out_lens);
size_t bf16_workspace_bytes(const cocr_handle *h, int B, int W);
// decode straight from the bf16 path's final activations (fused projection + argmax)
// img: fp32 pixels, or uint8 pixels when `u8` (scaled by 1/255 inside the first kernel)
int bf16_forward_decode(cocr_handle *h, const void *img, bool u8, int64_t sb, int64_t sw, int64_t sh,
                        const int32_t *widths, int B, int W, cudaStream_t st, int32_t *labels, int32_t *starts,
                        int32_t *ends, float *conf, int32_t *counts, int32_t *out_lens);
int bf16_logits(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths,
                int B, int W, cudaStream_t st, float *logits, int32_t *out_lens);

}  // namespace cocr
