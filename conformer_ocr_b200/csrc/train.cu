// One training step of RecognitionModel (model.py:129-145) on one GPU: train()-mode forward (BatchNorm batch statistics,
// dropout), log_softmax + CTCLoss(sum, zero_infinity), and the hand-written backward pass through the decoder, the 16
// conformer blocks and the convolutional front-end; plus the AdamW update (model.py:272-274).  Two arithmetic modes
// (cocr_train_config.compute): 0 = fp32 CUDA-core kernels (this file + train_kernels.cuh: the parity path), 1 = every dense
// contraction on tcgen05 with bf16 operands through the batched GEMM of bgemm_tc.cu (train_tc.cuh), fp32 everywhere else.
//
// Parameters, gradients and optimiser moments live in FLAT fp32 buffers owned by the caller (torch tensors); the layout
// (state-dict key -> offset) is exported by cocr_train_entry.  Trainable tensors come first, the BatchNorm running
// statistics last, so one all-reduce over [0, n_trainable) of the gradient buffer is the whole data-parallel exchange.
#include <math.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "train_kernels.cuh"

namespace cocr {

struct TrainLayerOff {
    size_t ffn_ln_g[2], ffn_ln_b[2], ffn_w1[2], ffn_b1[2], ffn_w2[2], ffn_b2[2];
    size_t att_ln_g, att_ln_b, w_qkv, b_qkv, w_pos, u_bias, v_bias, w_out, b_out;
    size_t conv_ln_g, conv_ln_b, w_pw1, b_pw1, w_dw, bn_g, bn_b, w_pw2, b_pw2;
    size_t fin_g, fin_b;
    size_t bn_rm, bn_rv;
};

struct TrainEntry {
    std::string key;
    size_t off, numel;
    bool trainable;
};

struct TrainState {
    std::vector<TrainEntry> entries;
    size_t n_total = 0, n_trainable = 0;
    size_t sub_w1, sub_b1, sub_wdw, sub_bdw, sub_wpw, sub_bpw, sub_wout, sub_bout, dec_w, dec_b;
    std::vector<TrainLayerOff> L;
    char *ws = nullptr;
    size_t ws_bytes = 0;
    float *pe = nullptr;        // (2T-1, d) relative positional table of the running step (an entry of `pes`)
    int pe_T = 0;
    std::map<int, float *> pes; // one table per frame count, kept until the handle dies: captured steps hold their addresses
    uint64_t *seed_dev = nullptr;   // [0]: seed of the running step, [1]: seed of cocr_dropout_mask
    // the launch sequence of one step, captured per argument set and replayed (about 1 800 launches per step otherwise)
    struct StepGraph {
        std::vector<uint64_t> key;
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
        int sightings = 0;
    };
    std::vector<StepGraph> graphs;
    cudaStream_t gstream = nullptr;
    cudaEvent_t g_in = nullptr, g_out = nullptr;
    cudaStream_t side = nullptr;            // weight-gradient stream of the tensor-core path
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {

using namespace train;

size_t add_entry(TrainState *s, const std::string &key, size_t numel, bool trainable, size_t &cursor) {
    const size_t off = cursor;
    s->entries.push_back({key, off, numel, trainable});
    cursor += numel;
    return off;
}

void build_layout(const Dims &dm, TrainState *s) {
    const size_t d = dm.d, ff = dm.ff, cc = dm.cc, K = dm.K;
    size_t cur = 0;
    auto pad = [&]() { cur = (cur + 7) / 8 * 8; };       // 16 bytes in the bf16 mirror of the buffer (TMA operand alignment)
    auto add = [&](const std::string &k, size_t n) { pad(); return add_entry(s, k, n, true, cur); };
    const std::string sp = "encoder.conv_subsample.";
    s->sub_w1 = add(sp + "conv.0.weight", cc * 9);
    s->sub_b1 = add(sp + "conv.0.bias", cc);
    s->sub_wdw = add(sp + "conv.2.weight", cc * 9);
    s->sub_bdw = add(sp + "conv.2.bias", cc);
    s->sub_wpw = add(sp + "conv.3.weight", cc * cc);
    s->sub_bpw = add(sp + "conv.3.bias", cc);
    s->sub_wout = add(sp + "out.0.weight", d * dm.feat);
    s->sub_bout = add(sp + "out.0.bias", d);
    s->L.assign(dm.L, TrainLayerOff());
    for (int l = 0; l < dm.L; ++l) {
        TrainLayerOff &o = s->L[l];
        const std::string p = "encoder.layers." + std::to_string(l) + ".sequential.";
        int f = 0;
        for (const char *fi : {"0", "3"}) {
            const std::string q = p + fi + ".module.sequential.";
            o.ffn_ln_g[f] = add(q + "0.weight", d);
            o.ffn_ln_b[f] = add(q + "0.bias", d);
            o.ffn_w1[f] = add(q + "1.linear.weight", ff * d);
            o.ffn_b1[f] = add(q + "1.linear.bias", ff);
            o.ffn_w2[f] = add(q + "4.linear.weight", d * ff);
            o.ffn_b2[f] = add(q + "4.linear.bias", d);
            ++f;
        }
        std::string q = p + "1.module.";
        o.att_ln_g = add(q + "layer_norm.weight", d);
        o.att_ln_b = add(q + "layer_norm.bias", d);
        // query | key | value stay contiguous (d % 4 == 0 is checked by the caller): one (3d, d) GEMM
        o.w_qkv = add(q + "attention.query_proj.linear.weight", d * d);
        add(q + "attention.key_proj.linear.weight", d * d);
        add(q + "attention.value_proj.linear.weight", d * d);
        o.b_qkv = add(q + "attention.query_proj.linear.bias", d);
        add(q + "attention.key_proj.linear.bias", d);
        add(q + "attention.value_proj.linear.bias", d);
        o.w_pos = add(q + "attention.pos_proj.linear.weight", d * d);
        o.u_bias = add(q + "attention.u_bias", d);
        o.v_bias = add(q + "attention.v_bias", d);
        o.w_out = add(q + "attention.out_proj.linear.weight", d * d);
        o.b_out = add(q + "attention.out_proj.linear.bias", d);
        q = p + "2.module.sequential.";
        o.conv_ln_g = add(q + "0.weight", d);
        o.conv_ln_b = add(q + "0.bias", d);
        o.w_pw1 = add(q + "2.conv.weight", 2 * d * d);
        o.b_pw1 = add(q + "2.conv.bias", 2 * d);
        o.w_dw = add(q + "4.conv.weight", d * K);
        o.bn_g = add(q + "5.weight", d);
        o.bn_b = add(q + "5.bias", d);
        o.w_pw2 = add(q + "7.conv.weight", d * d);
        o.b_pw2 = add(q + "7.conv.bias", d);
        o.fin_g = add(p + "4.weight", d);
        o.fin_b = add(p + "4.bias", d);
    }
    s->dec_w = add("decoder.weight", (size_t)dm.C * d);
    s->dec_b = add("decoder.bias", (size_t)dm.C);
    pad();
    s->n_trainable = cur;
    for (int l = 0; l < dm.L; ++l) {
        const std::string q = "encoder.layers." + std::to_string(l) + ".sequential.2.module.sequential.";
        pad();
        s->L[l].bn_rm = add_entry(s, q + "5.running_mean", d, false, cur);
        pad();
        s->L[l].bn_rv = add_entry(s, q + "5.running_var", d, false, cur);
    }
    pad();
    s->n_total = cur;
}

int train_state(cocr_handle *h, TrainState **out) {
    if (!h->train) {
        const Dims &dm = h->dm;
        if (dm.d % 8 != 0 || 256 % dm.cc != 0 || dm.C < 2)
            return set_error(h, COCR_ERR_UNSUPPORTED, "training needs encoder_dim %% 8 == 0, conv channels dividing 256 and a decoder (d=%d cc=%d C=%d)",
                             dm.d, dm.cc, dm.C);
        if (dm.dh > 64) return set_error(h, COCR_ERR_UNSUPPORTED, "training attention supports head dim <= 64 (got %d)", dm.dh);
        h->train = new TrainState();
        build_layout(dm, h->train);
        void *sp = nullptr;
        COCR_CUDA(h, cudaMalloc(&sp, 16));
        h->train->seed_dev = (uint64_t *)sp;
    }
    *out = h->train;
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
inline unsigned ew_grid(int64_t n) { int64_t g = (n + 255) / 256; return (unsigned)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g)); }

struct Ctx {
    cocr_handle *h;
    cudaStream_t st;
    float *part;           // split-K partial sums
    size_t part_floats;
    // Weight gradients only feed the optimiser at the end of the step, so the tensor-core path runs them on a side stream next to
    // the data-gradient chain (fork after their operands exist, join before an operand buffer is written again).
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_busy = false;
    int fork() {           // side stream waits for everything enqueued on the main stream so far
        if (!side) return COCR_OK;
        COCR_CUDA(h, cudaEventRecord(ev_fork, st));
        COCR_CUDA(h, cudaStreamWaitEvent(side, ev_fork, 0));
        side_busy = true;
        return COCR_OK;
    }
    int join() {           // main stream waits for the side stream
        if (!side || !side_busy) return COCR_OK;
        COCR_CUDA(h, cudaEventRecord(ev_join, side));
        COCR_CUDA(h, cudaStreamWaitEvent(st, ev_join, 0));
        side_busy = false;
        return COCR_OK;
    }
    cudaStream_t wstream() const { return side ? side : st; }
};

#define TK_LAUNCHED(c) COCR_LAUNCHED((c).h)

// C = act(A W^T + bias): A (M,K) lda, W (N,K) ldw
int gemm_nt(Ctx &c, const float *A, int lda, const float *W, int ldw, const float *bias, float *C, int ldc, int M, int N, int K, int act,
            const char *tag = "train_gemm_nt") {
    if (M <= 0 || N <= 0) return COCR_OK;
    dim3 grid((M + 63) / 64, (N + 63) / 64, 1);
    ProfScope ps(c.h, tag, c.st, 2.0 * M * (double)N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N));
    k_gemm_f32<true, true><<<grid, 256, 0, c.st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, K, 0, act);
    TK_LAUNCHED(c);
    return COCR_OK;
}
// dX (M,K) = dY (M,N) W (N,K)
int gemm_nn(Ctx &c, const float *dY, int ldy, const float *W, int ldw, float *dX, int ldx, int M, int N, int K) {
    if (M <= 0 || K <= 0) return COCR_OK;
    dim3 grid((M + 63) / 64, (K + 63) / 64, 1);
    ProfScope ps(c.h, "train_gemm_dgrad", c.st, 2.0 * M * (double)N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N));
    k_gemm_f32<true, false><<<grid, 256, 0, c.st>>>(dY, ldy, W, ldw, nullptr, dX, ldx, M, K, N, N, 0, 0);
    TK_LAUNCHED(c);
    return COCR_OK;
}
// dW (N,K) = dY^T (N,M) X (M,K), split over M, summed in slice order
int gemm_tn(Ctx &c, const float *dY, int ldy, const float *X, int ldx, float *dW, int M, int N, int K) {
    if (N <= 0 || K <= 0) return COCR_OK;
    const int gx = (N + 63) / 64, gy = (K + 63) / 64;
    int Z = (148 * 4) / (gx * gy);
    if (Z < 1) Z = 1;
    if (Z > 64) Z = 64;
    const int maxz = (M + 63) / 64;
    if (Z > maxz) Z = maxz < 1 ? 1 : maxz;
    while (Z > 1 && (size_t)Z * N * K > c.part_floats) --Z;
    int kchunk = ((M + Z - 1) / Z + 15) / 16 * 16;
    Z = (M + kchunk - 1) / kchunk;
    if (Z < 1) Z = 1;
    ProfScope ps(c.h, "train_gemm_wgrad", c.st, 2.0 * M * (double)N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N));
    if (Z == 1) {
        k_gemm_f32<false, false><<<dim3(gx, gy, 1), 256, 0, c.st>>>(dY, ldy, X, ldx, nullptr, dW, K, N, K, M, kchunk, 0, 0);
        TK_LAUNCHED(c);
        return COCR_OK;
    }
    k_gemm_f32<false, false><<<dim3(gx, gy, Z), 256, 0, c.st>>>(dY, ldy, X, ldx, nullptr, c.part, K, N, K, M, kchunk, (size_t)N * K, 0);
    TK_LAUNCHED(c);
    k_reduce_z<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, c.st>>>(c.part, Z, (int64_t)N * K, dW);
    TK_LAUNCHED(c);
    return COCR_OK;
}
int colsum(Ctx &c, const float *A, int lda, int M, int N, float *out) {
    const int rpb = 32;
    dim3 grid((M + rpb - 1) / rpb, (N + 255) / 256);
    k_colsum<<<grid, 256, 0, c.st>>>(A, lda, M, N, rpb, out);
    TK_LAUNCHED(c);
    return COCR_OK;
}
int ln_fwd(Ctx &c, const float *x, const float *g, const float *b, float *y, float *stats, int M, int d) {
    ProfScope ps(c.h, "train_layernorm", c.st, 0.0, 8.0 * M * d);
    k_ln_fwd<<<(M + 7) / 8, 256, 0, c.st>>>(x, g, b, y, stats, M, d);
    TK_LAUNCHED(c);
    return COCR_OK;
}
// parameter gradients of a LayerNorm and dx_out = resid + LN'(dy)
int ln_bwd(Ctx &c, const float *dy, const float *x, const float *stats, const float *g, const float *resid, float *dx_out, float *dgamma,
           float *dbeta, int M, int d) {
    ProfScope ps(c.h, "train_layernorm_bwd", c.st, 0.0, 16.0 * M * d);
    if (d % 4 == 0 && d <= 256) {
        const int rpb = 32;
        k_ln_bwd_fused_v4<<<(M + rpb - 1) / rpb, 256, 0, c.st>>>(dy, x, stats, g, resid, dx_out, M, d, rpb, dgamma, dbeta);
        TK_LAUNCHED(c);
        return COCR_OK;
    }
    if (d <= 32 * LNB_MAXC) {
        const int rpb = 32;
        k_ln_bwd_fused<<<(M + rpb - 1) / rpb, 256, 0, c.st>>>(dy, x, stats, g, resid, dx_out, M, d, rpb, dgamma, dbeta);
        TK_LAUNCHED(c);
        return COCR_OK;
    }
    const int rpb = 32;
    k_ln_param_grad<<<dim3((M + rpb - 1) / rpb, (d + 255) / 256), 256, 0, c.st>>>(dy, x, stats, M, d, rpb, dgamma, dbeta);
    TK_LAUNCHED(c);
    k_ln_bwd<<<(M + 7) / 8, 256, 0, c.st>>>(dy, x, stats, g, resid, dx_out, M, d);
    TK_LAUNCHED(c);
    return COCR_OK;
}

template <typename F>
int set_smem(cocr_handle *h, F *fn, size_t bytes) {
    if (bytes > 48 * 1024) COCR_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return COCR_OK;
}

template <int DH>
int att_forward(Ctx &c, const float *qkv, const float *pos, const float *ub, const float *vb, float *P, float *ctx, int B, int T, int d, int nh,
                int dh, const Drop &dr) {
    const size_t LD = DH + 1;
    const dim3 grid((T + TQ - 1) / TQ, nh, B);
    {
        const size_t smem = 4 * ((size_t)TK * LD + (size_t)(TQ + TK - 1) * LD + (size_t)TQ * (TK + 1));
        COCR_TRY(set_smem(c.h, k_att_scores<DH>, smem));
        ProfScope ps(c.h, "train_att_scores", c.st, 4.0 * B * T * (double)T * d, 4.0 * B * nh * (double)T * T);
        k_att_scores<DH><<<grid, 256, smem, c.st>>>(qkv, pos, ub, vb, P, T, d, dh, 1.0f / sqrtf((float)dh));
        TK_LAUNCHED(c);
    }
    {
        const int64_t rows = (int64_t)B * nh * T;
        ProfScope ps(c.h, "train_att_softmax", c.st, 0.0, 8.0 * rows * T);
        k_softmax_rows<<<(unsigned)((rows + 7) / 8), 256, 0, c.st>>>(P, rows, T);
        TK_LAUNCHED(c);
    }
    {
        const size_t smem = 4 * ((size_t)TK * LD + (size_t)TQ * (TK + 1));
        COCR_TRY(set_smem(c.h, k_att_pv<DH>, smem));
        ProfScope ps(c.h, "train_att_pv", c.st, 2.0 * B * T * (double)T * d, 4.0 * B * nh * (double)T * T);
        k_att_pv<DH><<<grid, 256, smem, c.st>>>(qkv, P, ctx, T, d, dh, dr);
        TK_LAUNCHED(c);
    }
    return COCR_OK;
}

template <int DH>
int att_backward(Ctx &c, const float *qkv, const float *pos, const float *ub, const float *vb, const float *P, const float *ctx, const float *dctx,
                 float *dS, float *dqkv, float *dpos, float *du, float *dvb, int B, int T, int d, int nh, int dh, const Drop &dr) {
    const size_t LD = DH + 1;
    const dim3 grid((T + TQ - 1) / TQ, nh, B);
    const float inv_sqrt = 1.0f / sqrtf((float)dh);
    {
        const size_t smem = 4 * ((size_t)2 * TK * LD + (size_t)(TQ + TK - 1) * LD + (size_t)TQ * (TK + 1));
        COCR_TRY(set_smem(c.h, k_att_bwd_rows<DH>, smem));
        ProfScope ps(c.h, "train_att_bwd_rows", c.st, 6.0 * B * T * (double)T * d, 8.0 * B * nh * (double)T * T);
        k_att_bwd_rows<DH><<<grid, 256, smem, c.st>>>(qkv, pos, P, ctx, dctx, dS, dqkv, du, dvb, T, d, dh, inv_sqrt, dr);
        TK_LAUNCHED(c);
    }
    {
        const size_t smem = 4 * (size_t)256 * LD;
        COCR_TRY(set_smem(c.h, k_att_bwd_cols<DH>, smem));
        ProfScope ps(c.h, "train_att_bwd_cols", c.st, 4.0 * B * T * (double)T * d, 8.0 * B * nh * (double)T * T);
        k_att_bwd_cols<DH><<<dim3((T + 31) / 32, nh, B), 256, smem, c.st>>>(qkv, ub, P, dS, dctx, dqkv, T, d, dh, dr);
        TK_LAUNCHED(c);
    }
    {
        const int R = 2 * T - 1;
        COCR_CUDA(c.h, cudaMemsetAsync(dpos, 0, (size_t)R * d * 4, c.st));
        const int bpb = B >= 8 ? (B + 7) / 8 : 1;
        const size_t smem = 4 * (size_t)256 * (DH + 4);
        COCR_TRY(set_smem(c.h, k_att_bwd_pos<DH>, smem));
        ProfScope ps(c.h, "train_att_bwd_pos", c.st, 2.0 * B * T * (double)T * d, 4.0 * B * nh * (double)T * T);
        k_att_bwd_pos<DH><<<dim3((R + 31) / 32, nh, (B + bpb - 1) / bpb), 256, smem, c.st>>>(qkv, vb, dS, dpos, B, T, d, dh, nh, bpb);
        TK_LAUNCHED(c);
    }
    return COCR_OK;
}

#define DISPATCH_DH(dh, call36, call16, call64) ((dh) <= 16 ? (call16) : (dh) <= 36 ? (call36) : (call64))

int dwconv(Ctx &c, const float *x, const float *w, float *y, int B, int T, int d, int K, int flip) {
    if (K == 31 && d % 2 == 0) {       // register sliding window: 8 outputs per thread and channel
        constexpr int TT = 64;
        const size_t smem = 4 * ((size_t)(TT + 30) * d + (size_t)31 * d);
        COCR_TRY(set_smem(c.h, k_dwconv1d_k31, smem));
        ProfScope ps(c.h, flip ? "train_dwconv_dgrad" : "train_dwconv", c.st, 2.0 * K * B * (double)T * d, 8.0 * B * T * d);
        k_dwconv1d_k31<<<dim3((T + TT - 1) / TT, B), 256, smem, c.st>>>(x, w, y, T, d, flip);
        TK_LAUNCHED(c);
        return COCR_OK;
    }
    const size_t smem = 4 * (size_t)(DWT + 2 * K - 1) * d;
    if (smem > 200 * 1024) return set_error(c.h, COCR_ERR_UNSUPPORTED, "depthwise conv tile needs %zu B of shared memory", smem);
    COCR_TRY(set_smem(c.h, k_dwconv1d, smem));
    ProfScope ps(c.h, flip ? "train_dwconv_dgrad" : "train_dwconv", c.st, 2.0 * K * B * (double)T * d, 8.0 * B * T * d);
    k_dwconv1d<<<dim3((T + DWT - 1) / DWT, B), 256, smem, c.st>>>(x, w, y, T, d, K, flip);
    TK_LAUNCHED(c);
    return COCR_OK;
}
int dwconv_wgrad(Ctx &c, const float *x, const float *dy, float *dw, int B, int T, int d, int K) {
    if (K == 31 && 4 * (size_t)(64 + 30 + 64 + 31) * d <= 200 * 1024) {
        const size_t smem = 4 * (size_t)(64 + 30 + 64 + 31) * d;
        COCR_TRY(set_smem(c.h, k_dwconv1d_wgrad_k31, smem));
        ProfScope ps(c.h, "train_dwconv_wgrad", c.st, 2.0 * K * B * (double)T * d, 8.0 * B * T * d);
        k_dwconv1d_wgrad_k31<<<dim3((T + 63) / 64, B), 256, smem, c.st>>>(x, dy, dw, T, d);
        TK_LAUNCHED(c);
        return COCR_OK;
    }
    int tile = 64;                                   // frames per CTA: larger tiles = fewer atomics per output, fewer CTAs
    while (tile > 16 && (4 * (size_t)(2 * tile + K - 1) * d > 200 * 1024 || tile >= 2 * T)) tile /= 2;
    const size_t smem = 4 * (size_t)(2 * tile + K - 1) * d;
    if (smem > 200 * 1024) return set_error(c.h, COCR_ERR_UNSUPPORTED, "depthwise conv tile needs %zu B of shared memory", smem);
    COCR_TRY(set_smem(c.h, k_dwconv1d_wgrad, smem));
    ProfScope ps(c.h, "train_dwconv_wgrad", c.st, 2.0 * K * B * (double)T * d, 8.0 * B * T * d);
    k_dwconv1d_wgrad<<<dim3((T + tile - 1) / tile, B), 256, smem, c.st>>>(x, dy, dw, T, d, K, tile);
    TK_LAUNCHED(c);
    return COCR_OK;
}

const uint64_t *g_seedp = nullptr;    // device seed word of the step being enqueued (set by train_step / dropout_mask)

Drop make_drop(const cocr_train_config *cfg, float p, uint32_t site) {
    Drop dr;
    (void)cfg;
    dr.seedp = g_seedp;
    dr.site = site;
    if (!(p > 0.0f)) { dr.thresh = 0; dr.inv_keep = 1.0f; return dr; }
    double t = (double)p * 4294967296.0;
    if (t > 4294967295.0) t = 4294967295.0;
    if (t < 1.0) t = 1.0;
    dr.thresh = (uint32_t)t;
    dr.inv_keep = (float)(1.0 / (1.0 - (double)p));
    return dr;
}

struct Bump {
    size_t off = 0;
    size_t take(size_t floats) { const size_t o = off; off = (off + floats * 4 + 255) / 256 * 256; return o; }
    size_t take16(size_t halves) { return take((halves + 1) / 2); }
};

// workspace carve-up of one (B, W, max target length)
struct Plan {
    int B, W, T2, T, M, R;
    size_t xs, stats, h1, qkv, pos, P, ctx, u, cdw, bnstat;     // saved for the backward pass (per layer strides below)
    size_t a1, a2, a3, wperm, logits, glogits, lens;
    size_t t_d0, t_d1, t_d2, t_ff0, t_ff1, t_qkv, dS, dpos, part, bnsums, dx, dxn, gwperm;
    size_t part_floats;
    // tensor-core path (bf16 operands): tc == 1
    int tc, Tp, Rp, dhp, HP, Cp;
    size_t a3b, wpermb, pb, peb, xnb, zb, actb, dyb, wideb, ctxb, encb, glogb, dposb, qu, qv, kb, vb, dctxb, posb, Sc, PS, dSb, dPSb, delta, Pb16;
    size_t total;
};

void make_plan(const Dims &dm, int B, int W, int tc, size_t n_total, Plan *p) {
    p->B = B; p->W = W;
    p->T2 = out_len1(W);
    p->T = out_len1(p->T2);
    p->M = B * p->T;
    p->R = 2 * p->T - 1;
    const size_t M = p->M, d = dm.d, L = dm.L, ff = dm.ff;
    Bump b;
    p->xs = b.take((1 + 5 * L) * M * d);
    p->stats = b.take(5 * L * M * 2);
    p->h1 = b.take(2 * L * M * ff);
    p->qkv = b.take((tc ? 1 : L) * M * 3 * d);
    p->pos = b.take((tc ? 1 : L) * (size_t)p->R * d);
    p->tc = tc;
    p->Tp = (p->T + 7) / 8 * 8;
    p->Rp = (p->R + 7) / 8 * 8;
    p->dhp = (dm.dh + 7) / 8 * 8;
    p->HP = dm.nh * p->dhp;
    p->Cp = (dm.C + 7) / 8 * 8;
    p->P = tc ? 0 : b.take(L * (size_t)B * dm.nh * p->T * p->T);
    p->ctx = b.take(L * M * d);
    p->u = b.take(L * M * 2 * d);
    p->cdw = b.take(L * M * d);
    p->bnstat = b.take(L * 2 * d);
    p->a1 = b.take((size_t)B * p->T2 * dm.F2 * dm.cc);
    p->a2 = b.take(M * dm.feat);
    p->a3 = b.take(M * dm.feat);
    p->wperm = b.take(d * dm.feat);
    p->gwperm = b.take(d * dm.feat);
    p->logits = b.take(M * (size_t)dm.C);
    p->glogits = b.take(M * (size_t)dm.C);
    p->lens = b.take((size_t)B + 64);
    p->t_d0 = b.take(M * d);
    p->t_d1 = b.take(M * d);
    p->t_d2 = b.take(M * d);
    const size_t wide = ff > 2 * d ? ff : 2 * d;
    p->t_ff0 = b.take(M * (wide > (size_t)dm.feat ? wide : (size_t)dm.feat));
    p->t_ff1 = b.take(M * (wide > (size_t)dm.feat ? wide : (size_t)dm.feat));
    p->t_qkv = b.take(M * 3 * d);
    p->dS = tc ? 0 : b.take((size_t)B * dm.nh * p->T * p->T);
    if (tc) {
        const size_t bh = (size_t)B * dm.nh, R = p->R;
        p->Pb16 = b.take16(2 * L * bh * p->T * p->Tp);        // per layer: P then dropout(P)
        p->pb = b.take16(n_total);
        p->a3b = b.take16(M * dm.feat);
        p->wpermb = b.take16(d * dm.feat);
        p->peb = b.take16(R * d);
        // operands saved per layer for the backward pass (bf16): LayerNorm outputs of the four modules, FFN activations,
        // conv-module activation, attention context and the head-pitched q+u | q+v | k | v | p
        p->xnb = b.take16(4 * L * M * d);
        p->zb = b.take16(L * M * d);
        p->actb = b.take16(2 * L * M * ff);
        p->dyb = b.take16(M * d);
        p->wideb = b.take16(M * (ff > 3 * d ? ff : 3 * d));
        p->ctxb = b.take16(L * M * d);
        p->encb = b.take16(M * d);
        p->glogb = b.take16(M * (size_t)p->Cp);
        p->dposb = b.take16(R * d);
        p->qu = b.take16(L * M * p->HP);
        p->qv = b.take16(L * M * p->HP);
        p->kb = b.take16(L * M * p->HP);
        p->vb = b.take16(L * M * p->HP);
        p->dctxb = b.take16(M * p->HP);
        p->posb = b.take16(L * R * p->HP);
        p->Sc = b.take(bh * p->T * p->Tp);
        p->PS = b.take(bh * p->T * p->Rp);
        p->dSb = b.take16(bh * p->T * p->Tp);
        p->dPSb = b.take16(bh * p->T * p->Rp);
        p->delta = b.take(M * dm.nh);
    }
    p->dpos = b.take((size_t)p->R * d);
    size_t biggest = ff * d;
    if ((size_t)dm.C * d > biggest) biggest = (size_t)dm.C * d;
    if (d * (size_t)dm.feat > biggest) biggest = d * (size_t)dm.feat;
    if (3 * d * d > biggest) biggest = 3 * d * d;
    p->part_floats = 64 * biggest;
    p->part = b.take(p->part_floats);
    p->bnsums = b.take(4 * d + 16);       // doubles: 2*d, as floats 4*d
    p->dx = b.take(M * d);
    p->dxn = b.take(M * d);
    p->total = b.off;
}

int ensure_train_ws(cocr_handle *h, TrainState *s, size_t bytes) {
    if (bytes <= s->ws_bytes) return COCR_OK;
    if (s->ws) {
        COCR_CUDA(h, cudaDeviceSynchronize());
        COCR_CUDA(h, cudaFree(s->ws));
        s->ws = nullptr;
        s->ws_bytes = 0;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return set_error(h, COCR_ERR_OOM, "CUDA out of memory. training workspace of %zu bytes", bytes);
    }
    COCR_CUDA(h, e);
    s->ws = (char *)p;
    s->ws_bytes = bytes;
    return COCR_OK;
}

int ensure_pe(cocr_handle *h, TrainState *s, int T, cudaStream_t st) {
    // A table is written once and never freed or rewritten while the handle lives: the step graphs captured for a width bake
    // its address in, and batches of alternating widths must not stall on a device synchronise + free + malloc per step.
    auto it = s->pes.find(T);
    if (it != s->pes.end()) {
        s->pe = it->second;
        s->pe_T = T;
        return COCR_OK;
    }
    const int R = 2 * T - 1, d = h->dm.d;
    void *p = nullptr;
    COCR_CUDA(h, cudaMalloc(&p, (size_t)R * d * 4));
    // embedding.py:42-56 in the reference's own fp32 steps (host), as ensure_pos_proj does for the inference path
    std::vector<float> pe((size_t)R * d);
    const float neg = (float)(-(log(10000.0) / (double)d));
    for (int k = 0; k < R; ++k) {
        const float r = (float)(T - 1 - k);
        for (int m = 0; m < d; m += 2) {
            const float arg = r * expf((float)m * neg);
            pe[(size_t)k * d + m] = sinf(arg);
            if (m + 1 < d) pe[(size_t)k * d + m + 1] = cosf(arg);
        }
    }
    cudaError_t e = cudaMemcpyAsync(p, pe.data(), pe.size() * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);        // `pe` is pageable host memory
    if (e != cudaSuccess) {
        cudaFree(p);
        COCR_CUDA(h, e);
    }
    s->pes[T] = s->pe = (float *)p;
    s->pe_T = T;
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// modules: forward (saves what the backward pass needs) and backward
// ------------------------------------------------------------------------------------------
struct Mod {
    Ctx &c;
    const Dims &dm;
    const Plan &p;
    char *ws;
    const float *P;      // parameters
    float *G;            // gradients
    const cocr_train_config *cfg;
    int l;
    float *f(size_t off) const { return (float *)(ws + off); }
    float *xs(int idx) const { return f(p.xs) + (size_t)idx * p.M * dm.d; }
    float *stats(int m) const { return f(p.stats) + ((size_t)l * 5 + m) * p.M * 2; }
    Drop drop(float prob, int site) const { return make_drop(cfg, prob, (uint32_t)(l * 16 + site + 1)); }
};

int ffn_forward(Mod &m, int half, const float *x_in, float *x_out) {
    const Dims &dm = m.dm;
    const TrainLayerOff &o = m.c.h->train->L[m.l];
    const int M = m.p.M, d = dm.d, ff = dm.ff;
    float *xn = m.f(m.p.t_d0), *act = m.f(m.p.t_ff0), *y = m.f(m.p.t_d1);
    float *h1 = m.f(m.p.h1) + ((size_t)m.l * 2 + half) * M * ff;
    COCR_TRY(ln_fwd(m.c, x_in, m.P + o.ffn_ln_g[half], m.P + o.ffn_ln_b[half], xn, m.stats(half ? 3 : 0), M, d));
    COCR_TRY(gemm_nt(m.c, xn, d, m.P + o.ffn_w1[half], d, m.P + o.ffn_b1[half], h1, ff, M, ff, d, 0, "train_ffn_up"));
    k_silu_drop<<<ew_grid((int64_t)M * ff), 256, 0, m.c.st>>>(h1, act, (int64_t)M * ff, m.drop(m.cfg->feed_forward_dropout_p, half ? 6 : 0));
    TK_LAUNCHED(m.c);
    COCR_TRY(gemm_nt(m.c, act, ff, m.P + o.ffn_w2[half], ff, m.P + o.ffn_b2[half], y, d, M, d, ff, 0, "train_ffn_down"));
    k_resid_drop<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(x_in, y, x_out, dm.ffn_factor, (int64_t)M * d,
                                                              m.drop(m.cfg->feed_forward_dropout_p, half ? 7 : 1));
    TK_LAUNCHED(m.c);
    return COCR_OK;
}

// dx (M,d): gradient w.r.t. the module output on entry, w.r.t. its input on return
int ffn_backward(Mod &m, int half, const float *x_in, float *dx) {
    const Dims &dm = m.dm;
    const TrainLayerOff &o = m.c.h->train->L[m.l];
    const int M = m.p.M, d = dm.d, ff = dm.ff;
    float *dy = m.f(m.p.t_d0), *xn = m.f(m.p.t_d1), *dxn = m.f(m.p.t_d2), *act = m.f(m.p.t_ff0), *dact = m.f(m.p.t_ff1);
    const float *h1 = m.f(m.p.h1) + ((size_t)m.l * 2 + half) * M * ff;
    const float *st = m.stats(half ? 3 : 0);
    k_resid_drop<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(nullptr, dx, dy, dm.ffn_factor, (int64_t)M * d,
                                                              m.drop(m.cfg->feed_forward_dropout_p, half ? 7 : 1));
    TK_LAUNCHED(m.c);
    const Drop d0 = m.drop(m.cfg->feed_forward_dropout_p, half ? 6 : 0);
    k_silu_drop<<<ew_grid((int64_t)M * ff), 256, 0, m.c.st>>>(h1, act, (int64_t)M * ff, d0);
    TK_LAUNCHED(m.c);
    COCR_TRY(gemm_tn(m.c, dy, d, act, ff, m.G + o.ffn_w2[half], M, d, ff));
    COCR_TRY(colsum(m.c, dy, d, M, d, m.G + o.ffn_b2[half]));
    COCR_TRY(gemm_nn(m.c, dy, d, m.P + o.ffn_w2[half], ff, dact, ff, M, d, ff));
    k_silu_drop_bwd<<<ew_grid((int64_t)M * ff), 256, 0, m.c.st>>>(h1, dact, (int64_t)M * ff, d0);
    TK_LAUNCHED(m.c);
    COCR_TRY(ln_fwd(m.c, x_in, m.P + o.ffn_ln_g[half], m.P + o.ffn_ln_b[half], xn, nullptr, M, d));
    COCR_TRY(gemm_tn(m.c, dact, ff, xn, d, m.G + o.ffn_w1[half], M, ff, d));
    COCR_TRY(colsum(m.c, dact, ff, M, ff, m.G + o.ffn_b1[half]));
    COCR_TRY(gemm_nn(m.c, dact, ff, m.P + o.ffn_w1[half], d, dxn, d, M, ff, d));
    COCR_TRY(ln_bwd(m.c, dxn, x_in, st, m.P + o.ffn_ln_g[half], dx, dx, m.G + o.ffn_ln_g[half], m.G + o.ffn_ln_b[half], M, d));
    return COCR_OK;
}

int att_forward_mod(Mod &m, const float *x_in, float *x_out) {
    const Dims &dm = m.dm;
    TrainState *s = m.c.h->train;
    const TrainLayerOff &o = s->L[m.l];
    const int M = m.p.M, d = dm.d, B = m.p.B, T = m.p.T, R = m.p.R;
    float *xn = m.f(m.p.t_d0), *y = m.f(m.p.t_d1);
    float *qkv = m.f(m.p.qkv) + (size_t)m.l * M * 3 * d, *pos = m.f(m.p.pos) + (size_t)m.l * R * d;
    float *Pm = m.f(m.p.P) + (size_t)m.l * B * dm.nh * T * T, *ctx = m.f(m.p.ctx) + (size_t)m.l * M * d;
    COCR_TRY(ln_fwd(m.c, x_in, m.P + o.att_ln_g, m.P + o.att_ln_b, xn, m.stats(1), M, d));
    COCR_TRY(gemm_nt(m.c, xn, d, m.P + o.w_qkv, d, m.P + o.b_qkv, qkv, 3 * d, M, 3 * d, d, 0, "train_qkv"));
    COCR_TRY(gemm_nt(m.c, s->pe, d, m.P + o.w_pos, d, nullptr, pos, d, R, d, d, 0, "train_pos_proj"));
    const Drop dr = m.drop(m.cfg->attention_dropout_p, 2);
    COCR_TRY(DISPATCH_DH(dm.dh, att_forward<36>(m.c, qkv, pos, m.P + o.u_bias, m.P + o.v_bias, Pm, ctx, B, T, d, dm.nh, dm.dh, dr),
                         att_forward<16>(m.c, qkv, pos, m.P + o.u_bias, m.P + o.v_bias, Pm, ctx, B, T, d, dm.nh, dm.dh, dr),
                         att_forward<64>(m.c, qkv, pos, m.P + o.u_bias, m.P + o.v_bias, Pm, ctx, B, T, d, dm.nh, dm.dh, dr)));
    COCR_TRY(gemm_nt(m.c, ctx, d, m.P + o.w_out, d, m.P + o.b_out, y, d, M, d, d, 0, "train_att_out"));
    k_resid_drop<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(x_in, y, x_out, 1.0f, (int64_t)M * d, m.drop(m.cfg->attention_dropout_p, 3));
    TK_LAUNCHED(m.c);
    return COCR_OK;
}

int att_backward_mod(Mod &m, const float *x_in, float *dx) {
    const Dims &dm = m.dm;
    TrainState *s = m.c.h->train;
    const TrainLayerOff &o = s->L[m.l];
    const int M = m.p.M, d = dm.d, B = m.p.B, T = m.p.T, R = m.p.R;
    float *dy = m.f(m.p.t_d0), *dctx = m.f(m.p.t_d1), *dxn = m.f(m.p.t_d2), *dqkv = m.f(m.p.t_qkv), *dS = m.f(m.p.dS), *dpos = m.f(m.p.dpos);
    const float *qkv = m.f(m.p.qkv) + (size_t)m.l * M * 3 * d, *pos = m.f(m.p.pos) + (size_t)m.l * R * d;
    const float *Pm = m.f(m.p.P) + (size_t)m.l * B * dm.nh * T * T, *ctx = m.f(m.p.ctx) + (size_t)m.l * M * d;
    k_resid_drop<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(nullptr, dx, dy, 1.0f, (int64_t)M * d, m.drop(m.cfg->attention_dropout_p, 3));
    TK_LAUNCHED(m.c);
    COCR_TRY(gemm_tn(m.c, dy, d, ctx, d, m.G + o.w_out, M, d, d));
    COCR_TRY(colsum(m.c, dy, d, M, d, m.G + o.b_out));
    COCR_TRY(gemm_nn(m.c, dy, d, m.P + o.w_out, d, dctx, d, M, d, d));
    const Drop dr = m.drop(m.cfg->attention_dropout_p, 2);
    COCR_TRY(DISPATCH_DH(dm.dh,
        att_backward<36>(m.c, qkv, pos, m.P + o.u_bias, m.P + o.v_bias, Pm, ctx, dctx, dS, dqkv, dpos, m.G + o.u_bias, m.G + o.v_bias, B, T, d, dm.nh, dm.dh, dr),
        att_backward<16>(m.c, qkv, pos, m.P + o.u_bias, m.P + o.v_bias, Pm, ctx, dctx, dS, dqkv, dpos, m.G + o.u_bias, m.G + o.v_bias, B, T, d, dm.nh, dm.dh, dr),
        att_backward<64>(m.c, qkv, pos, m.P + o.u_bias, m.P + o.v_bias, Pm, ctx, dctx, dS, dqkv, dpos, m.G + o.u_bias, m.G + o.v_bias, B, T, d, dm.nh, dm.dh, dr)));
    COCR_TRY(gemm_tn(m.c, dpos, d, s->pe, d, m.G + o.w_pos, R, d, d));
    float *xn = dy;                                  // dy is dead from here on
    COCR_TRY(ln_fwd(m.c, x_in, m.P + o.att_ln_g, m.P + o.att_ln_b, xn, nullptr, M, d));
    COCR_TRY(gemm_tn(m.c, dqkv, 3 * d, xn, d, m.G + o.w_qkv, M, 3 * d, d));
    COCR_TRY(colsum(m.c, dqkv, 3 * d, M, 3 * d, m.G + o.b_qkv));
    COCR_TRY(gemm_nn(m.c, dqkv, 3 * d, m.P + o.w_qkv, d, dxn, d, M, 3 * d, d));
    COCR_TRY(ln_bwd(m.c, dxn, x_in, m.stats(1), m.P + o.att_ln_g, dx, dx, m.G + o.att_ln_g, m.G + o.att_ln_b, M, d));
    return COCR_OK;
}

int conv_forward_mod(Mod &m, const float *x_in, float *x_out, float *params_rw) {
    const Dims &dm = m.dm;
    const TrainLayerOff &o = m.c.h->train->L[m.l];
    const int M = m.p.M, d = dm.d, B = m.p.B, T = m.p.T;
    float *xn = m.f(m.p.t_d0), *y = m.f(m.p.t_d1), *g = m.f(m.p.t_d2), *z = m.f(m.p.t_ff0);
    float *u = m.f(m.p.u) + (size_t)m.l * M * 2 * d, *cd = m.f(m.p.cdw) + (size_t)m.l * M * d, *bst = m.f(m.p.bnstat) + (size_t)m.l * 2 * d;
    double *sums = (double *)m.f(m.p.bnsums);
    COCR_TRY(ln_fwd(m.c, x_in, m.P + o.conv_ln_g, m.P + o.conv_ln_b, xn, m.stats(2), M, d));
    COCR_TRY(gemm_nt(m.c, xn, d, m.P + o.w_pw1, d, m.P + o.b_pw1, u, 2 * d, M, 2 * d, d, 0, "train_pw1"));
    k_glu_fwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(u, g, M, d);
    TK_LAUNCHED(m.c);
    COCR_TRY(dwconv(m.c, g, m.P + o.w_dw, cd, B, T, d, dm.K, 0));
    COCR_CUDA(m.c.h, cudaMemsetAsync(sums, 0, (size_t)2 * d * 8, m.c.st));
    const int rpb = 32;
    k_bn_sums<<<dim3((M + rpb - 1) / rpb, (d + 255) / 256), 256, 0, m.c.st>>>(cd, M, d, rpb, sums);
    TK_LAUNCHED(m.c);
    k_bn_finalize<<<(d + 255) / 256, 256, 0, m.c.st>>>(sums, M, d, m.cfg->bn_momentum, bst, params_rw ? params_rw + o.bn_rm : nullptr,
                                                       params_rw ? params_rw + o.bn_rv : nullptr);
    TK_LAUNCHED(m.c);
    k_bn_silu_fwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(cd, bst, m.P + o.bn_g, m.P + o.bn_b, z, M, d);
    TK_LAUNCHED(m.c);
    COCR_TRY(gemm_nt(m.c, z, d, m.P + o.w_pw2, d, m.P + o.b_pw2, y, d, M, d, d, 0, "train_pw2"));
    k_resid_drop<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(x_in, y, x_out, 1.0f, (int64_t)M * d, m.drop(m.cfg->conv_dropout_p, 4));
    TK_LAUNCHED(m.c);
    return COCR_OK;
}

int conv_backward_mod(Mod &m, const float *x_in, float *dx) {
    const Dims &dm = m.dm;
    const TrainLayerOff &o = m.c.h->train->L[m.l];
    const int M = m.p.M, d = dm.d, B = m.p.B, T = m.p.T;
    float *dy = m.f(m.p.t_d0), *z = m.f(m.p.t_d1), *dz = m.f(m.p.t_d2), *du = m.f(m.p.t_ff0), *g = m.f(m.p.t_ff1);
    const float *u = m.f(m.p.u) + (size_t)m.l * M * 2 * d, *cd = m.f(m.p.cdw) + (size_t)m.l * M * d, *bst = m.f(m.p.bnstat) + (size_t)m.l * 2 * d;
    double *sums = (double *)m.f(m.p.bnsums);
    k_resid_drop<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(nullptr, dx, dy, 1.0f, (int64_t)M * d, m.drop(m.cfg->conv_dropout_p, 4));
    TK_LAUNCHED(m.c);
    k_bn_silu_fwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(cd, bst, m.P + o.bn_g, m.P + o.bn_b, z, M, d);
    TK_LAUNCHED(m.c);
    COCR_TRY(gemm_tn(m.c, dy, d, z, d, m.G + o.w_pw2, M, d, d));
    COCR_TRY(colsum(m.c, dy, d, M, d, m.G + o.b_pw2));
    COCR_TRY(gemm_nn(m.c, dy, d, m.P + o.w_pw2, d, dz, d, M, d, d));
    // BatchNorm backward (batch statistics): dz -> dy_bn -> dc, in place
    COCR_CUDA(m.c.h, cudaMemsetAsync(sums, 0, (size_t)2 * d * 8, m.c.st));
    const int rpb = 32;
    k_bn_bwd_reduce<<<dim3((M + rpb - 1) / rpb, (d + 255) / 256), 256, 0, m.c.st>>>(cd, bst, m.P + o.bn_g, m.P + o.bn_b, dz, M, d, rpb, sums);
    TK_LAUNCHED(m.c);
    k_bn_bwd_apply<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(cd, bst, m.P + o.bn_g, sums, dz, M, d, m.G + o.bn_g, m.G + o.bn_b);
    TK_LAUNCHED(m.c);
    float *dc = dz;
    // depthwise conv: weight gradient needs its input g = GLU(u) again
    k_glu_fwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(u, g, M, d);
    TK_LAUNCHED(m.c);
    COCR_TRY(dwconv_wgrad(m.c, g, dc, m.G + o.w_dw, B, T, d, dm.K));
    float *dg = dy;                                   // dy is dead
    COCR_TRY(dwconv(m.c, dc, m.P + o.w_dw, dg, B, T, d, dm.K, 1));
    k_glu_bwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(u, dg, du, M, d);
    TK_LAUNCHED(m.c);
    float *xn = z;                                    // z is dead
    COCR_TRY(ln_fwd(m.c, x_in, m.P + o.conv_ln_g, m.P + o.conv_ln_b, xn, nullptr, M, d));
    COCR_TRY(gemm_tn(m.c, du, 2 * d, xn, d, m.G + o.w_pw1, M, 2 * d, d));
    COCR_TRY(colsum(m.c, du, 2 * d, M, 2 * d, m.G + o.b_pw1));
    float *dxn = dz;                                  // dc is dead
    COCR_TRY(gemm_nn(m.c, du, 2 * d, m.P + o.w_pw1, d, dxn, d, M, 2 * d, d));
    COCR_TRY(ln_bwd(m.c, dxn, x_in, m.stats(2), m.P + o.conv_ln_g, dx, dx, m.G + o.conv_ln_g, m.G + o.conv_ln_b, M, d));
    return COCR_OK;
}

#include "train_tc.cuh"

}  // namespace

void train_destroy(cocr_handle *h) {
    if (!h->train) return;
    if (h->train->ws) cudaFree(h->train->ws);
    for (auto &kv : h->train->pes) cudaFree(kv.second);
    if (h->train->seed_dev) cudaFree(h->train->seed_dev);
    for (auto &g : h->train->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->train->gstream) cudaStreamDestroy(h->train->gstream);
    if (h->train->side) cudaStreamDestroy(h->train->side);
    if (h->train->ev_fork) cudaEventDestroy(h->train->ev_fork);
    if (h->train->ev_join) cudaEventDestroy(h->train->ev_join);
    if (h->train->g_in) cudaEventDestroy(h->train->g_in);
    if (h->train->g_out) cudaEventDestroy(h->train->g_out);
    delete h->train;
    h->train = nullptr;
}

int train_layout(cocr_handle *h, int64_t *n_total, int64_t *n_trainable, int32_t *n_entries) {
    TrainState *s;
    COCR_TRY(train_state(h, &s));
    if (n_total) *n_total = (int64_t)s->n_total;
    if (n_trainable) *n_trainable = (int64_t)s->n_trainable;
    if (n_entries) *n_entries = (int32_t)s->entries.size();
    return COCR_OK;
}

int train_entry(cocr_handle *h, int32_t i, char *key, size_t cap, int64_t *off, int64_t *numel, int32_t *trainable) {
    TrainState *s;
    COCR_TRY(train_state(h, &s));
    if (i < 0 || (size_t)i >= s->entries.size()) return set_error(h, COCR_ERR_BAD_ARG, "cocr_train_entry: index %d out of range", i);
    const TrainEntry &e = s->entries[i];
    if (key && cap) {
        if (e.key.size() + 1 > cap) return set_error(h, COCR_ERR_BAD_ARG, "cocr_train_entry: key buffer too small");
        memcpy(key, e.key.c_str(), e.key.size() + 1);
    }
    if (off) *off = (int64_t)e.off;
    if (numel) *numel = (int64_t)e.numel;
    if (trainable) *trainable = e.trainable ? 1 : 0;
    return COCR_OK;
}

int train_workspace_bytes(cocr_handle *h, int B, int W, size_t *bytes) {
    TrainState *s;
    COCR_TRY(train_state(h, &s));
    Plan p;
    make_plan(h->dm, B, W, 1, s->n_total, &p);
    *bytes = p.total;
    return COCR_OK;
}

int dropout_mask(cocr_handle *h, uint64_t seed, uint32_t site, float prob, int64_t n, cudaStream_t st, float *out) {
    TrainState *s;
    COCR_TRY(train_state(h, &s));
    cocr_train_config cfg = {};
    k_set_seed<<<1, 1, 0, st>>>(s->seed_dev + 1, seed);
    COCR_LAUNCHED(h);
    g_seedp = s->seed_dev + 1;
    k_dropout_mask<<<ew_grid(n), 256, 0, st>>>(make_drop(&cfg, prob, site), n, out);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

int adamw_step(cocr_handle *h, float *p, const float *g, float *m, float *v, int64_t n, float lr, float b1, float b2, float eps, float wd, int step,
               cudaStream_t st) {
    if (n <= 0) return COCR_OK;
    const float bc1 = (float)(1.0 - pow((double)b1, (double)step)), bc2s = (float)sqrt(1.0 - pow((double)b2, (double)step));
    ProfScope ps(h, "train_adamw", st, 0.0, 28.0 * n);
    k_adamw<<<ew_grid(n), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2s);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

static int train_step_body(cocr_handle *h, TrainState *s, const Plan &p, float *params, float *grads, const float *img, int64_t sb, int64_t sw,
                           int64_t sh, const int32_t *widths, int B, int W, const int32_t *targets, const int32_t *target_lens, int max_target_len,
                           const cocr_train_config *cfg, cudaStream_t st, float *loss_out, float *loss_per_line, float *taps);

int train_step(cocr_handle *h, float *params, float *grads, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int B, int W,
               const int32_t *targets, const int32_t *target_lens, int max_target_len, const cocr_train_config *cfg, cudaStream_t st, float *loss_out,
               float *loss_per_line, float *taps) {
    TrainState *s;
    COCR_TRY(train_state(h, &s));
    const Dims &dm = h->dm;
    Plan p;
    const int tc = cfg->compute == 1 ? 1 : 0;
    make_plan(dm, B, W, tc, s->n_total, &p);
    COCR_TRY(ensure_train_ws(h, s, p.total));
    COCR_TRY(ensure_pe(h, s, p.T, st));
    // the seed lives in device memory: the captured step is replayed with a new one
    k_set_seed<<<1, 1, 0, st>>>(s->seed_dev, cfg->seed);
    COCR_LAUNCHED(h);
    g_seedp = s->seed_dev;
    static const bool no_graph = getenv("COCR_NO_GRAPH") != nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (taps || h->profiling || no_graph || cap != cudaStreamCaptureStatusNone)
        return train_step_body(h, s, p, params, grads, img, sb, sw, sh, widths, B, W, targets, target_lens, max_target_len, cfg, st, loss_out,
                               loss_per_line, taps);
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return (uint64_t)u; };
    std::vector<uint64_t> key = {(uint64_t)B, (uint64_t)W, (uint64_t)max_target_len, (uint64_t)tc, (uint64_t)params, (uint64_t)grads, (uint64_t)img,
                                 (uint64_t)sb, (uint64_t)sw, (uint64_t)sh, (uint64_t)widths, (uint64_t)targets, (uint64_t)target_lens, (uint64_t)loss_out,
                                 (uint64_t)loss_per_line, (uint64_t)s->ws, (uint64_t)s->pe, bits(cfg->input_dropout_p), bits(cfg->feed_forward_dropout_p),
                                 bits(cfg->attention_dropout_p), bits(cfg->conv_dropout_p), bits(cfg->bn_momentum)};
    TrainState::StepGraph *sg = nullptr;
    for (auto &g : s->graphs)
        if (g.key == key) { sg = &g; break; }
    if (!sg) {
        if (s->graphs.size() >= 8) {
            if (s->graphs.front().exec) cudaGraphExecDestroy(s->graphs.front().exec);
            s->graphs.erase(s->graphs.begin());
        }
        s->graphs.emplace_back();
        sg = &s->graphs.back();
        sg->key = key;
    }
    if (!s->gstream) {
        COCR_CUDA(h, cudaStreamCreateWithFlags(&s->gstream, cudaStreamNonBlocking));
        COCR_CUDA(h, cudaEventCreateWithFlags(&s->g_in, cudaEventDisableTiming));
        COCR_CUDA(h, cudaEventCreateWithFlags(&s->g_out, cudaEventDisableTiming));
    }
    if (!sg->exec) {
        if (++sg->sightings < 2)       // first call with these arguments: plain launches (also settles every lazily-set function attribute)
            return train_step_body(h, s, p, params, grads, img, sb, sw, sh, widths, B, W, targets, target_lens, max_target_len, cfg, st, loss_out,
                                   loss_per_line, nullptr);
        const int64_t l0 = h->launches;
        COCR_CUDA(h, cudaStreamBeginCapture(s->gstream, cudaStreamCaptureModeThreadLocal));
        const int rc = train_step_body(h, s, p, params, grads, img, sb, sw, sh, widths, B, W, targets, target_lens, max_target_len, cfg, s->gstream,
                                       loss_out, loss_per_line, nullptr);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(s->gstream, &graph);
        if (rc != COCR_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc != COCR_OK) return rc;
            return set_error(h, COCR_ERR_CUDA, "capture of the training step failed: %s", cudaGetErrorString(e));
        }
        sg->launches = h->launches - l0;
        h->launches = l0;
        const cudaError_t ei = cudaGraphInstantiate(&sg->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) {
            sg->exec = nullptr;
            return set_error(h, COCR_ERR_CUDA, "cudaGraphInstantiate (training step): %s", cudaGetErrorString(ei));
        }
    }
    // replay on the handle's own stream (the caller's may be the legacy default stream, which cannot be captured), fenced by two events
    COCR_CUDA(h, cudaEventRecord(s->g_in, st));
    COCR_CUDA(h, cudaStreamWaitEvent(s->gstream, s->g_in, 0));
    COCR_CUDA(h, cudaGraphLaunch(sg->exec, s->gstream));
    COCR_CUDA(h, cudaEventRecord(s->g_out, s->gstream));
    COCR_CUDA(h, cudaStreamWaitEvent(st, s->g_out, 0));
    h->launches += sg->launches;
    return COCR_OK;
}

static int train_step_body(cocr_handle *h, TrainState *s, const Plan &p, float *params, float *grads, const float *img, int64_t sb, int64_t sw,
                           int64_t sh, const int32_t *widths, int B, int W, const int32_t *targets, const int32_t *target_lens, int max_target_len,
                           const cocr_train_config *cfg, cudaStream_t st, float *loss_out, float *loss_per_line, float *taps) {
    const Dims &dm = h->dm;
    const int tc = p.tc;
    char *ws = s->ws;
    Ctx c{h, st, (float *)(ws + p.part), p.part_floats};
    static const bool no_side = getenv("COCR_TRAIN_NO_SIDE_STREAM") != nullptr;
    if (tc && !no_side && !h->profiling) {
        if (!s->side) {
            COCR_CUDA(h, cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
            COCR_CUDA(h, cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
            COCR_CUDA(h, cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
        }
        c.side = s->side; c.ev_fork = s->ev_fork; c.ev_join = s->ev_join;
    }
    const int M = p.M, d = dm.d, L = dm.L, C = dm.C;
    const float *P = params;
    float *G = grads;
    auto F = [&](size_t off) { return (float *)(ws + off); };
    auto xs = [&](int idx) { return F(p.xs) + (size_t)idx * M * d; };
    COCR_CUDA(h, cudaMemsetAsync(G, 0, s->n_total * 4, st));
    int32_t *lens = (int32_t *)(ws + p.lens);
    COCR_TRY(launch_out_lens(h, widths, B, W, lens, st));

    if (tc) {     // bf16 mirror of the parameters (same offsets) and of the positional table: the operands of every GEMM of this step
        ProfScope ps(h, "train_tc_weight_cast", st, 0.0, 6.0 * s->n_total);
        k_cast_bf16<<<ew_grid((int64_t)s->n_total), 256, 0, st>>>(P, (int64_t)s->n_total, (__nv_bfloat16 *)(ws + p.pb), (int64_t)s->n_total, 1, (int)s->n_total);
        COCR_LAUNCHED(h);
        k_cast_bf16<<<ew_grid((int64_t)p.R * d), 256, 0, st>>>(s->pe, d, (__nv_bfloat16 *)(ws + p.peb), d, p.R, d);
        COCR_LAUNCHED(h);
    }
    // ---------------- forward ----------------
    float *a1 = F(p.a1), *a2 = F(p.a2), *a3 = F(p.a3), *wperm = F(p.wperm);
    {
        const int64_t n1 = (int64_t)B * p.T2 * dm.F2 * dm.cc, n2 = (int64_t)M * dm.feat;
        ProfScope ps(h, "train_frontend", st, 18.0 * (n1 + n2), 4.0 * ((double)B * W * dm.H + 2.0 * n1 + n2));
        k_fe_conv1<<<ew_grid(n1), 256, 0, st>>>(img, sb, sw, sh, P + s->sub_w1, P + s->sub_b1, a1, B, W, dm.H, p.T2, dm.F2, dm.cc);
        COCR_LAUNCHED(h);
        k_fe_dw<<<ew_grid(n2), 256, 0, st>>>(a1, P + s->sub_wdw, P + s->sub_bdw, a2, B, p.T2, dm.F2, p.T, dm.F, dm.cc);
        COCR_LAUNCHED(h);
    }
    COCR_TRY(gemm_nt(c, a2, dm.cc, P + s->sub_wpw, dm.cc, P + s->sub_bpw, a3, dm.cc, M * dm.F, dm.cc, dm.cc, 2, "train_fe_pw"));
    k_permute_wout<<<ew_grid((int64_t)d * dm.feat), 256, 0, st>>>(P + s->sub_wout, wperm, d, dm.cc, dm.F, 0);
    COCR_LAUNCHED(h);
    {
        float *y = F(p.t_d1);
        if (tc) {
            COCR_TRY(tc_cast(c, a3, dm.feat, (__nv_bfloat16 *)(ws + p.a3b), dm.feat, M, dm.feat));
            COCR_TRY(tc_cast(c, wperm, dm.feat, (__nv_bfloat16 *)(ws + p.wpermb), dm.feat, d, dm.feat));
            COCR_TRY(tc_linear(c, (__nv_bfloat16 *)(ws + p.a3b), dm.feat, (__nv_bfloat16 *)(ws + p.wpermb), P + s->sub_bout, y, d, M, d, dm.feat, "train_tc_fe_out"));
        } else {
            COCR_TRY(gemm_nt(c, a3, dm.feat, wperm, dm.feat, P + s->sub_bout, y, d, M, d, dm.feat, 0, "train_fe_out"));
        }
        k_resid_drop<<<ew_grid((int64_t)M * d), 256, 0, st>>>(nullptr, y, xs(0), 1.0f, (int64_t)M * d, make_drop(cfg, cfg->input_dropout_p, 0));
        COCR_LAUNCHED(h);
    }
    for (int l = 0; l < L; ++l) {
        Mod m{c, dm, p, ws, P, G, cfg, l};
        const TrainLayerOff &o = s->L[l];
        const int x0 = 5 * l;
        if (tc) {
            COCR_TRY(ffn_forward_tc(m, 0, xs(x0), xs(x0 + 1)));
            COCR_TRY(att_forward_tc(m, xs(x0 + 1), xs(x0 + 2)));
            COCR_TRY(conv_forward_tc(m, xs(x0 + 2), xs(x0 + 3), params));
            COCR_TRY(ffn_forward_tc(m, 1, xs(x0 + 3), xs(x0 + 4)));
        } else {
            COCR_TRY(ffn_forward(m, 0, xs(x0), xs(x0 + 1)));
            COCR_TRY(att_forward_mod(m, xs(x0 + 1), xs(x0 + 2)));
            COCR_TRY(conv_forward_mod(m, xs(x0 + 2), xs(x0 + 3), params));
            COCR_TRY(ffn_forward(m, 1, xs(x0 + 3), xs(x0 + 4)));
        }
        COCR_TRY(ln_fwd(c, xs(x0 + 4), P + o.fin_g, P + o.fin_b, xs(x0 + 5), m.stats(4), M, d));
    }
    const float *enc = xs(5 * L);
    float *logits = F(p.logits), *glog = F(p.glogits);
    TcBuf tb = {};
    if (tc) {
        tb = tc_buffers(ws, p);
        COCR_TRY(tc_cast(c, enc, d, tb.encb, d, M, d));
        COCR_TRY(tc_linear(c, tb.encb, d, tb.pb + s->dec_w, P + s->dec_b, logits, C, M, C, d, "train_tc_decoder"));
    } else {
        COCR_TRY(gemm_nt(c, enc, d, P + s->dec_w, d, P + s->dec_b, logits, C, M, C, d, 0, "train_decoder"));
    }
    COCR_TRY(launch_ctc_loss(h, logits, targets, target_lens, lens, B, p.T, C, max_target_len, st, loss_out, loss_per_line, glog));
    if (taps) COCR_CUDA(h, cudaMemcpyAsync(taps, xs(0), (size_t)(1 + 5 * L) * M * d * 4, cudaMemcpyDeviceToDevice, st));

    // ---------------- backward ----------------
    float *dx = F(p.dx);
    float *gtaps = taps ? taps + (size_t)(1 + 5 * L) * M * d : nullptr;
    auto tap = [&](int idx) -> int {
        if (gtaps) COCR_CUDA(h, cudaMemcpyAsync(gtaps + (size_t)idx * M * d, dx, (size_t)M * d * 4, cudaMemcpyDeviceToDevice, st));
        return COCR_OK;
    };
    if (tc) {
        COCR_TRY(tc_cast(c, glog, C, tb.glogb, p.Cp, M, C));
        COCR_TRY(c.fork());
        COCR_TRY(tc_wgrad(c, tb.glogb, p.Cp, tb.encb, d, G + s->dec_w, M, C, d));
        COCR_TRY(tc_dgrad(c, tb.glogb, p.Cp, tb.pb + s->dec_w, dx, d, M, C, d));
    } else {
        COCR_TRY(gemm_tn(c, glog, C, enc, d, G + s->dec_w, M, C, d));
        COCR_TRY(gemm_nn(c, glog, C, P + s->dec_w, d, dx, d, M, C, d));
    }
    COCR_TRY(colsum(c, glog, C, M, C, G + s->dec_b));
    COCR_TRY(tap(5 * L));
    for (int l = L - 1; l >= 0; --l) {
        Mod m{c, dm, p, ws, P, G, cfg, l};
        const TrainLayerOff &o = s->L[l];
        const int x0 = 5 * l;
        COCR_TRY(ln_bwd(c, dx, xs(x0 + 4), m.stats(4), P + o.fin_g, nullptr, dx, G + o.fin_g, G + o.fin_b, M, d));
        COCR_TRY(tap(x0 + 4));
        COCR_TRY(tc ? ffn_backward_tc(m, 1, xs(x0 + 3), dx) : ffn_backward(m, 1, xs(x0 + 3), dx));
        COCR_TRY(tap(x0 + 3));
        COCR_TRY(tc ? conv_backward_tc(m, xs(x0 + 2), dx) : conv_backward_mod(m, xs(x0 + 2), dx));
        COCR_TRY(tap(x0 + 2));
        COCR_TRY(tc ? att_backward_tc(m, xs(x0 + 1), dx) : att_backward_mod(m, xs(x0 + 1), dx));
        COCR_TRY(tap(x0 + 1));
        COCR_TRY(tc ? ffn_backward_tc(m, 0, xs(x0), dx) : ffn_backward(m, 0, xs(x0), dx));
        COCR_TRY(tap(x0));
    }
    // front-end (its weight gradients are consumed right away: main stream only)
    COCR_TRY(c.join());
    c.side = nullptr;
    {
        float *dy = F(p.t_d0), *gwp = F(p.gwperm), *da3 = F(p.t_ff0), *da2 = F(p.t_ff1);
        k_resid_drop<<<ew_grid((int64_t)M * d), 256, 0, st>>>(nullptr, dx, dy, 1.0f, (int64_t)M * d, make_drop(cfg, cfg->input_dropout_p, 0),
                                                              tc ? tb.dyb : nullptr);
        COCR_LAUNCHED(h);
        if (tc) COCR_TRY(tc_wgrad(c, tb.dyb, d, (__nv_bfloat16 *)(ws + p.a3b), dm.feat, gwp, M, d, dm.feat));
        else COCR_TRY(gemm_tn(c, dy, d, a3, dm.feat, gwp, M, d, dm.feat));
        k_permute_wout<<<ew_grid((int64_t)d * dm.feat), 256, 0, st>>>(gwp, G + s->sub_wout, d, dm.cc, dm.F, 1);
        COCR_LAUNCHED(h);
        COCR_TRY(colsum(c, dy, d, M, d, G + s->sub_bout));
        if (tc) COCR_TRY(tc_dgrad(c, tb.dyb, d, (__nv_bfloat16 *)(ws + p.wpermb), da3, dm.feat, M, d, dm.feat));
        else COCR_TRY(gemm_nn(c, dy, d, wperm, dm.feat, da3, dm.feat, M, d, dm.feat));
        const int64_t n2 = (int64_t)M * dm.feat, n1 = (int64_t)B * p.T2 * dm.F2 * dm.cc;
        k_relu_bwd<<<ew_grid(n2), 256, 0, st>>>(a3, da3, n2);
        COCR_LAUNCHED(h);
        const int rows = M * dm.F;
        COCR_TRY(gemm_tn(c, da3, dm.cc, a2, dm.cc, G + s->sub_wpw, rows, dm.cc, dm.cc));
        COCR_TRY(colsum(c, da3, dm.cc, rows, dm.cc, G + s->sub_bpw));
        COCR_TRY(gemm_nn(c, da3, dm.cc, P + s->sub_wpw, dm.cc, da2, dm.cc, rows, dm.cc, dm.cc));
        ProfScope ps(h, "train_frontend_bwd", st, 54.0 * (n1 + n2), 4.0 * (3.0 * n1 + 2.0 * n2 + (double)B * W * dm.H));
        k_fe_dw_wgrad<<<148 * 4, 256, 0, st>>>(a1, da2, G + s->sub_wdw, G + s->sub_bdw, B, p.T2, dm.F2, p.T, dm.F, dm.cc);
        COCR_LAUNCHED(h);
        k_fe_dw_dgrad_relu<<<ew_grid(n1), 256, 0, st>>>(a1, da2, P + s->sub_wdw, B, p.T2, dm.F2, p.T, dm.F, dm.cc);
        COCR_LAUNCHED(h);
        k_fe_conv1_wgrad<<<148 * 4, 256, 0, st>>>(img, sb, sw, sh, a1, G + s->sub_w1, G + s->sub_b1, B, W, dm.H, p.T2, dm.F2, dm.cc);
        COCR_LAUNCHED(h);
    }
    return COCR_OK;
}

}  // namespace cocr
