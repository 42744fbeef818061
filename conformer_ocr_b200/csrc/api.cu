// C ABI of libcocr_b200 (include/cocr_b200.h): handle life cycle, weight ingestion / folding,
// workspace management and dispatch to the fp32 or bf16 device path.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "common.h"

namespace cocr {

static std::mutex g_err_mu;
static std::string g_err = "";

void set_global_error(const char *msg) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = msg;
}

int set_error(cocr_handle *h, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else set_global_error(buf);
    return code;
}

int check_cuda(cocr_handle *h, cudaError_t e, const char *what) {
    if (e == cudaSuccess) return COCR_OK;
    int code = (e == cudaErrorMemoryAllocation) ? COCR_ERR_OOM : COCR_ERR_CUDA;
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();   // clear the sticky-less OOM so later calls work (model.py:146-151 skips the batch)
        return set_error(h, code, "CUDA out of memory. (%s)", what);
    }
    return set_error(h, code, "CUDA error %s: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e), what);
}

int ensure_workspace(cocr_handle *h, size_t bytes) {
    if (bytes <= h->ws.bytes) return COCR_OK;
    if (h->ws.base) {
        COCR_CUDA(h, cudaDeviceSynchronize());
        COCR_CUDA(h, cudaFree(h->ws.base));
        h->ws.base = nullptr;
        h->ws.bytes = 0;
    }
    void *p = nullptr;
    COCR_CUDA(h, cudaMalloc(&p, bytes));
    h->ws.base = (char *)p;
    h->ws.bytes = bytes;
    return COCR_OK;
}

// Relative positional table (embedding.py:35-67) rows k=0..2T-2 <-> offset r=T-1-k, computed on the
// host in the reference's own fp32 steps, then projected per layer with pos_proj (attention.py:85).
int ensure_pos_proj(cocr_handle *h, int T, cudaStream_t st) {
    if (h->pos_T == T && h->pos_proj) return COCR_OK;
    const Dims &dm = h->dm;
    const int R = 2 * T - 1, d = dm.d;
    if (h->pos_proj) {
        COCR_CUDA(h, cudaDeviceSynchronize());
        COCR_CUDA(h, cudaFree(h->pos_proj));
        h->pos_proj = nullptr;
        h->pos_T = 0;
    }
    void *p = nullptr;
    COCR_CUDA(h, cudaMalloc(&p, sizeof(float) * (size_t)(dm.L + 1) * R * d));
    h->pos_proj = (float *)p;
    float *pe_dev = h->pos_proj + (size_t)dm.L * R * d;
    std::vector<float> pe((size_t)R * d);
    const float neg = (float)(-(log(10000.0) / (double)d));
    for (int k = 0; k < R; ++k) {
        float r = (float)(T - 1 - k);
        for (int m = 0; m < d; m += 2) {
            float div = expf((float)m * neg);
            float arg = r * div;
            pe[(size_t)k * d + m] = sinf(arg);
            if (m + 1 < d) pe[(size_t)k * d + m + 1] = cosf(arg);
        }
    }
    COCR_CUDA(h, cudaMemcpyAsync(pe_dev, pe.data(), pe.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    for (int l = 0; l < dm.L; ++l)
        COCR_TRY(launch_sgemm(h, pe_dev, d, h->layers[l].w_pos, d, nullptr, h->pos_proj + (size_t)l * R * d, d, R, d, d, 0,
                              nullptr, 0, 0.f, st));
    COCR_CUDA(h, cudaStreamSynchronize(st));   // `pe` is pageable host memory
    h->pos_T = T;
    return COCR_OK;
}

static void add_expect(cocr_handle *h, const std::string &k, std::vector<int64_t> shape) { h->expect[k] = std::move(shape); }

static void build_expected(cocr_handle *h) {
    const Dims &dm = h->dm;
    const int64_t d = dm.d, ff = dm.ff, cc = dm.cc, K = dm.K, nh = dm.nh, dh = dm.dh;
    const std::string s = "encoder.conv_subsample.";
    add_expect(h, s + "conv.0.weight", {cc, 1, 3, 3});
    add_expect(h, s + "conv.0.bias", {cc});
    add_expect(h, s + "conv.2.weight", {cc, 1, 3, 3});
    add_expect(h, s + "conv.2.bias", {cc});
    add_expect(h, s + "conv.3.weight", {cc, cc, 1, 1});
    add_expect(h, s + "conv.3.bias", {cc});
    add_expect(h, s + "out.0.weight", {d, (int64_t)dm.feat});
    add_expect(h, s + "out.0.bias", {d});
    for (int l = 0; l < dm.L; ++l) {
        const std::string p = "encoder.layers." + std::to_string(l) + ".sequential.";
        for (const char *f : {"0", "3"}) {
            const std::string q = p + f + ".module.sequential.";
            add_expect(h, q + "0.weight", {d});
            add_expect(h, q + "0.bias", {d});
            add_expect(h, q + "1.linear.weight", {ff, d});
            add_expect(h, q + "1.linear.bias", {ff});
            add_expect(h, q + "4.linear.weight", {d, ff});
            add_expect(h, q + "4.linear.bias", {d});
        }
        std::string q = p + "1.module.";
        add_expect(h, q + "layer_norm.weight", {d});
        add_expect(h, q + "layer_norm.bias", {d});
        add_expect(h, q + "attention.u_bias", {nh, dh});
        add_expect(h, q + "attention.v_bias", {nh, dh});
        for (const char *nm : {"query", "key", "value", "out"}) {
            add_expect(h, q + "attention." + nm + "_proj.linear.weight", {d, d});
            add_expect(h, q + "attention." + nm + "_proj.linear.bias", {d});
        }
        add_expect(h, q + "attention.pos_proj.linear.weight", {d, d});
        q = p + "2.module.sequential.";
        add_expect(h, q + "0.weight", {d});
        add_expect(h, q + "0.bias", {d});
        add_expect(h, q + "2.conv.weight", {2 * d, d, 1});
        add_expect(h, q + "2.conv.bias", {2 * d});
        add_expect(h, q + "4.conv.weight", {d, 1, K});
        add_expect(h, q + "5.weight", {d});
        add_expect(h, q + "5.bias", {d});
        add_expect(h, q + "5.running_mean", {d});
        add_expect(h, q + "5.running_var", {d});
        add_expect(h, q + "7.conv.weight", {d, d, 1});
        add_expect(h, q + "7.conv.bias", {d});
        add_expect(h, p + "4.weight", {d});
        add_expect(h, p + "4.bias", {d});
    }
    if (dm.C > 0) {
        add_expect(h, "decoder.weight", {(int64_t)dm.C, d});
        add_expect(h, "decoder.bias", {(int64_t)dm.C});
    }
}

}  // namespace cocr

using namespace cocr;

extern "C" {

const char *cocr_version(void) { return "cocr_b200 0.1 (sm_100a; fp32 CUDA-core path + bf16 tcgen05 path)"; }

const char *cocr_last_error(const cocr_handle *h) {
    if (h) return h->err.c_str();
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_err_mu);
    copy = g_err;
    return copy.c_str();
}

int32_t cocr_out_frames(int32_t width) {
    float l = (float)width;
    l = floorf((l - 1.0f) / 2.0f + 1.0f);
    l = floorf((l - 1.0f) / 2.0f + 1.0f);
    return (int32_t)l;
}

int cocr_create(const cocr_hparams *hp, int device, cocr_handle **out) {
    if (!hp || !out) return set_error(nullptr, COCR_ERR_BAD_ARG, "cocr_create: null argument");
    *out = nullptr;
    // the reference's own constructor checks
    if (hp->encoder_dim <= 0 || hp->num_heads <= 0 || hp->encoder_dim % hp->num_heads != 0)
        return set_error(nullptr, COCR_ERR_BAD_ARG, "d_model %% num_heads should be zero.");                    // attention.py:53
    if (hp->conv_kernel_size < 1 || (hp->conv_kernel_size - 1) % 2 != 0)
        return set_error(nullptr, COCR_ERR_BAD_ARG, "kernel_size should be a odd number for 'SAME' padding");    // convolution.py:132
    if (hp->conv_expansion != 2)
        return set_error(nullptr, COCR_ERR_BAD_ARG, "Currently, Only Supports expansion_factor 2");              // convolution.py:133
    if (hp->subsampling_factor <= 0 || (hp->subsampling_factor & (hp->subsampling_factor - 1)) != 0)
        return set_error(nullptr, COCR_ERR_BAD_ARG, "Sampling factor should be a power of 2.");                  // convolution.py:180-181
    if (hp->subsampling_factor != 4)
        return set_error(nullptr, COCR_ERR_UNSUPPORTED, "only subsampling_factor 4 is implemented (got %d)", hp->subsampling_factor);
    if (hp->height < 1 || hp->num_layers < 0 || hp->ff_expansion < 1 || hp->subsampling_conv_channels < 1 || hp->num_classes < 0)
        return set_error(nullptr, COCR_ERR_BAD_ARG, "cocr_create: non-positive hyper-parameter");
    if (hp->precision != COCR_FP32 && hp->precision != COCR_BF16)
        return set_error(nullptr, COCR_ERR_BAD_ARG, "cocr_create: unknown precision %d", hp->precision);
    if (hp->mask_mode != COCR_MASK_REFERENCE_UNMASKED && hp->mask_mode != COCR_MASK_RAGGED)
        return set_error(nullptr, COCR_ERR_BAD_ARG, "cocr_create: unknown mask mode %d", hp->mask_mode);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev)
        return set_error(nullptr, COCR_ERR_CUDA, "cocr_create: CUDA device %d not available (%s)", device,
                         e == cudaSuccess ? "index out of range" : cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return set_error(nullptr, COCR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return set_error(nullptr, COCR_ERR_UNSUPPORTED, "libcocr_b200 is built for sm_100a only; device %d is sm_%d%d", device,
                         prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return set_error(nullptr, COCR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

    cocr_handle *h = new cocr_handle();
    h->hp = *hp;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    Dims &dm = h->dm;
    dm.H = hp->height; dm.d = hp->encoder_dim; dm.L = hp->num_layers; dm.nh = hp->num_heads;
    dm.dh = dm.d / dm.nh; dm.ff = dm.d * hp->ff_expansion; dm.K = hp->conv_kernel_size;
    dm.cc = hp->subsampling_conv_channels; dm.C = hp->num_classes;
    dm.F2 = out_len1(dm.H); dm.F = out_len1(dm.F2); dm.feat = dm.cc * dm.F;
    dm.ffn_factor = hp->half_step_residual ? 0.5f : 1.0f;
    build_expected(h);
    *out = h;
    return COCR_OK;
}

int cocr_destroy(cocr_handle *h) {
    if (!h) return COCR_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    bf16_destroy(h);
    train_destroy(h);
    if (h->arena) cudaFree(h->arena);
    if (h->ws.base) cudaFree(h->ws.base);
    if (h->pos_proj) cudaFree(h->pos_proj);
    for (auto &sl : h->slots) {
        if (sl.img) cudaFree(sl.img);
        if (sl.widths) cudaFree(sl.widths);
        if (sl.res) cudaFree(sl.res);
        if (sl.res_host) cudaFreeHost(sl.res_host);
        if (sl.h2d_done) cudaEventDestroy(sl.h2d_done);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.fwd_done) cudaEventDestroy(sl.fwd_done);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    delete h;
    return COCR_OK;
}

int cocr_set_weight(cocr_handle *h, const char *key, const float *data, int ndim, const int64_t *shape) {
    if (!h || !key || !data || ndim < 0 || (ndim > 0 && !shape)) return set_error(h, COCR_ERR_BAD_ARG, "cocr_set_weight: null argument");
    std::string k(key);
    if (k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return COCR_OK;
    auto it = h->expect.find(k);
    if (it == h->expect.end()) return set_error(h, COCR_ERR_BAD_ARG, "unexpected key in state_dict: %s", key);
    std::vector<int64_t> shp(shape, shape + ndim);
    if (shp != it->second) {
        std::string a, b;
        for (auto v : shp) a += std::to_string(v) + ",";
        for (auto v : it->second) b += std::to_string(v) + ",";
        return set_error(h, COCR_ERR_BAD_ARG, "size mismatch for %s: got (%s) expected (%s)", key, a.c_str(), b.c_str());
    }
    HostTensor t;
    t.shape = shp;
    t.data.assign(data, data + t.numel());
    h->host_w[k] = std::move(t);
    h->finalized = false;
    return COCR_OK;
}

int cocr_finalize_weights(cocr_handle *h) {
    if (!h) return COCR_ERR_BAD_ARG;
    COCR_CUDA(h, cudaSetDevice(h->device));
    for (auto &kv : h->expect)
        if (!h->host_w.count(kv.first)) return set_error(h, COCR_ERR_STATE, "missing key in state_dict: %s", kv.first.c_str());
    const Dims &dm = h->dm;
    const int d = dm.d;
    std::vector<float> arena;
    auto push = [&](const float *p, size_t n) {
        size_t off = arena.size();
        arena.insert(arena.end(), p, p + n);
        arena.resize((arena.size() + 63) / 64 * 64, 0.0f);
        return off;
    };
    auto W = [&](const std::string &k) -> const HostTensor & { return h->host_w.at(k); };
    auto pushk = [&](const std::string &k) { const HostTensor &t = W(k); return push(t.data.data(), t.data.size()); };

    struct Off { size_t v[40]; };
    const std::string s = "encoder.conv_subsample.";
    size_t o_w1 = pushk(s + "conv.0.weight"), o_b1 = pushk(s + "conv.0.bias");
    size_t o_wdw = pushk(s + "conv.2.weight"), o_bdw = pushk(s + "conv.2.bias");
    size_t o_wpw = pushk(s + "conv.3.weight"), o_bpw = pushk(s + "conv.3.bias");
    // out.0.weight (d, cc*F) column c*F+f  ->  column f*cc+c (channel-last activations)
    std::vector<float> wout((size_t)d * dm.feat);
    {
        const std::vector<float> &src = W(s + "out.0.weight").data;
        for (int n = 0; n < d; ++n)
            for (int c = 0; c < dm.cc; ++c)
                for (int f = 0; f < dm.F; ++f) wout[(size_t)n * dm.feat + f * dm.cc + c] = src[(size_t)n * dm.feat + c * dm.F + f];
    }
    size_t o_wout = push(wout.data(), wout.size()), o_bout = pushk(s + "out.0.bias");
    std::vector<Off> lo(dm.L);
    for (int l = 0; l < dm.L; ++l) {
        const std::string p = "encoder.layers." + std::to_string(l) + ".sequential.";
        size_t *v = lo[l].v;
        int n = 0;
        for (const char *f : {"0", "3"}) {
            const std::string q = p + f + ".module.sequential.";
            v[n++] = pushk(q + "0.weight"); v[n++] = pushk(q + "0.bias");
            v[n++] = pushk(q + "1.linear.weight"); v[n++] = pushk(q + "1.linear.bias");
            v[n++] = pushk(q + "4.linear.weight"); v[n++] = pushk(q + "4.linear.bias");
        }
        std::string q = p + "1.module.";
        v[n++] = pushk(q + "layer_norm.weight"); v[n++] = pushk(q + "layer_norm.bias");
        std::vector<float> wqkv, bqkv;
        for (const char *nm : {"query", "key", "value"}) {
            const auto &w = W(q + "attention." + nm + "_proj.linear.weight").data;
            const auto &b = W(q + "attention." + nm + "_proj.linear.bias").data;
            wqkv.insert(wqkv.end(), w.begin(), w.end());
            bqkv.insert(bqkv.end(), b.begin(), b.end());
        }
        v[n++] = push(wqkv.data(), wqkv.size()); v[n++] = push(bqkv.data(), bqkv.size());
        v[n++] = pushk(q + "attention.pos_proj.linear.weight");
        v[n++] = pushk(q + "attention.u_bias"); v[n++] = pushk(q + "attention.v_bias");
        v[n++] = pushk(q + "attention.out_proj.linear.weight"); v[n++] = pushk(q + "attention.out_proj.linear.bias");
        q = p + "2.module.sequential.";
        v[n++] = pushk(q + "0.weight"); v[n++] = pushk(q + "0.bias");
        v[n++] = pushk(q + "2.conv.weight"); v[n++] = pushk(q + "2.conv.bias");
        v[n++] = pushk(q + "4.conv.weight");
        // BatchNorm1d in eval mode folded to y = x*scale + shift (convolution.py:141; eps 1e-5)
        std::vector<float> sc(d), sh(d);
        {
            const auto &g = W(q + "5.weight").data, &b = W(q + "5.bias").data;
            const auto &mu = W(q + "5.running_mean").data, &var = W(q + "5.running_var").data;
            for (int c = 0; c < d; ++c) {
                float r = 1.0f / sqrtf(var[c] + 1e-5f);
                sc[c] = g[c] * r;
                sh[c] = b[c] - mu[c] * sc[c];
            }
        }
        v[n++] = push(sc.data(), d); v[n++] = push(sh.data(), d);
        v[n++] = pushk(q + "7.conv.weight"); v[n++] = pushk(q + "7.conv.bias");
        v[n++] = pushk(p + "4.weight"); v[n++] = pushk(p + "4.bias");
    }
    size_t o_dw = 0, o_db = 0;
    if (dm.C > 0) { o_dw = pushk("decoder.weight"); o_db = pushk("decoder.bias"); }

    if (h->arena) { COCR_CUDA(h, cudaDeviceSynchronize()); COCR_CUDA(h, cudaFree(h->arena)); h->arena = nullptr; }
    void *p = nullptr;
    COCR_CUDA(h, cudaMalloc(&p, arena.size() * sizeof(float)));
    h->arena = (float *)p;
    h->arena_floats = arena.size();
    COCR_CUDA(h, cudaMemcpy(h->arena, arena.data(), arena.size() * sizeof(float), cudaMemcpyHostToDevice));
    const float *A = h->arena;
    h->sub_w1 = A + o_w1; h->sub_b1 = A + o_b1; h->sub_wdw = A + o_wdw; h->sub_bdw = A + o_bdw;
    h->sub_wpw = A + o_wpw; h->sub_bpw = A + o_bpw; h->sub_wout = A + o_wout; h->sub_bout = A + o_bout;
    h->dec_w = dm.C > 0 ? A + o_dw : nullptr; h->dec_b = dm.C > 0 ? A + o_db : nullptr;
    h->layers.assign(dm.L, LayerF32());
    for (int l = 0; l < dm.L; ++l) {
        LayerF32 &w = h->layers[l];
        const size_t *v = lo[l].v;
        int n = 0;
        for (int f = 0; f < 2; ++f) {
            w.ffn_ln_g[f] = A + v[n++]; w.ffn_ln_b[f] = A + v[n++];
            w.ffn_w1[f] = A + v[n++]; w.ffn_b1[f] = A + v[n++];
            w.ffn_w2[f] = A + v[n++]; w.ffn_b2[f] = A + v[n++];
        }
        w.att_ln_g = A + v[n++]; w.att_ln_b = A + v[n++];
        w.w_qkv = A + v[n++]; w.b_qkv = A + v[n++];
        w.w_pos = A + v[n++];
        w.u_bias = A + v[n++]; w.v_bias = A + v[n++];
        w.w_out = A + v[n++]; w.b_out = A + v[n++];
        w.conv_ln_g = A + v[n++]; w.conv_ln_b = A + v[n++];
        w.w_pw1 = A + v[n++]; w.b_pw1 = A + v[n++];
        w.w_dw = A + v[n++];
        w.bn_scale = A + v[n++]; w.bn_shift = A + v[n++];
        w.w_pw2 = A + v[n++]; w.b_pw2 = A + v[n++];
        w.final_ln_g = A + v[n++]; w.final_ln_b = A + v[n++];
    }
    // conv.0 / conv.2 of the front-end in kernel-parameter order (w1[32*9], b1[32], wdw[32*9], bdw[32])
    h->fe_weights.clear();
    if (dm.cc == 32) {
        for (const char *k : {"conv.0.weight", "conv.0.bias", "conv.2.weight", "conv.2.bias"}) {
            const std::vector<float> &v = W(s + k).data;
            h->fe_weights.insert(h->fe_weights.end(), v.begin(), v.end());
        }
    }
    h->pos_T = 0;   // positional projections depend on the weights
    if (h->hp.precision == COCR_BF16) COCR_TRY(bf16_init(h));
    h->finalized = true;
    return COCR_OK;
}

static int check_call(cocr_handle *h, const void *img, int B, int W) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!h->finalized) return set_error(h, COCR_ERR_STATE, "weights not finalised (call cocr_finalize_weights)");
    if (!img || B < 1 || W < 1) return set_error(h, COCR_ERR_BAD_ARG, "bad batch: img=%p B=%d W=%d", img, B, W);
    if ((int64_t)B * cocr_out_frames(W) >= (int64_t)1 << 31) return set_error(h, COCR_ERR_UNSUPPORTED, "batch too large");
    COCR_CUDA(h, cudaSetDevice(h->device));
    return COCR_OK;
}

int cocr_encode(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int32_t B,
                int32_t W, void *stream, float *enc_out, int32_t *out_lens) {
    COCR_TRY(check_call(h, img, B, W));
    if (!enc_out) return set_error(h, COCR_ERR_BAD_ARG, "cocr_encode: enc_out is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->hp.precision == COCR_BF16) return bf16_encode(h, img, sb, sw, sh, widths, B, W, st, enc_out, out_lens);
    return fp32_encode(h, img, sb, sw, sh, widths, B, W, st, enc_out, out_lens);
}

int cocr_logits(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int32_t B,
                int32_t W, void *stream, float *logits, int32_t *out_lens) {
    COCR_TRY(check_call(h, img, B, W));
    if (!logits) return set_error(h, COCR_ERR_BAD_ARG, "cocr_logits: logits is null");
    if (h->dm.C <= 0) return set_error(h, COCR_ERR_STATE, "handle was created without a decoder (num_classes = 0)");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->hp.precision == COCR_BF16) return bf16_logits(h, img, sb, sw, sh, widths, B, W, st, logits, out_lens);
    COCR_TRY(fp32_encode(h, img, sb, sw, sh, widths, B, W, st, nullptr, out_lens));
    Fp32Plan p;
    fp32_plan(h, B, W, &p);
    return fp32_logits(h, (const float *)(h->ws.base + p.off_x), p.M, st, logits);
}

int cocr_greedy_decode(cocr_handle *h, const float *logits, const int32_t *lens, int32_t B, int32_t T, int32_t C,
                       void *stream, int32_t *labels, int32_t *starts, int32_t *ends, float *conf, int32_t *counts) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!logits || !labels || !starts || !ends || !conf || !counts || B < 1 || T < 0 || C < 1)
        return set_error(h, COCR_ERR_BAD_ARG, "cocr_greedy_decode: bad argument");
    COCR_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) { COCR_CUDA(h, cudaMemsetAsync(counts, 0, sizeof(int32_t) * B, st)); return COCR_OK; }
    // scratch for (value, index) per frame lives behind the forward workspace so both can coexist
    size_t need = (size_t)B * T * 8 + 512;
    static_assert(sizeof(float) == 4, "");
    void *scratch = nullptr;
    COCR_CUDA(h, cudaMallocAsync(&scratch, need, st));
    float *bv = (float *)scratch;
    int32_t *bi = (int32_t *)((char *)scratch + (((size_t)B * T * 4 + 255) / 256) * 256);
    int rc = launch_greedy_decode(h, logits, lens, B, T, C, st, bv, bi, labels, starts, ends, conf, counts);
    cudaFreeAsync(scratch, st);
    return rc;
}

static int forward_decode_any(cocr_handle *h, const void *img, bool u8, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths,
                              int32_t B, int32_t W, void *stream, int32_t *labels, int32_t *starts, int32_t *ends, float *conf,
                              int32_t *counts, int32_t *out_lens) {
    COCR_TRY(check_call(h, img, B, W));
    if (!labels || !starts || !ends || !conf || !counts || !out_lens)
        return set_error(h, COCR_ERR_BAD_ARG, "cocr_forward_decode: null output");
    if (h->dm.C <= 0) return set_error(h, COCR_ERR_STATE, "handle was created without a decoder (num_classes = 0)");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->hp.precision == COCR_BF16)
        return bf16_forward_decode(h, img, u8, sb, sw, sh, widths, B, W, st, labels, starts, ends, conf, counts, out_lens);
    if (u8) return set_error(h, COCR_ERR_UNSUPPORTED, "uint8 input is implemented on the bf16 path only");
    // fp32: logits go through the `h` scratch tile by tile would save memory; they fit the workspace for the
    // parity-sized batches this path serves, so project into a dedicated block behind the plan.
    const float *imgf = static_cast<const float *>(img);
    Fp32Plan p;
    fp32_plan(h, B, W, &p);
    size_t logits_off = p.total;
    size_t total = logits_off + (size_t)p.M * h->dm.C * sizeof(float);
    COCR_TRY(ensure_workspace(h, total));
    COCR_TRY(fp32_encode(h, imgf, sb, sw, sh, widths, B, W, st, nullptr, out_lens));
    float *logits = (float *)(h->ws.base + logits_off);
    COCR_TRY(fp32_logits(h, (const float *)(h->ws.base + p.off_x), p.M, st, logits));
    return launch_greedy_decode(h, logits, out_lens, B, p.T, h->dm.C, st, (float *)(h->ws.base + p.off_best_val),
                                (int32_t *)(h->ws.base + p.off_best_idx), labels, starts, ends, conf, counts);
}

int cocr_forward_decode(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths,
                        int32_t B, int32_t W, void *stream, int32_t *labels, int32_t *starts, int32_t *ends, float *conf,
                        int32_t *counts, int32_t *out_lens) {
    return forward_decode_any(h, img, false, sb, sw, sh, widths, B, W, stream, labels, starts, ends, conf, counts, out_lens);
}

int cocr_forward_decode_u8(cocr_handle *h, const uint8_t *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths,
                           int32_t B, int32_t W, void *stream, int32_t *labels, int32_t *starts, int32_t *ends, float *conf,
                           int32_t *counts, int32_t *out_lens) {
    if (h && h->hp.precision != COCR_BF16) return set_error(h, COCR_ERR_UNSUPPORTED, "uint8 input is implemented on the bf16 path only");
    return forward_decode_any(h, img, true, sb, sw, sh, widths, B, W, stream, labels, starts, ends, conf, counts, out_lens);
}

static int host_begin(cocr_handle *h, int32_t slot, const void *img_host, size_t pix_bytes, const int32_t *widths_host, int32_t B, int32_t W,
                      void *stream) {
    COCR_TRY(check_call(h, img_host, B, W));
    if (slot < 0 || slot > 1) return set_error(h, COCR_ERR_BAD_ARG, "cocr_recognize_host_begin: slot must be 0 or 1");
    cocr_handle::HostSlot &sl = h->slots[slot];
    if (sl.busy) return set_error(h, COCR_ERR_STATE, "cocr_recognize_host_begin: slot %d still holds an unfinished batch", slot);
    cudaStream_t st = (cudaStream_t)stream;
    const int H = h->dm.H;
    const int T = cocr_out_frames(W);
    if (!h->copy_stream) COCR_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->d2h_stream) COCR_CUDA(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    if (!sl.h2d_done) {
        COCR_CUDA(h, cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
        COCR_CUDA(h, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        COCR_CUDA(h, cudaEventCreateWithFlags(&sl.fwd_done, cudaEventDisableTiming));
    }
    const size_t img_bytes = pix_bytes * (size_t)B * H * W;
    if (img_bytes > sl.img_bytes) {
        if (sl.img) { COCR_CUDA(h, cudaDeviceSynchronize()); COCR_CUDA(h, cudaFree(sl.img)); sl.img = nullptr; sl.img_bytes = 0; }
        void *p = nullptr;
        COCR_CUDA(h, cudaMalloc(&p, img_bytes));
        sl.img = p;
        sl.img_bytes = img_bytes;
    }
    if ((size_t)B > sl.widths_n) {
        if (sl.widths) { COCR_CUDA(h, cudaDeviceSynchronize()); COCR_CUDA(h, cudaFree(sl.widths)); sl.widths = nullptr; }
        void *p = nullptr;
        COCR_CUDA(h, cudaMalloc(&p, sizeof(int32_t) * B));
        sl.widths = (int32_t *)p;
        sl.widths_n = B;
    }
    const size_t bt = (size_t)B * T;
    const size_t res_bytes = 4 * (4 * bt + 2 * (size_t)B);
    if (res_bytes > sl.res_bytes) {
        if (sl.res) { COCR_CUDA(h, cudaDeviceSynchronize()); COCR_CUDA(h, cudaFree(sl.res)); COCR_CUDA(h, cudaFreeHost(sl.res_host)); }
        sl.res = nullptr; sl.res_host = nullptr; sl.res_bytes = 0;
        void *p = nullptr;
        COCR_CUDA(h, cudaMalloc(&p, res_bytes));
        sl.res = (char *)p;
        COCR_CUDA(h, cudaMallocHost(&p, res_bytes));
        sl.res_host = (char *)p;
        sl.res_bytes = res_bytes;
    }
    // H2D on the copy stream: it may run while the previous batch (other slot) is still computing on `stream`.  The
    // slot's own previous forward pass finished before its results were collected (cocr_recognize_host_end), so the
    // image buffer is free.
    COCR_CUDA(h, cudaMemcpyAsync(sl.img, img_host, img_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    const int32_t *wd = nullptr;
    if (widths_host) {
        COCR_CUDA(h, cudaMemcpyAsync(sl.widths, widths_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, h->copy_stream));
        wd = sl.widths;
    }
    COCR_CUDA(h, cudaEventRecord(sl.h2d_done, h->copy_stream));
    COCR_CUDA(h, cudaStreamWaitEvent(st, sl.h2d_done, 0));
    int32_t *d_lab = (int32_t *)sl.res, *d_st = d_lab + bt, *d_en = d_st + bt;
    float *d_cf = (float *)(d_en + bt);
    int32_t *d_cnt = (int32_t *)(d_cf + bt), *d_len = d_cnt + B;
    COCR_TRY(forward_decode_any(h, sl.img, pix_bytes == 1, (int64_t)H * W, 1, W, wd, B, W, stream, d_lab, d_st, d_en, d_cf, d_cnt, d_len));
    // the records leave on their own stream: the forward pass of the other slot, already queued on `stream`, starts at once
    COCR_CUDA(h, cudaEventRecord(sl.fwd_done, st));
    COCR_CUDA(h, cudaStreamWaitEvent(h->d2h_stream, sl.fwd_done, 0));
    COCR_CUDA(h, cudaMemcpyAsync(sl.res_host, sl.res, res_bytes, cudaMemcpyDeviceToHost, h->d2h_stream));
    COCR_CUDA(h, cudaEventRecord(sl.done, h->d2h_stream));
    sl.B = B; sl.T = T; sl.busy = true;
    return COCR_OK;
}

int cocr_recognize_host_begin(cocr_handle *h, int32_t slot, const float *img_host, const int32_t *widths_host, int32_t B, int32_t W,
                              void *stream) {
    return host_begin(h, slot, img_host, sizeof(float), widths_host, B, W, stream);
}

int cocr_recognize_host_u8_begin(cocr_handle *h, int32_t slot, const uint8_t *img_host, const int32_t *widths_host, int32_t B, int32_t W,
                                 void *stream) {
    if (h && h->hp.precision != COCR_BF16) return set_error(h, COCR_ERR_UNSUPPORTED, "uint8 input is implemented on the bf16 path only");
    return host_begin(h, slot, img_host, 1, widths_host, B, W, stream);
}

int cocr_recognize_host_end(cocr_handle *h, int32_t slot, int32_t *labels, int32_t *starts, int32_t *ends, float *conf, int32_t *counts,
                            int32_t *out_lens) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (slot < 0 || slot > 1) return set_error(h, COCR_ERR_BAD_ARG, "cocr_recognize_host_end: slot must be 0 or 1");
    if (!labels || !starts || !ends || !conf || !counts || !out_lens) return set_error(h, COCR_ERR_BAD_ARG, "cocr_recognize_host_end: null output");
    cocr_handle::HostSlot &sl = h->slots[slot];
    if (!sl.busy) return set_error(h, COCR_ERR_STATE, "cocr_recognize_host_end: slot %d holds no batch", slot);
    COCR_CUDA(h, cudaSetDevice(h->device));
    COCR_CUDA(h, cudaEventSynchronize(sl.done));
    sl.busy = false;
    const size_t bt = (size_t)sl.B * sl.T, B = (size_t)sl.B;
    const char *r = sl.res_host;
    memcpy(labels, r, 4 * bt); r += 4 * bt;
    memcpy(starts, r, 4 * bt); r += 4 * bt;
    memcpy(ends, r, 4 * bt); r += 4 * bt;
    memcpy(conf, r, 4 * bt); r += 4 * bt;
    memcpy(counts, r, 4 * B); r += 4 * B;
    memcpy(out_lens, r, 4 * B);
    return COCR_OK;
}

int cocr_recognize_host(cocr_handle *h, const float *img_host, const int32_t *widths_host, int32_t B, int32_t W,
                        void *stream, int32_t *labels, int32_t *starts, int32_t *ends, float *conf, int32_t *counts,
                        int32_t *out_lens) {
    if (h && (!labels || !starts || !ends || !conf || !counts || !out_lens))
        return set_error(h, COCR_ERR_BAD_ARG, "cocr_recognize_host: null output");
    const int slot = (h && h->slots[0].busy) ? 1 : 0;
    COCR_TRY(cocr_recognize_host_begin(h, slot, img_host, widths_host, B, W, stream));
    return cocr_recognize_host_end(h, slot, labels, starts, ends, conf, counts, out_lens);
}

int cocr_log_softmax(cocr_handle *h, const float *logits, int64_t rows, int32_t C, void *stream, float *out) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!logits || !out || rows < 0 || C < 1) return set_error(h, COCR_ERR_BAD_ARG, "cocr_log_softmax: bad argument");
    COCR_CUDA(h, cudaSetDevice(h->device));
    return launch_log_softmax(h, logits, rows, C, (cudaStream_t)stream, out);
}

int cocr_ctc_loss(cocr_handle *h, const float *logits, const int32_t *targets, const int32_t *target_lens, const int32_t *input_lens,
                  int32_t B, int32_t T, int32_t C, int32_t max_target_len, void *stream, float *loss_sum, float *loss_per_line,
                  float *grad_logits) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!logits || !targets || !target_lens || !loss_sum || B < 1 || T < 1 || C < 2 || max_target_len < 0)
        return set_error(h, COCR_ERR_BAD_ARG, "cocr_ctc_loss: bad argument");
    if ((int64_t)B * T * (2 * (int64_t)max_target_len + 1) >= ((int64_t)1 << 31))
        return set_error(h, COCR_ERR_UNSUPPORTED, "cocr_ctc_loss: alpha table too large (B*T*(2L+1) >= 2^31)");
    COCR_CUDA(h, cudaSetDevice(h->device));
    return launch_ctc_loss(h, logits, targets, target_lens, input_lens, B, T, C, max_target_len, (cudaStream_t)stream, loss_sum,
                           loss_per_line, grad_logits);
}

int cocr_project(cocr_handle *h, const float *enc, int64_t rows, void *stream, float *logits) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!h->finalized) return set_error(h, COCR_ERR_STATE, "weights not finalised (call cocr_finalize_weights)");
    if (h->dm.C <= 0) return set_error(h, COCR_ERR_STATE, "handle was created without a decoder (num_classes = 0)");
    if (!enc || !logits || rows < 1 || rows >= ((int64_t)1 << 31)) return set_error(h, COCR_ERR_BAD_ARG, "cocr_project: bad argument");
    COCR_CUDA(h, cudaSetDevice(h->device));
    return fp32_logits(h, enc, (int)rows, (cudaStream_t)stream, logits);
}

int cocr_decoder_backward(cocr_handle *h, const float *grad_logits, const float *enc, int64_t rows, void *stream, float *grad_w,
                          float *grad_b, float *grad_enc) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!h->finalized) return set_error(h, COCR_ERR_STATE, "weights not finalised (call cocr_finalize_weights)");
    if (h->dm.C <= 0) return set_error(h, COCR_ERR_STATE, "handle was created without a decoder (num_classes = 0)");
    if (!grad_logits || !enc || !grad_w || !grad_b || rows < 1 || rows >= ((int64_t)1 << 31))
        return set_error(h, COCR_ERR_BAD_ARG, "cocr_decoder_backward: bad argument");
    COCR_CUDA(h, cudaSetDevice(h->device));
    return launch_decoder_backward(h, grad_logits, enc, (int)rows, (cudaStream_t)stream, grad_w, grad_b, grad_enc);
}

int cocr_train_layout(cocr_handle *h, int64_t *n_total, int64_t *n_trainable, int32_t *n_entries) {
    if (!h) return COCR_ERR_BAD_ARG;
    return train_layout(h, n_total, n_trainable, n_entries);
}

int cocr_train_entry(cocr_handle *h, int32_t index, char *key_buf, size_t key_cap, int64_t *offset, int64_t *numel, int32_t *trainable) {
    if (!h) return COCR_ERR_BAD_ARG;
    return train_entry(h, index, key_buf, key_cap, offset, numel, trainable);
}

int cocr_train_workspace_bytes(cocr_handle *h, int32_t B, int32_t W, size_t *bytes) {
    if (!h || !bytes || B < 1 || W < 1) return set_error(h, COCR_ERR_BAD_ARG, "cocr_train_workspace_bytes: bad argument");
    return train_workspace_bytes(h, B, W, bytes);
}

int cocr_train_step(cocr_handle *h, float *params, float *grads, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths,
                    int32_t B, int32_t W, const int32_t *targets, const int32_t *target_lens, int32_t max_target_len,
                    const cocr_train_config *cfg, void *stream, float *loss_sum, float *loss_per_line, float *taps) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!params || !grads || !img || !targets || !target_lens || !cfg || !loss_sum || B < 1 || W < 1 || max_target_len < 0)
        return set_error(h, COCR_ERR_BAD_ARG, "cocr_train_step: null or non-positive argument");
    const float ps[4] = {cfg->input_dropout_p, cfg->feed_forward_dropout_p, cfg->attention_dropout_p, cfg->conv_dropout_p};
    for (float p : ps)
        if (!(p >= 0.0f && p < 1.0f)) return set_error(h, COCR_ERR_BAD_ARG, "dropout probability has to be in [0, 1), got %g", (double)p);
    if ((int64_t)B * cocr_out_frames(W) >= (int64_t)1 << 31) return set_error(h, COCR_ERR_UNSUPPORTED, "batch too large");
    COCR_CUDA(h, cudaSetDevice(h->device));
    return train_step(h, params, grads, img, sb, sw, sh, widths, B, W, targets, target_lens, max_target_len, cfg, (cudaStream_t)stream, loss_sum,
                      loss_per_line, taps);
}

int cocr_adamw_step(cocr_handle *h, float *params, const float *grads, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int32_t step, void *stream) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!params || !grads || !m || !v || n < 0 || step < 1) return set_error(h, COCR_ERR_BAD_ARG, "cocr_adamw_step: bad argument");
    COCR_CUDA(h, cudaSetDevice(h->device));
    return adamw_step(h, params, grads, m, v, n, lr, beta1, beta2, eps, weight_decay, step, (cudaStream_t)stream);
}

int cocr_dropout_mask(cocr_handle *h, uint64_t seed, uint32_t site, float p, int64_t n, void *stream, float *out) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (!out || n < 0 || !(p >= 0.0f && p < 1.0f)) return set_error(h, COCR_ERR_BAD_ARG, "cocr_dropout_mask: bad argument");
    COCR_CUDA(h, cudaSetDevice(h->device));
    return dropout_mask(h, seed, site, p, n, (cudaStream_t)stream, out);
}

int cocr_slot_records_dev(cocr_handle *h, int32_t slot, void **records_dev, size_t *bytes, int32_t *B, int32_t *T) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (slot < 0 || slot > 1 || !records_dev || !bytes) return set_error(h, COCR_ERR_BAD_ARG, "cocr_slot_records_dev: bad argument");
    cocr_handle::HostSlot &sl = h->slots[slot];
    if (!sl.res || sl.B < 1) return set_error(h, COCR_ERR_STATE, "cocr_slot_records_dev: slot %d has not held a batch", slot);
    *records_dev = sl.res;
    *bytes = 4 * (4 * (size_t)sl.B * sl.T + 2 * (size_t)sl.B);
    if (B) *B = sl.B;
    if (T) *T = sl.T;
    return COCR_OK;
}

int cocr_workspace_bytes(cocr_handle *h, int32_t B, int32_t W, size_t *bytes) {
    if (!h || !bytes || B < 1 || W < 1) return set_error(h, COCR_ERR_BAD_ARG, "cocr_workspace_bytes: bad argument");
    if (h->hp.precision == COCR_BF16) { *bytes = bf16_workspace_bytes(h, B, W); return COCR_OK; }
    Fp32Plan p;
    fp32_plan(h, B, W, &p);
    *bytes = p.total + (size_t)p.M * h->dm.C * sizeof(float);
    return COCR_OK;
}

int cocr_kernel_plan(cocr_handle *h, char *buf, size_t cap, int32_t *fallbacks) {
    if (!h) return COCR_ERR_BAD_ARG;
    if (fallbacks) *fallbacks = 0;
    if (h->hp.precision != COCR_BF16) {
        const char *msg = "all components: CUDA-core fp32 kernels (precision fp32 is the parity path)\n";
        if (buf && cap) { strncpy(buf, msg, cap - 1); buf[cap - 1] = 0; }
        return COCR_OK;
    }
    if (!h->finalized) return set_error(h, COCR_ERR_STATE, "weights not finalised (call cocr_finalize_weights)");
    return bf16_kernel_plan(h, buf, cap, fallbacks);
}

int cocr_set_profiling(cocr_handle *h, int enabled) {
    if (!h) return COCR_ERR_BAD_ARG;
    h->profiling = enabled != 0;
    return COCR_OK;
}

int cocr_profile_report(cocr_handle *h, char *buf, size_t cap) {
    if (!h || !buf || cap == 0) return set_error(h, COCR_ERR_BAD_ARG, "cocr_profile_report: bad argument");
    COCR_CUDA(h, cudaSetDevice(h->device));
    COCR_CUDA(h, cudaDeviceSynchronize());
    struct Agg { long n = 0; double ms = 0, flops = 0, bytes = 0; };
    std::map<std::string, Agg> agg;
    for (auto &r : h->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        Agg &a = agg[r.name];
        a.n++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->prof.clear();
    std::string out;
    char line[256];
    for (auto &kv : agg) {
        snprintf(line, sizeof(line), "%s %ld %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops,
                 kv.second.bytes);
        out += line;
    }
    snprintf(buf, cap, "%s", out.c_str());
    return COCR_OK;
}

int64_t cocr_launch_count(const cocr_handle *h) { return h ? h->launches : 0; }

}  // extern "C"
