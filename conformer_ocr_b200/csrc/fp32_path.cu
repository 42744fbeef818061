// fp32 (CUDA-core) implementation of the conformer hot path.
//
// This is the parity-grade path (cocr_precision COCR_FP32): every contraction accumulates in
// fp32 with FMAs in the reference's operand order class, transcendental functions use the
// accurate libdevice versions, and the relative-position term is computed in the reference's
// own formulation (projected table + index shift) rather than the rotated form the bf16 path
// uses.  It also provides the memory-bound kernels (front-end convs, LayerNorm, depthwise
// conv module, argmax + CTC collapse) shared with the bf16 path.
//
// Reference symbols: conformer/convolution.py:151-247 (front-end), feed_forward.py:45-55,
// attention.py:72-151, convolution.py:107-148, encoder.py:67-103,183-204, pred.py:118-178.
#include <math.h>

#include "common.h"

namespace cocr {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_acc(float x) { return x * sigmoid_acc(x); }

__device__ __forceinline__ float ldf(const float *p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float *p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// calc_length (convolution.py:240-247), two stride-2 stages, float arithmetic like the reference
// ------------------------------------------------------------------------------------------
__global__ void k_out_lens(const int32_t *__restrict__ widths, int B, int W, int32_t *__restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float l = (float)(widths ? widths[b] : W);
    l = floorf((l - 1.0f) / 2.0f + 1.0f);
    l = floorf((l - 1.0f) / 2.0f + 1.0f);
    out[b] = (int32_t)l;
}

int launch_out_lens(cocr_handle *h, const int32_t *widths, int B, int W, int32_t *out_lens, cudaStream_t st) {
    k_out_lens<<<(B + 127) / 128, 128, 0, st>>>(widths, B, W, out_lens);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// front-end: Conv2d(1->cc,3x3,s2,p1)+ReLU ; depthwise Conv2d(cc,3x3,s2,p1)   (convolution.py:192-205)
// activations are channel-last: a1 (B,T2,F2,cc), a2 (B,T,F,cc)
// ------------------------------------------------------------------------------------------
__global__ void k_sub_conv1(const float *__restrict__ img, int64_t sb, int64_t sw, int64_t sh,
                            const float *__restrict__ w1, const float *__restrict__ b1, float *__restrict__ a1,
                            int B, int W, int H, int T2, int F2, int cc) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)B * T2 * F2 * cc;
    if (idx >= total) return;
    int c = (int)(idx % cc);
    int64_t r = idx / cc;
    int f2 = (int)(r % F2);
    r /= F2;
    int t2 = (int)(r % T2);
    int b = (int)(r / T2);
    float acc = b1[c];
    const float *wc = w1 + c * 9;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        int w = 2 * t2 - 1 + kh;
        if (w < 0 || w >= W) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            int y = 2 * f2 - 1 + kw;
            if (y < 0 || y >= H) continue;
            acc = fmaf(wc[kh * 3 + kw], img[b * sb + w * sw + y * sh], acc);
        }
    }
    a1[idx] = fmaxf(acc, 0.0f);
}

__global__ void k_sub_dw(const float *__restrict__ a1, const float *__restrict__ wdw, const float *__restrict__ bdw,
                         float *__restrict__ a2, int B, int T2, int F2, int T, int F, int cc) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)B * T * F * cc;
    if (idx >= total) return;
    int c = (int)(idx % cc);
    int64_t r = idx / cc;
    int f = (int)(r % F);
    r /= F;
    int t = (int)(r % T);
    int b = (int)(r / T);
    float acc = bdw[c];
    const float *wc = wdw + c * 9;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        int t2 = 2 * t - 1 + kh;
        if (t2 < 0 || t2 >= T2) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            int f2 = 2 * f - 1 + kw;
            if (f2 < 0 || f2 >= F2) continue;
            acc = fmaf(wc[kh * 3 + kw], a1[(((int64_t)b * T2 + t2) * F2 + f2) * cc + c], acc);
        }
    }
    a2[idx] = acc;
}

// ------------------------------------------------------------------------------------------
// generic fp32 GEMM  C = epi(A W^T + bias)
// ------------------------------------------------------------------------------------------
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
k_sgemm(const float *__restrict__ A, int lda, const float *__restrict__ Wt, int ldw, const float *__restrict__ bias,
        float *__restrict__ C, int ldc, int M, int N, int K, int act, const float *res, int ldres, float alpha,
        int out_bf16) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    const int lrow = tid / 4;          // 0..63
    const int lk = (tid % 4) * 4;      // 0,4,8,12
    const bool a_vec = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0);
    const bool w_vec = ((ldw & 3) == 0) && ((((uintptr_t)Wt) & 15) == 0);

    for (int k0 = 0; k0 < K; k0 += BK) {
        {   // A tile
            int64_t m = m0 + lrow;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < M) {
                const float *p = A + m * lda + k0 + lk;
                if (a_vec && k0 + lk + 3 < K) {
                    float4 q = *reinterpret_cast<const float4 *>(p);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k0 + lk + i < K) v[i] = p[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[lk + i][lrow] = v[i];
        }
        {   // W tile
            int n = n0 + lrow;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (n < N) {
                const float *p = Wt + (int64_t)n * ldw + k0 + lk;
                if (w_vec && k0 + lk + 3 < K) {
                    float4 q = *reinterpret_cast<const float4 *>(p);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k0 + lk + i < K) v[i] = p[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) Ws[lk + i][lrow] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 w = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w};
            float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.0f);
            if (act == 1) v = silu_acc(v);
            else if (act == 2) v = fmaxf(v, 0.0f);
            if (res) v = res[m * ldres + n] + alpha * v;
            if (out_bf16) reinterpret_cast<__nv_bfloat16 *>(C)[m * ldc + n] = __float2bfloat16_rn(v);
            else C[m * ldc + n] = v;
        }
    }
}

int launch_sgemm(cocr_handle *h, const float *A, int lda, const float *Wt, int ldw, const float *bias, float *Cout,
                 int ldc, int M, int N, int K, int act, const float *res, int ldres, float alpha, cudaStream_t st,
                 int out_bf16) {
    if (M <= 0 || N <= 0) return COCR_OK;
    dim3 grid((M + 63) / 64, (N + 63) / 64);
    ProfScope ps(h, "sgemm_f32", st, 2.0 * M * N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N));
    k_sgemm<64, 64, 16><<<grid, 256, 0, st>>>(A, lda, Wt, ldw, bias, Cout, ldc, M, N, K, act, res, ldres, alpha, out_bf16);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, biased variance), one warp per row
// ------------------------------------------------------------------------------------------
__global__ void k_layernorm(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ g,
                            const float *__restrict__ b, int M, int d) {
    int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *xr = x + (int64_t)row * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    float mean = warp_sum(s) / (float)d;
    float v = 0.f;
    for (int c = lane; c < d; c += 32) {
        float t = xr[c] - mean;
        v = fmaf(t, t, v);
    }
    float rstd = 1.0f / sqrtf(warp_sum(v) / (float)d + 1e-5f);
    float *yr = y + (int64_t)row * d;
    for (int c = lane; c < d; c += 32) yr[c] = (xr[c] - mean) * rstd * g[c] + b[c];
}

int launch_layernorm(cocr_handle *h, const float *x, float *y, const float *g, const float *b, int M, int d,
                     cudaStream_t st) {
    ProfScope ps(h, "layernorm_f32", st, 0.0, 8.0 * M * d);
    k_layernorm<<<(M + 7) / 8, 256, 0, st>>>(x, y, g, b, M, d);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// relative-position multi-head attention, reference formulation (attention.py:82-101):
//   score[i,j] = ((q_i+u).k_j + (q_i+v).p[T-1-i+j]) / sqrt(dh) ; softmax over j ; @ v
// one CTA = 32 queries of one (line, head); keys streamed in tiles of 64, online softmax.
// ------------------------------------------------------------------------------------------
constexpr int AQT = 32;
constexpr int AKT = 64;

template <int DH, typename TA>
__global__ void __launch_bounds__(256)
k_attn_f32(const TA *__restrict__ qkv, const float *__restrict__ pos, const float *__restrict__ ub,
           const float *__restrict__ vb, const int32_t *__restrict__ lens, TA *__restrict__ ctx, int T, int d,
           int dh, float inv_sqrt_dim) {
    extern __shared__ float sm[];
    constexpr int LD = DH + 1;
    float *ks = sm;                          // [AKT][LD]
    float *vs = ks + AKT * LD;               // [AKT][LD]
    float *ps = vs + AKT * LD;               // [AQT+AKT-1][LD]
    float *S = ps + (AQT + AKT - 1) * LD;    // [AQT][AKT+1]

    const int tid = threadIdx.x;
    const int ti = tid / 8, tj = tid % 8;
    const int i0 = blockIdx.x * AQT;
    const int hd = blockIdx.y;
    const int b = blockIdx.z;
    const int i = i0 + ti;
    const int Tk = lens ? min(max(lens[b], 0), T) : T;    // ragged mode: keys beyond the line are excluded
    const int ld = 3 * d;
    const TA *base = qkv + (int64_t)b * T * ld;

    float qu[DH], qv[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) {
        float q = (c < dh && i < T) ? ldf(base + (int64_t)i * ld + hd * dh + c) : 0.0f;
        qu[c] = (c < dh) ? q + ub[hd * dh + c] : 0.0f;
        qv[c] = (c < dh) ? q + vb[hd * dh + c] : 0.0f;
    }
    float m_run = -INFINITY, l_run = 0.0f;
    float o[(DH + 7) / 8];
#pragma unroll
    for (int k = 0; k < (DH + 7) / 8; ++k) o[k] = 0.0f;

    for (int j0 = 0; j0 < Tk; j0 += AKT) {
        __syncthreads();
        for (int e = tid; e < AKT * DH; e += 256) {
            int r = e / DH, c = e % DH;
            int j = j0 + r;
            bool ok = (j < Tk) && (c < dh);
            ks[r * LD + c] = ok ? ldf(base + (int64_t)j * ld + d + hd * dh + c) : 0.0f;
            vs[r * LD + c] = ok ? ldf(base + (int64_t)j * ld + 2 * d + hd * dh + c) : 0.0f;
        }
        const int pbase = T - 1 - (i0 + AQT - 1) + j0;    // table row of local row 0
        for (int e = tid; e < (AQT + AKT - 1) * DH; e += 256) {
            int r = e / DH, c = e % DH;
            int pr = pbase + r;
            bool ok = (pr >= 0) && (pr <= 2 * T - 2) && (c < dh);
            ps[r * LD + c] = ok ? pos[(int64_t)pr * d + hd * dh + c] : 0.0f;
        }
        __syncthreads();

        float s[AKT / 8];
        float tmax = -INFINITY;
#pragma unroll
        for (int n = 0; n < AKT / 8; ++n) {
            int jl = tj + 8 * n;
            const float *kr = ks + jl * LD;
            const float *pr = ps + (AQT - 1 - ti + jl) * LD;
            float c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int c = 0; c < DH; ++c) {
                c1 = fmaf(qu[c], kr[c], c1);
                c2 = fmaf(qv[c], pr[c], c2);
            }
            float sc = (c1 + c2) * inv_sqrt_dim;
            if (j0 + jl >= Tk) sc = -INFINITY;
            s[n] = sc;
            tmax = fmaxf(tmax, sc);
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
        float m_new = fmaxf(m_run, tmax);
        float corr = (m_run == -INFINITY) ? 0.0f : expf(m_run - m_new);
        float psum = 0.f;
#pragma unroll
        for (int n = 0; n < AKT / 8; ++n) {
            float p = (s[n] == -INFINITY) ? 0.0f : expf(s[n] - m_new);
            psum += p;
            S[ti * (AKT + 1) + tj + 8 * n] = p;
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
        l_run = l_run * corr + psum;
        m_run = m_new;
#pragma unroll
        for (int k = 0; k < (DH + 7) / 8; ++k) o[k] *= corr;
        __syncwarp();
        for (int jl = 0; jl < AKT; ++jl) {
            float p = S[ti * (AKT + 1) + jl];
            const float *vr = vs + jl * LD;
#pragma unroll
            for (int k = 0; k < (DH + 7) / 8; ++k) {
                int c = tj + 8 * k;
                if (c < DH) o[k] = fmaf(p, vr[c], o[k]);
            }
        }
    }
    if (i < T) {
        float inv = (l_run > 0.f) ? 1.0f / l_run : 0.0f;
#pragma unroll
        for (int k = 0; k < (DH + 7) / 8; ++k) {
            int c = tj + 8 * k;
            if (c < dh) stf(ctx + ((int64_t)b * T + i) * d + hd * dh + c, o[k] * inv);
        }
    }
}

template <int DH, typename TA>
static int launch_attn_t(cocr_handle *h, const TA *qkv, const float *pos, const float *ub, const float *vb,
                         const int32_t *lens, TA *ctx, int B, int T, int d, int nh, int dh, cudaStream_t st) {
    size_t smem = sizeof(float) * ((size_t)(2 * AKT + AQT + AKT - 1) * (DH + 1) + (size_t)AQT * (AKT + 1));
    static PerDevice<bool> attr_done;
    if (!attr_done.get()) {
        cudaFuncSetAttribute(k_attn_f32<DH, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done.get() = true;
    }
    dim3 grid((T + AQT - 1) / AQT, nh, B);
    ProfScope ps(h, "attention_f32", st, 6.0 * B * T * (double)T * d, 16.0 * B * T * d);
    k_attn_f32<DH, TA><<<grid, 256, smem, st>>>(qkv, pos, ub, vb, lens, ctx, T, d, dh, 1.0f / sqrtf((float)dh));
    COCR_LAUNCHED(h);
    return COCR_OK;
}

template <typename TA>
static int launch_attn_any(cocr_handle *h, const TA *qkv, const float *pos, const float *ub, const float *vb,
                           const int32_t *lens, TA *ctx, int B, int T, cudaStream_t st) {
    const Dims &dm = h->dm;
    if (dm.dh <= 16) return launch_attn_t<16, TA>(h, qkv, pos, ub, vb, lens, ctx, B, T, dm.d, dm.nh, dm.dh, st);
    if (dm.dh <= 36) return launch_attn_t<36, TA>(h, qkv, pos, ub, vb, lens, ctx, B, T, dm.d, dm.nh, dm.dh, st);
    if (dm.dh <= 64) return launch_attn_t<64, TA>(h, qkv, pos, ub, vb, lens, ctx, B, T, dm.d, dm.nh, dm.dh, st);
    return set_error(h, COCR_ERR_UNSUPPORTED, "SIMT attention supports head dim <= 64 (got %d)", dm.dh);
}
static int launch_attn(cocr_handle *h, const float *qkv, const float *pos, const float *ub, const float *vb,
                       const int32_t *lens, float *ctx, int B, int T, cudaStream_t st) {
    return launch_attn_any<float>(h, qkv, pos, ub, vb, lens, ctx, B, T, st);
}
int launch_attn_simt_bf16(cocr_handle *h, const __nv_bfloat16 *qkv, const float *pos, const float *ub, const float *vb,
                          const int32_t *lens, __nv_bfloat16 *ctx, int B, int T, cudaStream_t st) {
    return launch_attn_any<__nv_bfloat16>(h, qkv, pos, ub, vb, lens, ctx, B, T, st);
}

// ------------------------------------------------------------------------------------------
// conv module middle: GLU -> depthwise conv (K taps, zero pad at batch-line edges) -> folded BN -> SiLU
// (convolution.py:139-142).  u (B,T,2d) -> g (B,T,d)
// ------------------------------------------------------------------------------------------
constexpr int DW_TT = 32;

template <typename TA>
__global__ void __launch_bounds__(256)
k_glu_dw_bn_silu(const TA *__restrict__ u, const float *__restrict__ wdw, const float *__restrict__ scale,
                 const float *__restrict__ shift, TA *__restrict__ g, int T, int d, int K) {
    extern __shared__ float sm[];
    float *tile = sm;                                // [DW_TT + K - 1][d]
    float *wt = sm + (size_t)(DW_TT + K - 1) * d;    // [K][d]
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * DW_TT;
    const int pad = (K - 1) / 2;
    const int rows = DW_TT + K - 1;
    for (int e = threadIdx.x; e < rows * d; e += blockDim.x) {
        int r = e / d, c = e % d;
        int t = t0 - pad + r;
        float v = 0.0f;
        if (t >= 0 && t < T) {
            const TA *p = u + ((int64_t)b * T + t) * 2 * d;
            v = ldf(p + c) * sigmoid_acc(ldf(p + d + c));
        }
        tile[e] = v;
    }
    for (int e = threadIdx.x; e < K * d; e += blockDim.x) {
        int k = e / d, c = e % d;
        wt[e] = wdw[c * K + k];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < DW_TT * d; e += blockDim.x) {
        int tt = e / d, c = e % d;
        int t = t0 + tt;
        if (t >= T) break;
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) acc = fmaf(wt[k * d + c], tile[(tt + k) * d + c], acc);
        float y = acc * scale[c] + shift[c];
        stf(g + ((int64_t)b * T + t) * d + c, silu_acc(y));
    }
}

template <typename TA>
static int launch_glu_dw_any(cocr_handle *h, const TA *u, const LayerF32 &lw, TA *g, int B, int T, cudaStream_t st) {
    const Dims &dm = h->dm;
    size_t smem = sizeof(float) * (size_t)(DW_TT + 2 * dm.K - 1) * dm.d;
    if (smem > 200 * 1024) return set_error(h, COCR_ERR_UNSUPPORTED, "conv module tile needs %zu B of shared memory", smem);
    static PerDevice<size_t> attr;
    if (smem > attr.get()) {
        cudaFuncSetAttribute(k_glu_dw_bn_silu<TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr.get() = smem;
    }
    dim3 grid((T + DW_TT - 1) / DW_TT, B);
    ProfScope ps(h, "glu_dwconv_bn_silu_f32", st, 2.0 * dm.K * B * (double)T * dm.d, 12.0 * B * T * dm.d);
    k_glu_dw_bn_silu<TA><<<grid, 256, smem, st>>>(u, lw.w_dw, lw.bn_scale, lw.bn_shift, g, T, dm.d, dm.K);
    COCR_LAUNCHED(h);
    return COCR_OK;
}
static int launch_glu_dw(cocr_handle *h, const float *u, const LayerF32 &lw, float *g, int B, int T, cudaStream_t st) {
    return launch_glu_dw_any<float>(h, u, lw, g, B, T, st);
}
int launch_glu_dw_bf16(cocr_handle *h, const __nv_bfloat16 *u, const LayerF32 &lw, __nv_bfloat16 *g, int B, int T,
                       cudaStream_t st) {
    return launch_glu_dw_any<__nv_bfloat16>(h, u, lw, g, B, T, st);
}

// ------------------------------------------------------------------------------------------
// greedy CTC decode: per-frame argmax (lowest index wins ties) + per-line run collapse
// ------------------------------------------------------------------------------------------
__global__ void k_argmax(const float *__restrict__ logits, int64_t M, int C, float *__restrict__ best_val,
                         int32_t *__restrict__ best_idx) {
    int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *p = logits + row * C;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        float v = p[c];
        if (v > bv) { bv = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
        best_val[row] = bv;
        best_idx[row] = (bi == 0x7fffffff) ? 0 : bi;
    }
}

constexpr int COLLAPSE_SMEM_T = 4096;             // frames of a line staged in shared memory (32 KB); longer lines read global memory

__global__ void __launch_bounds__(256)
k_ctc_collapse(const float *__restrict__ best_val, const int32_t *__restrict__ best_idx,
               const int32_t *__restrict__ lens, int T, int32_t *__restrict__ labels, int32_t *__restrict__ starts,
               int32_t *__restrict__ ends, float *__restrict__ conf, int32_t *__restrict__ counts) {
    __shared__ int warp_tot[8];
    __shared__ int s_offset;
    extern __shared__ int32_t s_line[];                 // [n_stage] labels | [n_stage] scores (as bits)
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int len = lens ? min(max(lens[b], 0), T) : T;
    const int32_t *idx = best_idx + (int64_t)b * T;
    const float *val = best_val + (int64_t)b * T;
    // A thread that opens a run walks to its end; with few, long runs that walk is a chain of dependent loads, so the line is
    // staged in shared memory first (the walk then costs ~30 instead of ~300 cycles per frame: 45 -> 8 us per batch)
    const int n_stage = min(len, COLLAPSE_SMEM_T);
    const float *s_val = reinterpret_cast<const float *>(s_line + n_stage);
    for (int t = tid; t < n_stage; t += 256) {
        s_line[t] = idx[t];
        s_line[n_stage + t] = __float_as_int(val[t]);
    }
    if (tid == 0) s_offset = 0;
    __syncthreads();
    auto lab_at = [&](int t) { return t < n_stage ? s_line[t] : idx[t]; };
    auto val_at = [&](int t) { return t < n_stage ? s_val[t] : val[t]; };
    for (int base = 0; base < len; base += 256) {
        int t = base + tid;
        int lab = 0;
        bool emit = false;
        if (t < len) {
            lab = lab_at(t);
            bool start = (t == 0) || (lab_at(t - 1) != lab);
            emit = start && (lab != 0);
        }
        unsigned bal = __ballot_sync(0xffffffffu, emit);
        int excl = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < wid) woff += warp_tot[w];
            tot += warp_tot[w];
        }
        int pos = s_offset + woff + excl;
        // Every run opened in this warp is walked by the WHOLE warp, 32 frames per step (ballot for the first frame whose label
        // differs, shuffle maximum over the frames before it): a line that decodes to one long run costs T/32 steps instead of
        // a T-long chain of dependent loads in one thread (the synthetic benchmark lines are exactly that case).
        for (unsigned m = bal; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const int t0 = __shfl_sync(0xffffffffu, t, src), lab0 = __shfl_sync(0xffffffffu, lab, src);
            int e = t0;
            float mx = val_at(t0);
            for (int nb = t0 + 1;; nb += 32) {
                const int tt = nb + lane;
                const bool same = tt < len && lab_at(tt) == lab0;
                const unsigned sm = __ballot_sync(0xffffffffu, same);
                const int n = (sm == 0xffffffffu) ? 32 : __ffs(~sm) - 1;        // frames of this step that continue the run
                float v = (lane < n) ? val_at(tt) : -INFINITY;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                mx = fmaxf(mx, v);
                e += n;
                if (n < 32) break;
            }
            if (lane == src) {
                int64_t o = (int64_t)b * T + pos;
                labels[o] = lab;
                starts[o] = t;
                ends[o] = e;
                conf[o] = mx;
            }
        }
        __syncthreads();
        if (tid == 0) s_offset += tot;
        __syncthreads();
    }
    if (tid == 0) counts[b] = s_offset;
}

int launch_collapse(cocr_handle *h, const float *best_val, const int32_t *best_idx, const int32_t *lens, int B, int T,
                    cudaStream_t st, int32_t *labels, int32_t *starts, int32_t *ends, float *conf, int32_t *counts) {
    if (B <= 0) return COCR_OK;
    ProfScope ps(h, "ctc_collapse", st, 0.0, 24.0 * B * T);
    const size_t smem = (size_t)(T < COLLAPSE_SMEM_T ? T : COLLAPSE_SMEM_T) * 8;
    k_ctc_collapse<<<B, 256, smem, st>>>(best_val, best_idx, lens, T, labels, starts, ends, conf, counts);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

int launch_greedy_decode(cocr_handle *h, const float *logits, const int32_t *lens, int B, int T, int C,
                         cudaStream_t st, float *best_val, int32_t *best_idx, int32_t *labels, int32_t *starts,
                         int32_t *ends, float *conf, int32_t *counts) {
    int64_t M = (int64_t)B * T;
    if (M > 0) {
        ProfScope ps(h, "argmax", st, 0.0, 4.0 * M * C + 8.0 * M);
        k_argmax<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(logits, M, C, best_val, best_idx);
        COCR_LAUNCHED(h);
    }
    return launch_collapse(h, best_val, best_idx, lens, B, T, st, labels, starts, ends, conf, counts);
}

// Conv2d(1->cc)+ReLU and the depthwise Conv2d of the subsampling front-end (channel-last outputs)
int launch_frontend_convs(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, int B, int W, float *a1,
                          float *a2, cudaStream_t st) {
    const Dims &dm = h->dm;
    const int T2 = out_len1(W), T = out_len1(T2);
    int64_t n1 = (int64_t)B * T2 * dm.F2 * dm.cc;
    {
        ProfScope ps(h, "sub_conv1_f32", st, 18.0 * n1, 4.0 * ((double)B * W * dm.H + n1));
        k_sub_conv1<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(img, sb, sw, sh, h->sub_w1, h->sub_b1, a1, B, W, dm.H, T2,
                                                                    dm.F2, dm.cc);
        COCR_LAUNCHED(h);
    }
    int64_t n2 = (int64_t)B * T * dm.F * dm.cc;
    {
        ProfScope ps(h, "sub_dwconv_f32", st, 18.0 * n2, 4.0 * ((double)n1 + n2));
        k_sub_dw<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(a1, h->sub_wdw, h->sub_bdw, a2, B, T2, dm.F2, T, dm.F, dm.cc);
        COCR_LAUNCHED(h);
    }
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// forward driver
// ------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void fp32_plan(const cocr_handle *h, int B, int W, Fp32Plan *p) {
    const Dims &dm = h->dm;
    p->B = B; p->W = W;
    p->T2 = out_len1(W);
    p->T = out_len1(p->T2);
    p->M = B * p->T;
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off = align_up(off + floats * sizeof(float), 256); return o; };
    p->off_a1 = take((size_t)B * p->T2 * dm.F2 * dm.cc);
    p->off_a2 = take((size_t)p->M * dm.F * dm.cc);
    p->off_feat = take((size_t)p->M * dm.feat);
    p->off_x = take((size_t)p->M * dm.d);
    p->off_xn = take((size_t)p->M * dm.d);
    size_t wide = (size_t)dm.ff > (size_t)3 * dm.d ? dm.ff : 3 * dm.d;
    p->off_h = take((size_t)p->M * wide);
    p->off_g = take((size_t)p->M * dm.d);
    p->off_best_val = take((size_t)p->M);
    p->off_best_idx = take((size_t)p->M);
    p->total = off;
}

int fp32_encode(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int B,
                int W, cudaStream_t st, float *enc_out, int32_t *out_lens) {
    const Dims &dm = h->dm;
    Fp32Plan p;
    fp32_plan(h, B, W, &p);
    COCR_TRY(ensure_workspace(h, p.total));
    COCR_TRY(ensure_pos_proj(h, p.T, st));
    char *ws = h->ws.base;
    float *a1 = (float *)(ws + p.off_a1), *a2 = (float *)(ws + p.off_a2), *feat = (float *)(ws + p.off_feat);
    float *x = (float *)(ws + p.off_x), *xn = (float *)(ws + p.off_xn), *hb = (float *)(ws + p.off_h);
    float *g = (float *)(ws + p.off_g);
    const int M = p.M, d = dm.d;

    if (out_lens) COCR_TRY(launch_out_lens(h, widths, B, W, out_lens, st));
    // ragged mode needs the per-line frame counts on the device
    const int32_t *lens_dev = nullptr;
    if (h->hp.mask_mode == COCR_MASK_RAGGED && widths) {
        int32_t *tmp = (int32_t *)(ws + p.off_best_idx);    // free until the decode step
        COCR_TRY(launch_out_lens(h, widths, B, W, tmp, st));
        lens_dev = tmp;
    }
    {
        COCR_TRY(launch_frontend_convs(h, img, sb, sw, sh, B, W, a1, a2, st));
    }
    {
        // 1x1 conv + ReLU as a GEMM over (b,t,f) rows; output row (b,t) holds f*cc+c
        COCR_TRY(launch_sgemm(h, a2, dm.cc, h->sub_wpw, dm.cc, h->sub_bpw, feat, dm.cc, M * dm.F, dm.cc, dm.cc, 2, nullptr,
                              0, 0.f, st));
        COCR_TRY(launch_sgemm(h, feat, dm.feat, h->sub_wout, dm.feat, h->sub_bout, x, d, M, d, dm.feat, 0, nullptr, 0, 0.f,
                              st));
    }
    for (int l = 0; l < dm.L; ++l) {
        const LayerF32 &lw = h->layers[l];
        for (int half = 0; half < 2; ++half) {
            if (half == 1) {
                // --- attention module (attention.py:143-151) ---
                COCR_TRY(launch_layernorm(h, x, xn, lw.att_ln_g, lw.att_ln_b, M, d, st));
                COCR_TRY(launch_sgemm(h, xn, d, lw.w_qkv, d, lw.b_qkv, hb, 3 * d, M, 3 * d, d, 0, nullptr, 0, 0.f, st));
                COCR_TRY(launch_attn(h, hb, h->pos_proj + (size_t)l * (2 * p.T - 1) * d, lw.u_bias, lw.v_bias, lens_dev, g,
                                     B, p.T, st));
                COCR_TRY(launch_sgemm(h, g, d, lw.w_out, d, lw.b_out, x, d, M, d, d, 0, x, d, 1.0f, st));
                // --- convolution module (convolution.py:135-148) ---
                COCR_TRY(launch_layernorm(h, x, xn, lw.conv_ln_g, lw.conv_ln_b, M, d, st));
                COCR_TRY(launch_sgemm(h, xn, d, lw.w_pw1, d, lw.b_pw1, hb, 2 * d, M, 2 * d, d, 0, nullptr, 0, 0.f, st));
                COCR_TRY(launch_glu_dw(h, hb, lw, g, B, p.T, st));
                COCR_TRY(launch_sgemm(h, g, d, lw.w_pw2, d, lw.b_pw2, x, d, M, d, d, 0, x, d, 1.0f, st));
            }
            // --- feed-forward module (feed_forward.py:45-55), residual factor encoder.py:68-75 ---
            COCR_TRY(launch_layernorm(h, x, xn, lw.ffn_ln_g[half], lw.ffn_ln_b[half], M, d, st));
            COCR_TRY(launch_sgemm(h, xn, d, lw.ffn_w1[half], d, lw.ffn_b1[half], hb, dm.ff, M, dm.ff, d, 1, nullptr, 0, 0.f,
                                  st));
            COCR_TRY(launch_sgemm(h, hb, dm.ff, lw.ffn_w2[half], dm.ff, lw.ffn_b2[half], x, d, M, d, dm.ff, 0, x, d,
                                  dm.ffn_factor, st));
        }
        COCR_TRY(launch_layernorm(h, x, x, lw.final_ln_g, lw.final_ln_b, M, d, st));
    }
    if (enc_out && enc_out != x)
        COCR_CUDA(h, cudaMemcpyAsync(enc_out, x, (size_t)M * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return COCR_OK;
}

int fp32_logits(cocr_handle *h, const float *enc, int M, cudaStream_t st, float *logits) {
    const Dims &dm = h->dm;
    return launch_sgemm(h, enc, dm.d, h->dec_w, dm.d, h->dec_b, logits, dm.C, M, dm.C, dm.d, 0, nullptr, 0, 0.f, st);
}

}  // namespace cocr
