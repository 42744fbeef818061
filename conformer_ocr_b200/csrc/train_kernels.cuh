// Kernels of the training step (train.cu): fp32 CUDA-core forward in train() mode and the hand-written backward pass of
// every module on the path.  Reference symbols: model.py:129-145 (_step), encoder.py:67-103, feed_forward.py:45-55,
// attention.py:72-151, convolution.py:135-148,190-237, modules.py:31-32.  Everything here is memory- or FMA-bound fp32;
// contractions accumulate in fp32 in a fixed order, column reductions (bias / LayerNorm / BatchNorm gradients) use atomics.
#pragma once

#include <math.h>

#include "common.h"
#include "dropout.cuh"

namespace cocr {
namespace train {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void k_dropout_mask(Drop dr, int64_t n, float *out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) out[e] = drop_scale(dr, (uint64_t)e);
}

// ------------------------------------------------------------------------------------------
// generic fp32 GEMM  C[i,j] = act(sum_k A(i,k) B(k,j) + bias[j]) for k in this z-slice
//   A_KC: A(i,k) = A[i*lda + k]  else A[k*lda + i];   B_KC: B(k,j) = B[j*ldb + k]  else B[k*ldb + j]
//   NT (forward linear):  A_KC, B_KC      NN (dgrad): A_KC, !B_KC      TN (wgrad): !A_KC, !B_KC
// ------------------------------------------------------------------------------------------
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256)
k_gemm_f32(const float *__restrict__ A, int lda, const float *__restrict__ Bp, int ldb, const float *__restrict__ bias,
           float *__restrict__ C, int ldc, int M, int N, int K, int kchunk, size_t c_zstride, int act) {
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const bool a_vec = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0);
    const bool b_vec = ((ldb & 3) == 0) && ((((uintptr_t)Bp) & 15) == 0);

    for (int k0 = kbeg; k0 < kend; k0 += 16) {
        if (A_KC) {
            const int lrow = tid / 4, lk = (tid % 4) * 4;
            const int m = m0 + lrow;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < M) {
                const float *p = A + (int64_t)m * lda + k0 + lk;
                if (a_vec && k0 + lk + 3 < kend) {
                    const float4 q = *reinterpret_cast<const float4 *>(p);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k0 + lk + i < kend) v[i] = p[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[lk + i][lrow] = v[i];
        } else {
            const int kk = tid / 16, c4 = (tid % 16) * 4;
            const int k = k0 + kk;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (k < kend) {
                const float *p = A + (int64_t)k * lda + m0 + c4;
                if (a_vec && m0 + c4 + 3 < M) {
                    const float4 q = *reinterpret_cast<const float4 *>(p);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (m0 + c4 + i < M) v[i] = p[i];
                }
            }
            *reinterpret_cast<float4 *>(&As[kk][c4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        if (B_KC) {
            const int lrow = tid / 4, lk = (tid % 4) * 4;
            const int n = n0 + lrow;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (n < N) {
                const float *p = Bp + (int64_t)n * ldb + k0 + lk;
                if (b_vec && k0 + lk + 3 < kend) {
                    const float4 q = *reinterpret_cast<const float4 *>(p);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k0 + lk + i < kend) v[i] = p[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[lk + i][lrow] = v[i];
        } else {
            const int kk = tid / 16, c4 = (tid % 16) * 4;
            const int k = k0 + kk;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (k < kend) {
                const float *p = Bp + (int64_t)k * ldb + n0 + c4;
                if (b_vec && n0 + c4 + 3 < N) {
                    const float4 q = *reinterpret_cast<const float4 *>(p);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (n0 + c4 + i < N) v[i] = p[i];
                }
            }
            *reinterpret_cast<float4 *>(&Bs[kk][c4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 w = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *Cz = C + (size_t)blockIdx.z * c_zstride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.0f);
            if (act == 2) v = fmaxf(v, 0.0f);
            Cz[(int64_t)m * ldc + n] = v;
        }
    }
}

// out[e] = sum_z part[z*n + e]
__global__ void k_reduce_z(const float *__restrict__ part, int Z, int64_t n, float *__restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s = 0.f;
    for (int z = 0; z < Z; ++z) s += part[(size_t)z * n + e];
    out[e] = s;
}

// out[c] += sum_m A[m*lda + c]   (out zeroed by the caller)
__global__ void k_colsum(const float *__restrict__ A, int lda, int M, int N, int rows_per_block, float *__restrict__ out) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    for (int m = r0; m < r1; ++m) s += A[(int64_t)m * lda + c];
    atomicAdd(out + c, s);
}

// ------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, biased variance), one warp per row
// ------------------------------------------------------------------------------------------
__global__ void k_ln_fwd(const float *__restrict__ x, const float *__restrict__ g, const float *__restrict__ b, float *__restrict__ y,
                         float *__restrict__ stats /* (M,2) mean, rstd; may be null */, int M, int d, __nv_bfloat16 *__restrict__ yb = nullptr) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *xr = x + (int64_t)row * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mean = warp_sum_f(s) / (float)d;
    float v = 0.f;
    for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; v = fmaf(t, t, v); }
    const float rstd = 1.0f / sqrtf(warp_sum_f(v) / (float)d + 1e-5f);
    for (int c = lane; c < d; c += 32) {
        const float v = (xr[c] - mean) * rstd * g[c] + b[c];
        if (y) y[(int64_t)row * d + c] = v;
        if (yb) yb[(int64_t)row * d + c] = __float2bfloat16_rn(v);
    }
    if (stats && lane == 0) { stats[2 * (int64_t)row] = mean; stats[2 * (int64_t)row + 1] = rstd; }
}

// dx_out[m] = (resid ? resid[m] : 0) + LN'(dy)[m]
__global__ void k_ln_bwd(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ g,
                         const float *resid, float *dx_out, int M, int d) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (row >= M) return;
    const float mean = stats[2 * (int64_t)row], rstd = stats[2 * (int64_t)row + 1];
    const float *xr = x + (int64_t)row * d, *dr = dy + (int64_t)row * d;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float xh = (xr[c] - mean) * rstd, dh = dr[c] * g[c];
        s1 += dh;
        s2 = fmaf(dh, xh, s2);
    }
    s1 = warp_sum_f(s1) / (float)d;
    s2 = warp_sum_f(s2) / (float)d;
    for (int c = lane; c < d; c += 32) {
        const float xh = (xr[c] - mean) * rstd, dh = dr[c] * g[c];
        const float v = rstd * (dh - s1 - xh * s2);
        dx_out[(int64_t)row * d + c] = (resid ? resid[(int64_t)row * d + c] : 0.0f) + v;
    }
}

// One pass: dx_out = resid + LN'(dy) and dgamma / dbeta.  A block takes `rows_per_block` rows, a warp one row at a time; lane l owns
// columns l, l+32, ... (d <= 32 * LNB_MAXC) and keeps its share of the column sums in registers; 8 warps are summed through shared memory.
constexpr int LNB_MAXC = 8;
__global__ void __launch_bounds__(256)
k_ln_bwd_fused(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ g,
               const float *resid, float *dx_out, int M, int d, int rows_per_block, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    __shared__ float red[2][8][32 * LNB_MAXC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float ag[LNB_MAXC], ab[LNB_MAXC], gv[LNB_MAXC];
#pragma unroll
    for (int k = 0; k < LNB_MAXC; ++k) {
        ag[k] = 0.f; ab[k] = 0.f;
        gv[k] = (lane + 32 * k < d) ? g[lane + 32 * k] : 0.f;
    }
    for (int row = r0 + warp; row < r1; row += 8) {
        const float mean = stats[2 * (int64_t)row], rstd = stats[2 * (int64_t)row + 1];
        const float *xr = x + (int64_t)row * d, *dr = dy + (int64_t)row * d;
        float xh[LNB_MAXC], dv[LNB_MAXC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < LNB_MAXC; ++k) {
            const int c = lane + 32 * k;
            xh[k] = c < d ? (xr[c] - mean) * rstd : 0.f;
            dv[k] = c < d ? dr[c] : 0.f;
            const float dh = dv[k] * gv[k];
            s1 += dh;
            s2 = fmaf(dh, xh[k], s2);
            ag[k] = fmaf(dv[k], xh[k], ag[k]);
            ab[k] += dv[k];
        }
        s1 = warp_sum_f(s1) / (float)d;
        s2 = warp_sum_f(s2) / (float)d;
#pragma unroll
        for (int k = 0; k < LNB_MAXC; ++k) {
            const int c = lane + 32 * k;
            if (c < d) {
                const float v = rstd * (dv[k] * gv[k] - s1 - xh[k] * s2);
                dx_out[(int64_t)row * d + c] = (resid ? resid[(int64_t)row * d + c] : 0.0f) + v;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < LNB_MAXC; ++k) {
        red[0][warp][lane + 32 * k] = ag[k];
        red[1][warp][lane + 32 * k] = ab[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += 256) {
        float sg = 0.f, sb = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { sg += red[0][w][c]; sb += red[1][w][c]; }
        atomicAdd(dgamma + c, sg);
        atomicAdd(dbeta + c, sb);
    }
}

// dgamma[c] += sum_m dy*xhat, dbeta[c] += sum_m dy
__global__ void k_ln_param_grad(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ stats, int M, int d,
                                int rows_per_block, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= d) return;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float sg = 0.f, sb = 0.f;
    for (int m = r0; m < r1; ++m) {
        const float v = dy[(int64_t)m * d + c];
        sg = fmaf(v, (x[(int64_t)m * d + c] - stats[2 * (int64_t)m]) * stats[2 * (int64_t)m + 1], sg);
        sb += v;
    }
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
}

// ------------------------------------------------------------------------------------------
// element-wise pieces
// ------------------------------------------------------------------------------------------
// out = dropout(silu(h))
__global__ void k_silu_drop(const float *__restrict__ h, float *__restrict__ out, int64_t n, Drop dr, __nv_bfloat16 *__restrict__ outb = nullptr) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float v = h[e];
        const float r = v * sigmoid_f(v) * drop_scale(dr, (uint64_t)e);
        if (out) out[e] = r;
        if (outb) outb[e] = __float2bfloat16_rn(r);
    }
}
// dh = dact * mask * silu'(h)   (in place on dact)
__global__ void k_silu_drop_bwd(const float *__restrict__ h, float *__restrict__ dact, int64_t n, Drop dr, __nv_bfloat16 *__restrict__ outb = nullptr) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float v = h[e], s = sigmoid_f(v);
        const float r = dact[e] * drop_scale(dr, (uint64_t)e) * (s * (1.0f + v * (1.0f - s)));
        dact[e] = r;
        if (outb) outb[e] = __float2bfloat16_rn(r);
    }
}
// out = x + factor * dropout(y)      (x may be null: out = factor*dropout(y))
__global__ void k_resid_drop(const float *x, const float *__restrict__ y, float *out, float factor, int64_t n, Drop dr,
                             __nv_bfloat16 *__restrict__ outb = nullptr) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float r = (x ? x[e] : 0.0f) + factor * y[e] * drop_scale(dr, (uint64_t)e);
        out[e] = r;
        if (outb) outb[e] = __float2bfloat16_rn(r);
    }
}
// ---- 4-wide variants used by the tensor-core path (cols % 4 == 0, 16-byte aligned rows) -------------------------------------------
// out = (x ? x : 0) + factor * dropout(y) over a (rows, cols) matrix; optional bf16 copy; optional colsum[c] += sum_rows out[., c]
// (the bias gradient of the layer that produced y when this runs in the backward pass).  A block takes `rpb` rows.
__global__ void __launch_bounds__(256)
k_resid_drop_v4(const float *x, const float *__restrict__ y, float *out, float factor, int64_t rows, int cols, Drop dr, __nv_bfloat16 *__restrict__ outb,
                float *__restrict__ colsum, int rpb) {
    extern __shared__ float cs[];
    if (colsum) {
        for (int c = threadIdx.x; c < cols; c += 256) cs[c] = 0.f;
        __syncthreads();
    }
    const int ncg = cols >> 2;
    const int64_t r0 = (int64_t)blockIdx.x * rpb;
    const int nr = (int)min((int64_t)rpb, rows - r0);
    for (int idx = threadIdx.x; idx < nr * ncg; idx += 256) {
        const int r = idx / ncg, c = (idx % ncg) << 2;
        const int64_t e = (r0 + r) * cols + c;
        const float4 yv = *reinterpret_cast<const float4 *>(y + e);
        const float4 m = drop_scale4(dr, (uint64_t)e);
        float4 o = make_float4(factor * yv.x * m.x, factor * yv.y * m.y, factor * yv.z * m.z, factor * yv.w * m.w);
        if (x) {
            const float4 xv = *reinterpret_cast<const float4 *>(x + e);
            o.x += xv.x; o.y += xv.y; o.z += xv.z; o.w += xv.w;
        }
        *reinterpret_cast<float4 *>(out + e) = o;
        if (outb) store_bf16x4(outb + e, o);
        if (colsum) {
            atomicAdd(&cs[c], o.x); atomicAdd(&cs[c + 1], o.y); atomicAdd(&cs[c + 2], o.z); atomicAdd(&cs[c + 3], o.w);
        }
    }
    if (colsum) {
        __syncthreads();
        for (int c = threadIdx.x; c < cols; c += 256) atomicAdd(colsum + c, cs[c]);
    }
}
// outb = bf16(dropout(silu(h)))
__global__ void k_silu_drop_v4(const float *__restrict__ h, int64_t n, Drop dr, __nv_bfloat16 *__restrict__ outb) {
    for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < n; e += (int64_t)gridDim.x * blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4 *>(h + e);
        const float4 m = drop_scale4(dr, (uint64_t)e);
        store_bf16x4(outb + e, make_float4(v.x * sigmoid_f(v.x) * m.x, v.y * sigmoid_f(v.y) * m.y, v.z * sigmoid_f(v.z) * m.z, v.w * sigmoid_f(v.w) * m.w));
    }
}
// dh = dact * mask * silu'(h) (in place, + bf16 copy) and colsum[c] += sum_rows dh[., c]
__global__ void __launch_bounds__(256)
k_silu_drop_bwd_v4(const float *__restrict__ h, float *__restrict__ dact, int64_t rows, int cols, Drop dr, __nv_bfloat16 *__restrict__ outb,
                   float *__restrict__ colsum, int rpb) {
    extern __shared__ float cs[];
    for (int c = threadIdx.x; c < cols; c += 256) cs[c] = 0.f;
    __syncthreads();
    const int ncg = cols >> 2;
    const int64_t r0 = (int64_t)blockIdx.x * rpb;
    const int nr = (int)min((int64_t)rpb, rows - r0);
    for (int idx = threadIdx.x; idx < nr * ncg; idx += 256) {
        const int r = idx / ncg, c = (idx % ncg) << 2;
        const int64_t e = (r0 + r) * cols + c;
        const float4 hv = *reinterpret_cast<const float4 *>(h + e), dv = *reinterpret_cast<const float4 *>(dact + e);
        const float4 m = drop_scale4(dr, (uint64_t)e);
        const float hs[4] = {hv.x, hv.y, hv.z, hv.w}, ds[4] = {dv.x * m.x, dv.y * m.y, dv.z * m.z, dv.w * m.w};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float sg = sigmoid_f(hs[i]);
            o[i] = ds[i] * (sg * (1.0f + hs[i] * (1.0f - sg)));
            atomicAdd(&cs[c + i], o[i]);
        }
        const float4 ov = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4 *>(dact + e) = ov;
        if (outb) store_bf16x4(outb + e, ov);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += 256) atomicAdd(colsum + c, cs[c]);
}

// GLU over the feature axis: u (M, 2d) -> g (M, d) = u[:, :d] * sigmoid(u[:, d:])   (convolution.py:139)
__global__ void k_glu_fwd(const float *__restrict__ u, float *__restrict__ g, int64_t M, int d) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = e / d;
        const int c = (int)(e % d);
        g[e] = u[m * 2 * d + c] * sigmoid_f(u[m * 2 * d + d + c]);
    }
}
__global__ void k_glu_bwd(const float *__restrict__ u, const float *__restrict__ dg, float *__restrict__ du, int64_t M, int d,
                          __nv_bfloat16 *__restrict__ dub = nullptr) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = e / d;
        const int c = (int)(e % d);
        const float a = u[m * 2 * d + c], s = sigmoid_f(u[m * 2 * d + d + c]), v = dg[e];
        const float da = v * s, db = v * a * s * (1.0f - s);
        du[m * 2 * d + c] = da;
        du[m * 2 * d + d + c] = db;
        if (dub) {
            dub[m * 2 * d + c] = __float2bfloat16_rn(da);
            dub[m * 2 * d + d + c] = __float2bfloat16_rn(db);
        }
    }
}

// ------------------------------------------------------------------------------------------
// depthwise Conv1d (kernel K, pad (K-1)/2, no bias, zero padding at the line edges): x, y (B,T,d); w (d,K)
// flip: correlate with the reversed taps (= gradient with respect to the input)
// ------------------------------------------------------------------------------------------
constexpr int DWT = 32;
__global__ void __launch_bounds__(256)
k_dwconv1d(const float *__restrict__ x, const float *__restrict__ w, float *__restrict__ y, int T, int d, int K, int flip) {
    extern __shared__ float sm[];
    float *tile = sm;                               // [DWT + K - 1][d]
    float *wt = sm + (size_t)(DWT + K - 1) * d;     // [K][d]
    const int b = blockIdx.y, t0 = blockIdx.x * DWT, pad = (K - 1) / 2, rows = DWT + K - 1;
    for (int e = threadIdx.x; e < rows * d; e += blockDim.x) {
        const int r = e / d, c = e % d, t = t0 - pad + r;
        tile[e] = (t >= 0 && t < T) ? x[((int64_t)b * T + t) * d + c] : 0.0f;
    }
    for (int e = threadIdx.x; e < K * d; e += blockDim.x) {
        const int k = e / d, c = e % d;
        wt[e] = w[c * K + (flip ? K - 1 - k : k)];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < DWT * d; e += blockDim.x) {
        const int tt = e / d, c = e % d, t = t0 + tt;
        if (t >= T) break;
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) acc = fmaf(wt[k * d + c], tile[(tt + k) * d + c], acc);
        y[((int64_t)b * T + t) * d + c] = acc;
    }
}
// K = 31: tiles of 64 frames; a thread owns a channel and 8 consecutive frames and slides a register window over its 38 input rows
// (one shared-memory read per input row and per tap instead of two per multiply-add)
__global__ void __launch_bounds__(256)
k_dwconv1d_k31(const float *__restrict__ x, const float *__restrict__ w, float *__restrict__ y, int T, int d, int flip) {
    constexpr int K = 31, TT = 64, OUT = 8;
    extern __shared__ float sm[];
    float *tile = sm;                          // [TT + 30][d]
    float *wt = sm + (size_t)(TT + K - 1) * d; // [31][d]
    const int b = blockIdx.y, t0 = blockIdx.x * TT, rows = TT + K - 1;
    for (int e = threadIdx.x; e < rows * d; e += 256) {
        const int r = e / d, c = e % d, t = t0 - 15 + r;
        tile[e] = (t >= 0 && t < T) ? x[((int64_t)b * T + t) * d + c] : 0.0f;
    }
    for (int e = threadIdx.x; e < K * d; e += 256) {
        const int k = e / d, c = e % d;
        wt[e] = w[c * K + (flip ? K - 1 - k : k)];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < (TT / OUT) * d; e += 256) {
        const int g = e / d, c = e % d;                    // group of 8 frames, channel (consecutive threads: consecutive channels)
        float acc[OUT], wr[K];
#pragma unroll
        for (int o = 0; o < OUT; ++o) acc[o] = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) wr[k] = wt[k * d + c];       // the channel's taps in registers
        const float *in = tile + (size_t)(g * OUT) * d + c;
#pragma unroll
        for (int r = 0; r < OUT + K - 1; ++r) {            // input row r of the window contributes to outputs o = r - k
            const float v = in[(size_t)r * d];
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                const int k = r - o;
                if (k >= 0 && k < K) acc[o] = fmaf(wr[k], v, acc[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
            const int t = t0 + g * OUT + o;
            if (t < T) y[((int64_t)b * T + t) * d + c] = acc[o];
        }
    }
}
// Weight gradient for K = 31: a thread owns a channel and 8 consecutive frames, reads its 8 output gradients and 38 input rows once
// and keeps the 31 tap sums in registers (46 shared-memory reads per 248 multiply-adds instead of 2 per 1); the 8 frame groups of a
// channel meet in a shared-memory accumulator, one global atomic per (channel, tap) and CTA.
__global__ void __launch_bounds__(256)
k_dwconv1d_wgrad_k31(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ dw, int T, int d) {
    constexpr int K = 31, TT = 64, OUT = 8;
    extern __shared__ float sm[];
    float *tile = sm;                                // [TT + 30][d]  x
    float *dt = tile + (size_t)(TT + K - 1) * d;     // [TT][d]       dy
    float *accs = dt + (size_t)TT * d;               // [K][d]
    const int b = blockIdx.y, t0 = blockIdx.x * TT, rows = TT + K - 1;
    for (int e = threadIdx.x; e < rows * d; e += 256) {
        const int r = e / d, c = e % d, t = t0 - 15 + r;
        tile[e] = (t >= 0 && t < T) ? x[((int64_t)b * T + t) * d + c] : 0.0f;
    }
    for (int e = threadIdx.x; e < TT * d; e += 256) {
        const int tt = e / d, c = e % d, t = t0 + tt;
        dt[e] = (t < T) ? dy[((int64_t)b * T + t) * d + c] : 0.0f;
    }
    for (int e = threadIdx.x; e < K * d; e += 256) accs[e] = 0.0f;
    __syncthreads();
    for (int e = threadIdx.x; e < (TT / OUT) * d; e += 256) {
        const int g = e / d, c = e % d;
        float acc[K], dv[OUT];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) dv[o] = dt[(size_t)(g * OUT + o) * d + c];
        const float *in = tile + (size_t)(g * OUT) * d + c;
#pragma unroll
        for (int r = 0; r < OUT + K - 1; ++r) {
            const float v = in[(size_t)r * d];
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                const int k = r - o;
                if (k >= 0 && k < K) acc[k] = fmaf(dv[o], v, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) atomicAdd(&accs[k * d + c], acc[k]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < K * d; e += 256) {
        const int k = e / d, c = e % d;
        atomicAdd(dw + c * K + k, accs[e]);
    }
}
// dw[c,k] += sum_{b,t} dy[b,t,c] * x[b,t+k-pad,c]      (tiles of DWT frames; the caller picks DWT so that the tile fits shared memory)
__global__ void __launch_bounds__(256)
k_dwconv1d_wgrad(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ dw, int T, int d, int K, int DWT) {
    extern __shared__ float sm[];
    float *tile = sm;                               // [DWT + K - 1][d]  x
    float *dt = sm + (size_t)(DWT + K - 1) * d;     // [DWT][d]          dy
    const int b = blockIdx.y, t0 = blockIdx.x * DWT, pad = (K - 1) / 2, rows = DWT + K - 1;
    for (int e = threadIdx.x; e < rows * d; e += blockDim.x) {
        const int r = e / d, c = e % d, t = t0 - pad + r;
        tile[e] = (t >= 0 && t < T) ? x[((int64_t)b * T + t) * d + c] : 0.0f;
    }
    for (int e = threadIdx.x; e < DWT * d; e += blockDim.x) {
        const int tt = e / d, c = e % d, t = t0 + tt;
        dt[e] = (t < T) ? dy[((int64_t)b * T + t) * d + c] : 0.0f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < K * d; e += blockDim.x) {
        const int k = e / d, c = e % d;
        float acc = 0.0f;
#pragma unroll 8
        for (int tt = 0; tt < DWT; ++tt) acc = fmaf(dt[tt * d + c], tile[(tt + k) * d + c], acc);
        atomicAdd(dw + c * K + k, acc);
    }
}

// ------------------------------------------------------------------------------------------
// BatchNorm1d in training mode over the M = B*T rows of c (M, d)   (convolution.py:141; eps 1e-5, momentum 0.1)
// ------------------------------------------------------------------------------------------
__global__ void k_bn_sums(const float *__restrict__ c, int M, int d, int rows_per_block, double *__restrict__ sums /* [2][d] */) {
    const int col = blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= d) return;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
    double s = 0.0, q = 0.0;
    for (int m = r0; m < r1; ++m) {
        const double v = c[(int64_t)m * d + col];
        s += v;
        q += v * v;
    }
    atomicAdd(sums + col, s);
    atomicAdd(sums + d + col, q);
}
__global__ void k_bn_finalize(const double *__restrict__ sums, int M, int d, float momentum, float *__restrict__ stat /* [2][d] mean, rstd */,
                              float *running_mean, float *running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    const double mean = sums[c] / M;
    double var = sums[d + c] / M - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[c] = (float)mean;
    stat[d + c] = (float)(1.0 / sqrt(var + 1e-5));
    if (running_mean) {
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}
// z = silu(gamma * (c - mean) * rstd + beta)
__global__ void k_bn_silu_fwd(const float *__restrict__ c, const float *__restrict__ stat, const float *__restrict__ gamma,
                              const float *__restrict__ beta, float *__restrict__ z, int64_t M, int d, __nv_bfloat16 *__restrict__ zb = nullptr) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % d);
        const float y = gamma[col] * (c[e] - stat[col]) * stat[d + col] + beta[col];
        const float r = y * sigmoid_f(y);
        if (z) z[e] = r;
        if (zb) zb[e] = __float2bfloat16_rn(r);
    }
}
// dy = dz * silu'(y) written over dz; sums[0][c] += dy, sums[1][c] += dy * chat
__global__ void k_bn_bwd_reduce(const float *__restrict__ c, const float *__restrict__ stat, const float *__restrict__ gamma,
                                const float *__restrict__ beta, float *__restrict__ dz, int M, int d, int rows_per_block, double *__restrict__ sums) {
    const int col = blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= d) return;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
    const float mean = stat[col], rstd = stat[d + col], ga = gamma[col], be = beta[col];
    double s = 0.0, q = 0.0;
    for (int m = r0; m < r1; ++m) {
        const int64_t e = (int64_t)m * d + col;
        const float ch = (c[e] - mean) * rstd, y = ga * ch + be, sg = sigmoid_f(y);
        const float dy = dz[e] * (sg * (1.0f + y * (1.0f - sg)));
        dz[e] = dy;
        s += dy;
        q += (double)dy * ch;
    }
    atomicAdd(sums + col, s);
    atomicAdd(sums + d + col, q);
}
// dc = gamma * rstd * (dy - sum(dy)/M - chat * sum(dy*chat)/M)    (dy in, dc out, in place); block 0 also writes dgamma, dbeta
__global__ void k_bn_bwd_apply(const float *__restrict__ c, const float *__restrict__ stat, const float *__restrict__ gamma,
                               const double *__restrict__ sums, float *__restrict__ dy, int64_t M, int d, float *dgamma, float *dbeta) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % d);
        const float rstd = stat[d + col], ch = (c[e] - stat[col]) * rstd;
        const float s = (float)(sums[col] / (double)M), q = (float)(sums[d + col] / (double)M);
        dy[e] = gamma[col] * rstd * (dy[e] - s - ch * q);
    }
    if (blockIdx.x == 0)
        for (int col = threadIdx.x; col < d; col += blockDim.x) {
            dbeta[col] = (float)sums[col];
            dgamma[col] = (float)sums[d + col];
        }
}

// ------------------------------------------------------------------------------------------
// relative-position attention, reference formulation (attention.py:82-101), probabilities kept for the backward pass.
//   qkv (M, 3d): q | k | v, head-major features; pos (2T-1, d); P (B, nh, T, T)
// thread layout of the tile kernels: 256 threads = 32 rows x 8 column groups
// ------------------------------------------------------------------------------------------
constexpr int TQ = 32;    // rows per CTA
constexpr int TK = 64;    // columns per step

// S[b,h,i,j] = ((q_i+u).k_j + (q_i+v).pos[T-1-i+j]) / sqrt(dh)
template <int DH>
__global__ void __launch_bounds__(256)
k_att_scores(const float *__restrict__ qkv, const float *__restrict__ pos, const float *__restrict__ ub, const float *__restrict__ vb,
             float *__restrict__ S, int T, int d, int dh, float inv_sqrt) {
    extern __shared__ float sm[];
    constexpr int LD = DH + 1;
    float *ks = sm;                         // [TK][LD]
    float *ps = ks + TK * LD;               // [TQ+TK-1][LD]
    float *st = ps + (TQ + TK - 1) * LD;    // [TQ][TK+1]
    const int tid = threadIdx.x, ti = tid / 8, tj = tid % 8;
    const int i0 = blockIdx.x * TQ, hd = blockIdx.y, b = blockIdx.z, i = i0 + ti;
    const int ld = 3 * d;
    const float *base = qkv + (int64_t)b * T * ld;
    float qu[DH], qv[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) {
        const float q = (c < dh && i < T) ? base[(int64_t)i * ld + hd * dh + c] : 0.0f;
        qu[c] = (c < dh) ? q + ub[hd * dh + c] : 0.0f;
        qv[c] = (c < dh) ? q + vb[hd * dh + c] : 0.0f;
    }
    float *Sb = S + ((int64_t)(b * gridDim.y + hd) * T) * T;
    for (int j0 = 0; j0 < T; j0 += TK) {
        __syncthreads();
        for (int e = tid; e < TK * DH; e += 256) {
            const int r = e / DH, c = e % DH, j = j0 + r;
            ks[r * LD + c] = (j < T && c < dh) ? base[(int64_t)j * ld + d + hd * dh + c] : 0.0f;
        }
        const int pbase = T - 1 - (i0 + TQ - 1) + j0;
        for (int e = tid; e < (TQ + TK - 1) * DH; e += 256) {
            const int r = e / DH, c = e % DH, pr = pbase + r;
            ps[r * LD + c] = (pr >= 0 && pr <= 2 * T - 2 && c < dh) ? pos[(int64_t)pr * d + hd * dh + c] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int n = 0; n < TK / 8; ++n) {
            const int jl = tj + 8 * n;
            const float *kr = ks + jl * LD, *pr = ps + (TQ - 1 - ti + jl) * LD;
            float c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int c = 0; c < DH; ++c) {
                c1 = fmaf(qu[c], kr[c], c1);
                c2 = fmaf(qv[c], pr[c], c2);
            }
            st[ti * (TK + 1) + jl] = (c1 + c2) * inv_sqrt;
        }
        __syncthreads();
        for (int e = tid; e < TQ * TK; e += 256) {
            const int r = e / TK, c = e % TK;
            if (i0 + r < T && j0 + c < T) Sb[(int64_t)(i0 + r) * T + j0 + c] = st[r * (TK + 1) + c];
        }
    }
}

// in-place softmax over rows of length T, one warp per row
__global__ void k_softmax_rows(float *__restrict__ S, int64_t rows, int T) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float *p = S + row * T;
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, p[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int j = lane; j < T; j += 32) {
        const float e = expf(p[j] - mx);
        p[j] = e;
        s += e;
    }
    s = warp_sum_f(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < T; j += 32) p[j] *= inv;
}

// ctx[b,i,h*dh+c] = sum_j dropout(P)[b,h,i,j] * v[b,j,h,c]
template <int DH>
__global__ void __launch_bounds__(256)
k_att_pv(const float *__restrict__ qkv, const float *__restrict__ P, float *__restrict__ ctx, int T, int d, int dh, Drop dr) {
    extern __shared__ float sm[];
    constexpr int LD = DH + 1;
    float *vs = sm;                    // [TK][LD]
    float *pt = vs + TK * LD;          // [TQ][TK+1]
    const int tid = threadIdx.x, ti = tid / 8, tj = tid % 8;
    const int i0 = blockIdx.x * TQ, hd = blockIdx.y, b = blockIdx.z, i = i0 + ti;
    const int ld = 3 * d;
    const float *base = qkv + (int64_t)b * T * ld;
    const int64_t prow0 = ((int64_t)(b * gridDim.y + hd) * T);
    float o[(DH + 7) / 8];
#pragma unroll
    for (int k = 0; k < (DH + 7) / 8; ++k) o[k] = 0.0f;
    for (int j0 = 0; j0 < T; j0 += TK) {
        __syncthreads();
        for (int e = tid; e < TK * DH; e += 256) {
            const int r = e / DH, c = e % DH, j = j0 + r;
            vs[r * LD + c] = (j < T && c < dh) ? base[(int64_t)j * ld + 2 * d + hd * dh + c] : 0.0f;
        }
        for (int e = tid; e < TQ * TK; e += 256) {
            const int r = e / TK, c = e % TK;
            float v = 0.0f;
            if (i0 + r < T && j0 + c < T) {
                const int64_t idx = (prow0 + i0 + r) * T + j0 + c;
                v = P[idx] * drop_scale(dr, (uint64_t)idx);
            }
            pt[r * (TK + 1) + c] = v;
        }
        __syncthreads();
        for (int jl = 0; jl < TK; ++jl) {
            const float p = pt[ti * (TK + 1) + jl];
            const float *vr = vs + jl * LD;
#pragma unroll
            for (int k = 0; k < (DH + 7) / 8; ++k) {
                const int c = tj + 8 * k;
                if (c < DH) o[k] = fmaf(p, vr[c], o[k]);
            }
        }
    }
    if (i < T) {
#pragma unroll
        for (int k = 0; k < (DH + 7) / 8; ++k) {
            const int c = tj + 8 * k;
            if (c < dh) ctx[((int64_t)b * T + i) * d + hd * dh + c] = o[k];
        }
    }
}

// Row pass of the backward: dS = P * (mask * (dctx.v) - delta) / sqrt(dh) with delta_i = dctx_i . ctx_i, written to dS;
// dq_i = sum_j dS[i,j] (k_j + pos[T-1-i+j]); du += sum_i sum_j dS k_j; dvb += sum_i sum_j dS pos
template <int DH>
__global__ void __launch_bounds__(256)
k_att_bwd_rows(const float *__restrict__ qkv, const float *__restrict__ pos, const float *__restrict__ P, const float *__restrict__ ctx,
               const float *__restrict__ dctx, float *__restrict__ dS, float *__restrict__ dqkv, float *__restrict__ du, float *__restrict__ dvb,
               int T, int d, int dh, float inv_sqrt, Drop dr) {
    extern __shared__ float sm[];
    constexpr int LD = DH + 1;
    float *ks = sm;                         // [TK][LD]
    float *vs = ks + TK * LD;               // [TK][LD]
    float *ps = vs + TK * LD;               // [TQ+TK-1][LD]
    float *pt = ps + (TQ + TK - 1) * LD;    // [TQ][TK+1]
    const int tid = threadIdx.x, ti = tid / 8, tj = tid % 8;
    const int i0 = blockIdx.x * TQ, hd = blockIdx.y, b = blockIdx.z, i = i0 + ti;
    const int ld = 3 * d;
    const float *base = qkv + (int64_t)b * T * ld;
    const int64_t prow0 = ((int64_t)(b * gridDim.y + hd) * T);
    float dc[DH], accA[DH], accB[DH];
    float delta = 0.f;
#pragma unroll
    for (int c = 0; c < DH; ++c) {
        const bool ok = (c < dh && i < T);
        dc[c] = ok ? dctx[((int64_t)b * T + i) * d + hd * dh + c] : 0.0f;
        delta = fmaf(dc[c], ok ? ctx[((int64_t)b * T + i) * d + hd * dh + c] : 0.0f, delta);
        accA[c] = 0.f;
        accB[c] = 0.f;
    }
    for (int j0 = 0; j0 < T; j0 += TK) {
        __syncthreads();
        for (int e = tid; e < TK * DH; e += 256) {
            const int r = e / DH, c = e % DH, j = j0 + r;
            const bool ok = (j < T && c < dh);
            ks[r * LD + c] = ok ? base[(int64_t)j * ld + d + hd * dh + c] : 0.0f;
            vs[r * LD + c] = ok ? base[(int64_t)j * ld + 2 * d + hd * dh + c] : 0.0f;
        }
        const int pbase = T - 1 - (i0 + TQ - 1) + j0;
        for (int e = tid; e < (TQ + TK - 1) * DH; e += 256) {
            const int r = e / DH, c = e % DH, pr = pbase + r;
            ps[r * LD + c] = (pr >= 0 && pr <= 2 * T - 2 && c < dh) ? pos[(int64_t)pr * d + hd * dh + c] : 0.0f;
        }
        for (int e = tid; e < TQ * TK; e += 256) {
            const int r = e / TK, c = e % TK;
            pt[r * (TK + 1) + c] = (i0 + r < T && j0 + c < T) ? P[(prow0 + i0 + r) * T + j0 + c] : 0.0f;
        }
        __syncthreads();
#pragma unroll 1
        for (int n = 0; n < TK / 8; ++n) {
            const int jl = tj + 8 * n;
            const float *vr = vs + jl * LD;
            float dp = 0.f;
#pragma unroll
            for (int c = 0; c < DH; ++c) dp = fmaf(dc[c], vr[c], dp);
            const float p = pt[ti * (TK + 1) + jl];
            float ds = 0.0f;
            if (i < T && j0 + jl < T) {
                const int64_t idx = (prow0 + i) * T + j0 + jl;
                ds = p * (dp * drop_scale(dr, (uint64_t)idx) - delta) * inv_sqrt;
            }
            pt[ti * (TK + 1) + jl] = ds;
            const float *kr = ks + jl * LD, *pr = ps + (TQ - 1 - ti + jl) * LD;
#pragma unroll
            for (int c = 0; c < DH; ++c) {
                accA[c] = fmaf(ds, kr[c], accA[c]);
                accB[c] = fmaf(ds, pr[c], accB[c]);
            }
        }
        __syncthreads();
        for (int e = tid; e < TQ * TK; e += 256) {
            const int r = e / TK, c = e % TK;
            if (i0 + r < T && j0 + c < T) dS[(prow0 + i0 + r) * T + j0 + c] = pt[r * (TK + 1) + c];
        }
    }
    // reduce over the 8 column groups of a row (lanes tj = 0..7 are consecutive lanes)
#pragma unroll
    for (int c = 0; c < DH; ++c) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            accA[c] += __shfl_xor_sync(0xffffffffu, accA[c], o);
            accB[c] += __shfl_xor_sync(0xffffffffu, accB[c], o);
        }
    }
    __syncthreads();
    float *ra = sm, *rb = sm + TQ * LD;      // [TQ][LD] each (aliases the tiles)
    if (tj == 0) {
#pragma unroll
        for (int c = 0; c < DH; ++c) {
            ra[ti * LD + c] = accA[c];
            rb[ti * LD + c] = accB[c];
            if (i < T && c < dh) dqkv[((int64_t)b * T + i) * ld + hd * dh + c] = accA[c] + accB[c];
        }
    }
    __syncthreads();
    if (tid < dh) {
        float sa = 0.f, sb = 0.f;
        for (int r = 0; r < TQ; ++r) {
            sa += ra[r * LD + tid];
            sb += rb[r * LD + tid];
        }
        atomicAdd(du + hd * dh + tid, sa);
        atomicAdd(dvb + hd * dh + tid, sb);
    }
}

// Column pass: dv_j = sum_i dropout(P)[i,j] dctx_i ; dk_j = sum_i dS[i,j] (q_i + u).  256 threads = 32 columns j x 8 row groups
template <int DH>
__global__ void __launch_bounds__(256)
k_att_bwd_cols(const float *__restrict__ qkv, const float *__restrict__ ub, const float *__restrict__ P, const float *__restrict__ dS,
               const float *__restrict__ dctx, float *__restrict__ dqkv, int T, int d, int dh, Drop dr) {
    extern __shared__ float sm[];
    constexpr int LD = DH + 1;
    float *qs = sm;                 // [TK][LD]  q + u
    float *ds_ = qs + TK * LD;      // [TK][LD]  dctx
    const int tid = threadIdx.x, tjj = tid % 32, tii = tid / 32;
    const int j0 = blockIdx.x * 32, hd = blockIdx.y, b = blockIdx.z, j = j0 + tjj;
    const int ld = 3 * d;
    const float *base = qkv + (int64_t)b * T * ld;
    const int64_t prow0 = ((int64_t)(b * gridDim.y + hd) * T);
    float accV[DH], accK[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) { accV[c] = 0.f; accK[c] = 0.f; }
    for (int i0 = 0; i0 < T; i0 += TK) {
        __syncthreads();
        for (int e = tid; e < TK * DH; e += 256) {
            const int r = e / DH, c = e % DH, i = i0 + r;
            const bool ok = (i < T && c < dh);
            qs[r * LD + c] = ok ? base[(int64_t)i * ld + hd * dh + c] + ub[hd * dh + c] : 0.0f;
            ds_[r * LD + c] = ok ? dctx[((int64_t)b * T + i) * d + hd * dh + c] : 0.0f;
        }
        __syncthreads();
#pragma unroll 1
        for (int n = 0; n < TK / 8; ++n) {
            const int il = tii + 8 * n, i = i0 + il;
            float p = 0.f, g = 0.f;
            if (i < T && j < T) {
                const int64_t idx = (prow0 + i) * T + j;
                p = P[idx] * drop_scale(dr, (uint64_t)idx);
                g = dS[idx];
            }
            const float *dr_ = ds_ + il * LD, *qr = qs + il * LD;
#pragma unroll
            for (int c = 0; c < DH; ++c) {
                accV[c] = fmaf(p, dr_[c], accV[c]);
                accK[c] = fmaf(g, qr[c], accK[c]);
            }
        }
    }
    // reduce over the 8 row groups through shared memory: red [8][32][LD]
    __syncthreads();
    float *red = sm;
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int c = 0; c < DH; ++c) red[(tii * 32 + tjj) * LD + c] = pass ? accK[c] : accV[c];
        __syncthreads();
        for (int e = tid; e < 32 * DH; e += 256) {
            const int r = e / DH, c = e % DH;
            if (j0 + r < T && c < dh) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += red[(w * 32 + r) * LD + c];
                dqkv[((int64_t)b * T + j0 + r) * ld + (pass ? d : 2 * d) + hd * dh + c] = s;
            }
        }
        __syncthreads();
    }
}

// dpos[r, h*dh+c] += sum_{b,i} dS[b,h,i, r-(T-1)+i] * (q[b,i,h,c] + v_bias[h,c])   lanes <-> 32 consecutive r
template <int DH>
__global__ void __launch_bounds__(256)
k_att_bwd_pos(const float *__restrict__ qkv, const float *__restrict__ vb, const float *__restrict__ dS, float *__restrict__ dpos,
              int B, int T, int d, int dh, int nh, int b_per_block) {
    extern __shared__ float sm[];
    constexpr int LD = (DH + 3) / 4 * 4;
    float *qt = sm;                       // [64][LD]  q + v_bias of 64 (b,i) rows
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int r0 = blockIdx.x * 32, hd = blockIdx.y, r = r0 + lane;
    const int b0 = blockIdx.z * b_per_block, b1 = min(B, b0 + b_per_block);
    const int ld = 3 * d;
    const int ilo = max(0, T - 1 - (r0 + 31)), ihi = min(T - 1, 2 * T - 2 - r0);      // rows with at least one valid column
    float acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
    if (ihi >= ilo) {
        const int ni = ihi - ilo + 1;
        const int total = (b1 - b0) * ni;
        for (int base = 0; base < total; base += 64) {
            __syncthreads();
            for (int e = tid; e < 64 * DH; e += 256) {
                const int rr = e / DH, c = e % DH, t = base + rr;
                float v = 0.f;
                if (t < total && c < dh) {
                    const int b = b0 + t / ni, i = ilo + t % ni;
                    v = qkv[((int64_t)b * T + i) * ld + hd * dh + c] + vb[hd * dh + c];
                }
                qt[rr * LD + c] = v;
            }
            __syncthreads();
#pragma unroll 1
            for (int n = 0; n < 8; ++n) {
                const int rr = w * 8 + n, t = base + rr;
                if (t >= total) break;
                const int b = b0 + t / ni, i = ilo + t % ni;
                const int j = r - (T - 1) + i;
                float g = 0.f;
                if (r <= 2 * T - 2 && j >= 0 && j < T) g = dS[(((int64_t)(b * nh + hd) * T) + i) * T + j];
                const float *qr = qt + rr * LD;
#pragma unroll
                for (int c = 0; c < DH; ++c) acc[c] = fmaf(g, qr[c], acc[c]);
            }
        }
    }
    __syncthreads();
    float *red = sm;                      // [8][32][DH+1]
    constexpr int LR = DH + 1;
#pragma unroll
    for (int c = 0; c < DH; ++c) red[(w * 32 + lane) * LR + c] = acc[c];
    __syncthreads();
    for (int e = tid; e < 32 * DH; e += 256) {
        const int rr = e / DH, c = e % DH;
        if (r0 + rr <= 2 * T - 2 && c < dh) {
            float s = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) s += red[(ww * 32 + rr) * LR + c];
            atomicAdd(dpos + (int64_t)(r0 + rr) * d + hd * dh + c, s);
        }
    }
}

// ------------------------------------------------------------------------------------------
// pieces of the tensor-core (bf16 operand) path
// ------------------------------------------------------------------------------------------
// dst (rows, ldd) bf16 <- src (rows, lds) fp32, columns [0, cols)
__global__ void k_cast_bf16(const float *__restrict__ src, int64_t lds, __nv_bfloat16 *__restrict__ dst, int64_t ldd, int64_t rows, int cols) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * cols; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols;
        const int c = (int)(e % cols);
        dst[r * ldd + c] = __float2bfloat16_rn(src[r * lds + c]);
    }
}
// head-pitched copy: dst (rows, nh*dhp) bf16 <- src (rows, lds) fp32 columns [col0 + h*dh + c] (+ bias[h*dh + c])
__global__ void k_head_pitch(const float *__restrict__ src, int64_t lds, int col0, const float *__restrict__ bias, __nv_bfloat16 *__restrict__ dst,
                             int64_t rows, int nh, int dh, int dhp) {
    const int d = nh * dh;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / d;
        const int c = (int)(e % d), hd = c / dh, cc = c % dh;
        dst[r * nh * dhp + hd * dhp + cc] = __float2bfloat16_rn(src[r * lds + col0 + c] + (bias ? bias[c] : 0.0f));
    }
}
// qkv (M, 3d) fp32 -> q+u | q+v | k | v, each head-pitched (M, nh*dhp) bf16
__global__ void k_att_prep(const float *__restrict__ qkv, const float *__restrict__ ub, const float *__restrict__ vb, __nv_bfloat16 *__restrict__ qu,
                           __nv_bfloat16 *__restrict__ qv, __nv_bfloat16 *__restrict__ kb, __nv_bfloat16 *__restrict__ vbuf, int64_t M, int nh, int dh,
                           int dhp) {
    const int d = nh * dh;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / d;
        const int c = (int)(e % d), hd = c / dh, cc = c % dh;
        const int64_t o = r * nh * dhp + hd * dhp + cc;
        const float q = qkv[r * 3 * d + c];
        qu[o] = __float2bfloat16_rn(q + ub[c]);
        qv[o] = __float2bfloat16_rn(q + vb[c]);
        kb[o] = __float2bfloat16_rn(qkv[r * 3 * d + d + c]);
        vbuf[o] = __float2bfloat16_rn(qkv[r * 3 * d + 2 * d + c]);
    }
}
// dctx (M, d) fp32 -> head-pitched bf16, and delta[(b,i), h] = sum_c dctx * ctx
__global__ void k_att_bwd_prep(const float *__restrict__ dctx, const float *__restrict__ ctx, __nv_bfloat16 *__restrict__ dctxb, float *__restrict__ delta,
                               int64_t M, int nh, int dh, int dhp) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * nh; e += (int64_t)gridDim.x * blockDim.x) {
        const float *a = dctx + e * dh, *b = ctx + e * dh;
        __nv_bfloat16 *o = dctxb + (e / nh) * nh * dhp + (e % nh) * dhp;
        float s = 0.f;
        for (int c = 0; c < dh; ++c) {
            s = fmaf(a[c], b[c], s);
            o[c] = __float2bfloat16_rn(a[c]);
        }
        delta[e] = s;
    }
}
// P = softmax((Sc + shift(PS)) / sqrt(dh)) per row; one warp per row (b, h, i).  Sc (rows, Tp) fp32, PS (rows, Rp) fp32: the shifted
// positional scores of row i are the CONTIGUOUS segment PS[i, T-1-i .. 2T-2-i] (attention.py:105-113).  Writes P and dropout(P) as bf16.
__global__ void k_att_softmax_tc(const float *__restrict__ Sc, const float *__restrict__ PS, __nv_bfloat16 *__restrict__ P, __nv_bfloat16 *__restrict__ Pd,
                                 int64_t rows, int T, int Tp, int Rp, float inv_sqrt, Drop dr) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int i = (int)(row % T);
    const float *sc = Sc + row * Tp, *ps = PS + row * Rp + (T - 1 - i);
    if (T <= 512) {   // the row lives in registers, 8 consecutive columns per lane and step (Tp is a multiple of 8): 16-byte loads of
                      // the content scores, 16-byte stores of the bf16 rows, one Philox call per 4 probabilities
        float v[16];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int j = 8 * lane + 256 * k;
            float sv[8];
            if (j < Tp) {
                const float4 s0 = *reinterpret_cast<const float4 *>(sc + j), s1 = *reinterpret_cast<const float4 *>(sc + j + 4);
                sv[0] = s0.x; sv[1] = s0.y; sv[2] = s0.z; sv[3] = s0.w; sv[4] = s1.x; sv[5] = s1.y; sv[6] = s1.z; sv[7] = s1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[8 * k + i] = (j + i < T) ? (sv[i] + ps[j + i]) * inv_sqrt : -INFINITY;
                mx = fmaxf(mx, v[8 * k + i]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            v[k] = (v[k] > -INFINITY) ? expf(v[k] - mx) : 0.f;
            s += v[k];
        }
        s = warp_sum_f(s);
        const float inv = 1.0f / s;
        const bool quad_rng = (T & 3) == 0;              // element index row*T + j is then a multiple of 4 with j
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int j = 8 * lane + 256 * k;
            if (j >= Tp) break;
            float pv[8], pd[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = v[8 * k + i] * inv;
            uint4 w;
            {
                const __nv_bfloat162 a = __floats2bfloat162_rn(pv[0], pv[1]), b = __floats2bfloat162_rn(pv[2], pv[3]),
                                     c = __floats2bfloat162_rn(pv[4], pv[5]), e = __floats2bfloat162_rn(pv[6], pv[7]);
                w = make_uint4(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b), *reinterpret_cast<const uint32_t *>(&c),
                               *reinterpret_cast<const uint32_t *>(&e));
            }
            *reinterpret_cast<uint4 *>(P + row * Tp + j) = w;
            if (Pd != P) {
                if (quad_rng) {
                    const float4 m0 = drop_scale4(dr, (uint64_t)(row * T + j)), m1 = drop_scale4(dr, (uint64_t)(row * T + j + 4));
                    pd[0] = pv[0] * m0.x; pd[1] = pv[1] * m0.y; pd[2] = pv[2] * m0.z; pd[3] = pv[3] * m0.w;
                    pd[4] = pv[4] * m1.x; pd[5] = pv[5] * m1.y; pd[6] = pv[6] * m1.z; pd[7] = pv[7] * m1.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) pd[i] = (j + i < T) ? pv[i] * drop_scale(dr, (uint64_t)(row * T + j + i)) : 0.f;
                }
                const __nv_bfloat162 a = __floats2bfloat162_rn(pd[0], pd[1]), b = __floats2bfloat162_rn(pd[2], pd[3]),
                                     c = __floats2bfloat162_rn(pd[4], pd[5]), e = __floats2bfloat162_rn(pd[6], pd[7]);
                *reinterpret_cast<uint4 *>(Pd + row * Tp + j) = make_uint4(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b),
                                                                            *reinterpret_cast<const uint32_t *>(&c), *reinterpret_cast<const uint32_t *>(&e));
            }
        }
        return;
    }
    if (T <= 512) {         // the row lives in registers: one read of the scores
        float v[16];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int j = lane + 32 * k;
            v[k] = j < T ? (sc[j] + ps[j]) * inv_sqrt : -INFINITY;
            mx = fmaxf(mx, v[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            v[k] = (lane + 32 * k < T) ? expf(v[k] - mx) : 0.f;
            s += v[k];
        }
        s = warp_sum_f(s);
        const float inv = 1.0f / s;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int j = lane + 32 * k;
            if (j >= Tp) break;
            const float pv = v[k] * inv;
            P[row * Tp + j] = __float2bfloat16_rn(pv);
            if (Pd != P) Pd[row * Tp + j] = __float2bfloat16_rn(j < T ? pv * drop_scale(dr, (uint64_t)(row * T + j)) : 0.f);
        }
        return;
    }
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, (sc[j] + ps[j]) * inv_sqrt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int j = lane; j < T; j += 32) s += expf((sc[j] + ps[j]) * inv_sqrt - mx);
    s = warp_sum_f(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < Tp; j += 32) {
        float pv = 0.f;
        if (j < T) pv = expf((sc[j] + ps[j]) * inv_sqrt - mx) * inv;
        P[row * Tp + j] = __float2bfloat16_rn(pv);
        if (Pd != P) Pd[row * Tp + j] = __float2bfloat16_rn(pv * drop_scale(dr, (uint64_t)(row * T + j)));
    }
}
// delta[(b,i), h] = sum_c dctx[b,i,h,c] * ctx[b,i,h,c]
__global__ void k_att_delta(const float *__restrict__ dctx, const float *__restrict__ ctx, float *__restrict__ delta, int64_t M, int nh, int dh) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * nh; e += (int64_t)gridDim.x * blockDim.x) {
        const float *a = dctx + e * dh, *b = ctx + e * dh;         // (M, nh*dh) row-major: element e covers dh consecutive columns
        float s = 0.f;
        for (int c = 0; c < dh; ++c) s = fmaf(a[c], b[c], s);
        delta[e] = s;
    }
}
// dS = P * (mask * dPd - delta) / sqrt(dh) -> dSb (rows, Tp) bf16 and its un-shifted copy dPSb (rows, Rp) bf16 (zeros outside the segment)
__global__ void k_att_ds_tc(const __nv_bfloat16 *__restrict__ P, const float *__restrict__ dPd, const float *__restrict__ delta,
                            __nv_bfloat16 *__restrict__ dSb, __nv_bfloat16 *__restrict__ dPSb, int64_t rows, int T, int Tp, int Rp, int nh,
                            float inv_sqrt, Drop dr) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int i = (int)(row % T);
    const int64_t bh = row / T;
    const int hd = (int)(bh % nh);
    const int64_t b = bh / nh;
    const float dl = delta[(b * T + i) * nh + hd];
    const int seg0 = T - 1 - i;
    __nv_bfloat16 *ps = dPSb + row * Rp;
    for (int r = lane; r < Rp; r += 32)
        if (r < seg0 || r >= seg0 + T) ps[r] = __float2bfloat16_rn(0.0f);
    // pairs of columns per lane (Tp is even): 4-byte loads / stores of the bf16 rows, 8-byte loads of the fp32 row
    for (int j = 2 * lane; j < Tp; j += 64) {
        const __nv_bfloat162 p2 = *reinterpret_cast<const __nv_bfloat162 *>(P + row * Tp + j);
        const float2 g2 = *reinterpret_cast<const float2 *>(dPd + row * Tp + j);
        float d0 = 0.f, d1 = 0.f;
        if (j < T) d0 = __low2float(p2) * (g2.x * drop_scale(dr, (uint64_t)(row * T + j)) - dl) * inv_sqrt;
        if (j + 1 < T) d1 = __high2float(p2) * (g2.y * drop_scale(dr, (uint64_t)(row * T + j + 1)) - dl) * inv_sqrt;
        const __nv_bfloat162 v2 = __floats2bfloat162_rn(d0, d1);
        *reinterpret_cast<__nv_bfloat162 *>(dSb + row * Tp + j) = v2;
        if (j < T) ps[seg0 + j] = __low2bfloat16(v2);
        if (j + 1 < T) ps[seg0 + j + 1] = __high2bfloat16(v2);
    }
}
// The same for T % 4 == 0 with every global access at least 8 bytes wide: 4 columns per lane and step (one Philox call), the un-shifted
// row assembled in shared memory (zeros, then the segment) and written out as 16-byte chunks.  Dynamic shared memory: 8 warps x Rp bf16.
__global__ void __launch_bounds__(256)
k_att_ds_tc_v4(const __nv_bfloat16 *__restrict__ P, const float *__restrict__ dPd, const float *__restrict__ delta, __nv_bfloat16 *__restrict__ dSb,
               __nv_bfloat16 *__restrict__ dPSb, int64_t rows, int T, int Tp, int Rp, int nh, float inv_sqrt, Drop dr) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __nv_bfloat16 *srow = reinterpret_cast<__nv_bfloat16 *>(smraw) + (size_t)warp * Rp;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const int i = (int)(row % T);
    const int64_t bh = row / T;
    const float dl = delta[((bh / nh) * T + i) * nh + (int)(bh % nh)];
    const int seg0 = T - 1 - i;
    for (int c = lane; c < Rp / 8; c += 32) *reinterpret_cast<uint4 *>(srow + 8 * c) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    for (int j = 4 * lane; j < Tp; j += 128) {
        float4 ds = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < T) {                                     // T % 4 == 0: the four columns are valid together
            const uint2 pw = *reinterpret_cast<const uint2 *>(P + row * Tp + j);
            const __nv_bfloat162 p01 = *reinterpret_cast<const __nv_bfloat162 *>(&pw.x), p23 = *reinterpret_cast<const __nv_bfloat162 *>(&pw.y);
            const float4 g = *reinterpret_cast<const float4 *>(dPd + row * Tp + j);
            const float4 mk = drop_scale4(dr, (uint64_t)(row * T + j));
            ds.x = __low2float(p01) * (g.x * mk.x - dl) * inv_sqrt;
            ds.y = __high2float(p01) * (g.y * mk.y - dl) * inv_sqrt;
            ds.z = __low2float(p23) * (g.z * mk.z - dl) * inv_sqrt;
            ds.w = __high2float(p23) * (g.w * mk.w - dl) * inv_sqrt;
        }
        const __nv_bfloat162 a = __floats2bfloat162_rn(ds.x, ds.y), b = __floats2bfloat162_rn(ds.z, ds.w);
        *reinterpret_cast<uint2 *>(dSb + row * Tp + j) = make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b));
        if (j < T) {
            srow[seg0 + j] = __low2bfloat16(a); srow[seg0 + j + 1] = __high2bfloat16(a);
            srow[seg0 + j + 2] = __low2bfloat16(b); srow[seg0 + j + 3] = __high2bfloat16(b);
        }
    }
    __syncwarp();
    for (int c = lane; c < Rp / 8; c += 32) *reinterpret_cast<uint4 *>(dPSb + row * Rp + 8 * c) = *reinterpret_cast<const uint4 *>(srow + 8 * c);
}
// dst[r, c] = a[r, c] + b[r, c] for c < cols (dst row pitch ldd)
__global__ void k_add2(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ dst, int64_t ldd, int64_t rows, int cols) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * cols; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols;
        const int c = (int)(e % cols);
        dst[r * ldd + c] = a[e] + b[e];
    }
}

__global__ void k_set_seed(uint64_t *dst, uint64_t v) { *dst = v; }


// ------------------------------------------------------------------------------------------
// 16-byte-access variants of the bandwidth-bound kernels (feature counts that are multiples of 4, which the default spec and
// every test geometry are); the scalar kernels above stay as the general path.  Measured on the dS kernel: 2.2 -> 4.7 TB/s.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// LayerNorm forward, one warp per row, d <= 256: lane owns the float4 chunks lane and lane + 32
__global__ void k_ln_fwd_v4(const float *__restrict__ x, const float *__restrict__ g, const float *__restrict__ b, float *__restrict__ y,
                            float *__restrict__ stats, int M, int d, __nv_bfloat16 *__restrict__ yb) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (row >= M) return;
    const int nq = d >> 2;
    const float *xr = x + (int64_t)row * d;
    float4 v[2];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int q = lane + 32 * k;
        v[k] = q < nq ? ld4(xr + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    const float mean = warp_sum_f(s) / (float)d;
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < nq) {
            const float a = v[k].x - mean, bb = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
            var += a * a + bb * bb + c * c + e * e;
        }
    const float rstd = 1.0f / sqrtf(warp_sum_f(var) / (float)d + 1e-5f);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int q = lane + 32 * k;
        if (q < nq) {
            const float4 gg = ld4(g + 4 * q), bv = ld4(b + 4 * q);
            const float4 o = make_float4((v[k].x - mean) * rstd * gg.x + bv.x, (v[k].y - mean) * rstd * gg.y + bv.y,
                                         (v[k].z - mean) * rstd * gg.z + bv.z, (v[k].w - mean) * rstd * gg.w + bv.w);
            if (y) st4(y + (int64_t)row * d + 4 * q, o);
            if (yb) store_bf16x4(yb + (int64_t)row * d + 4 * q, o);
        }
    }
    if (stats && lane == 0) { stats[2 * (int64_t)row] = mean; stats[2 * (int64_t)row + 1] = rstd; }
}

// fused LayerNorm backward (dx, dgamma, dbeta), d <= 256 and d % 4 == 0
__global__ void __launch_bounds__(256)
k_ln_bwd_fused_v4(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ g,
                  const float *resid, float *dx_out, int M, int d, int rows_per_block, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    __shared__ float red[2][8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = d >> 2;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float4 ag[2], ab[2], gv[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        ab[k] = ag[k];
        gv[k] = (lane + 32 * k < nq) ? ld4(g + 4 * (lane + 32 * k)) : ag[k];
    }
    for (int row = r0 + warp; row < r1; row += 8) {
        const float mean = stats[2 * (int64_t)row], rstd = stats[2 * (int64_t)row + 1];
        const float *xr = x + (int64_t)row * d, *dr = dy + (int64_t)row * d;
        float4 xh[2], dv[2];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int q = lane + 32 * k;
            if (q < nq) {
                const float4 xv = ld4(xr + 4 * q);
                dv[k] = ld4(dr + 4 * q);
                xh[k] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
            } else {
                dv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                xh[k] = dv[k];
            }
            const float4 dh = make_float4(dv[k].x * gv[k].x, dv[k].y * gv[k].y, dv[k].z * gv[k].z, dv[k].w * gv[k].w);
            s1 += dh.x + dh.y + dh.z + dh.w;
            s2 += dh.x * xh[k].x + dh.y * xh[k].y + dh.z * xh[k].z + dh.w * xh[k].w;
            ag[k].x = fmaf(dv[k].x, xh[k].x, ag[k].x); ag[k].y = fmaf(dv[k].y, xh[k].y, ag[k].y);
            ag[k].z = fmaf(dv[k].z, xh[k].z, ag[k].z); ag[k].w = fmaf(dv[k].w, xh[k].w, ag[k].w);
            ab[k].x += dv[k].x; ab[k].y += dv[k].y; ab[k].z += dv[k].z; ab[k].w += dv[k].w;
        }
        s1 = warp_sum_f(s1) / (float)d;
        s2 = warp_sum_f(s2) / (float)d;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int q = lane + 32 * k;
            if (q < nq) {
                float4 o = make_float4(rstd * (dv[k].x * gv[k].x - s1 - xh[k].x * s2), rstd * (dv[k].y * gv[k].y - s1 - xh[k].y * s2),
                                       rstd * (dv[k].z * gv[k].z - s1 - xh[k].z * s2), rstd * (dv[k].w * gv[k].w - s1 - xh[k].w * s2));
                if (resid) {
                    const float4 r = ld4(resid + (int64_t)row * d + 4 * q);
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                st4(dx_out + (int64_t)row * d + 4 * q, o);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int q = lane + 32 * k;
        if (q < 64) {
            st4(&red[0][warp][4 * q], ag[k]);
            st4(&red[1][warp][4 * q], ab[k]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += 256) {
        float sg = 0.f, sb = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { sg += red[0][w][c]; sb += red[1][w][c]; }
        atomicAdd(dgamma + c, sg);
        atomicAdd(dbeta + c, sb);
    }
}

// dst (rows, ldd) bf16 <- src (rows, lds) fp32, columns [0, cols); cols, lds, ldd multiples of 4
__global__ void k_cast_bf16_v4(const float *__restrict__ src, int64_t lds, __nv_bfloat16 *__restrict__ dst, int64_t ldd, int64_t rows, int cols) {
    const int nq = cols >> 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * nq; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / nq;
        const int c = (int)(e % nq) << 2;
        store_bf16x4(dst + r * ldd + c, ld4(src + r * lds + c));
    }
}
// qkv (M, 3d) fp32 -> q+u | q+v | k | v head-pitched bf16; dh % 4 == 0, dhp % 4 == 0
__global__ void k_att_prep_v4(const float *__restrict__ qkv, const float *__restrict__ ub, const float *__restrict__ vb, __nv_bfloat16 *__restrict__ qu,
                              __nv_bfloat16 *__restrict__ qv, __nv_bfloat16 *__restrict__ kb, __nv_bfloat16 *__restrict__ vbuf, int64_t M, int nh, int dh,
                              int dhp) {
    const int d = nh * dh, nq = d >> 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * nq; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / nq;
        const int c = (int)(e % nq) << 2, hd = c / dh, cc = c % dh;
        const int64_t o = r * nh * dhp + hd * dhp + cc;
        const float4 q = ld4(qkv + r * 3 * d + c), u = ld4(ub + c), v = ld4(vb + c);
        store_bf16x4(qu + o, make_float4(q.x + u.x, q.y + u.y, q.z + u.z, q.w + u.w));
        store_bf16x4(qv + o, make_float4(q.x + v.x, q.y + v.y, q.z + v.z, q.w + v.w));
        store_bf16x4(kb + o, ld4(qkv + r * 3 * d + d + c));
        store_bf16x4(vbuf + o, ld4(qkv + r * 3 * d + 2 * d + c));
    }
}
// dctx (M, d) fp32 -> head-pitched bf16 and delta[(b,i), h] = sum_c dctx * ctx; one thread per (row, head, 4 columns), quads of ... 
// kept simple: one thread per (row, head), 16-byte loads, 8-byte stores
__global__ void k_att_bwd_prep_v4(const float *__restrict__ dctx, const float *__restrict__ ctx, __nv_bfloat16 *__restrict__ dctxb, float *__restrict__ delta,
                                  int64_t M, int nh, int dh, int dhp) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * nh; e += (int64_t)gridDim.x * blockDim.x) {
        const float *a = dctx + e * dh, *b = ctx + e * dh;
        __nv_bfloat16 *o = dctxb + (e / nh) * nh * dhp + (e % nh) * dhp;
        float s = 0.f;
        for (int c = 0; c < dh; c += 4) {
            const float4 av = ld4(a + c), bv = ld4(b + c);
            s += av.x * bv.x + av.y * bv.y + av.z * bv.z + av.w * bv.w;
            store_bf16x4(o + c, av);
        }
        delta[e] = s;
    }
}
// GLU forward / backward over (M, 2d) -> (M, d), d % 4 == 0
__global__ void k_glu_fwd_v4(const float *__restrict__ u, float *__restrict__ g, int64_t M, int d) {
    const int nq = d >> 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * nq; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = e / nq;
        const int c = (int)(e % nq) << 2;
        const float4 a = ld4(u + m * 2 * d + c), b = ld4(u + m * 2 * d + d + c);
        st4(g + m * d + c, make_float4(a.x * sigmoid_f(b.x), a.y * sigmoid_f(b.y), a.z * sigmoid_f(b.z), a.w * sigmoid_f(b.w)));
    }
}
__global__ void k_glu_bwd_v4(const float *__restrict__ u, const float *__restrict__ dg, float *__restrict__ du, int64_t M, int d,
                             __nv_bfloat16 *__restrict__ dub) {
    const int nq = d >> 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * nq; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = e / nq;
        const int c = (int)(e % nq) << 2;
        const float4 a = ld4(u + m * 2 * d + c), b = ld4(u + m * 2 * d + d + c), v = ld4(dg + m * d + c);
        const float4 sg = make_float4(sigmoid_f(b.x), sigmoid_f(b.y), sigmoid_f(b.z), sigmoid_f(b.w));
        const float4 da = make_float4(v.x * sg.x, v.y * sg.y, v.z * sg.z, v.w * sg.w);
        const float4 db = make_float4(v.x * a.x * sg.x * (1.f - sg.x), v.y * a.y * sg.y * (1.f - sg.y), v.z * a.z * sg.z * (1.f - sg.z),
                                      v.w * a.w * sg.w * (1.f - sg.w));
        st4(du + m * 2 * d + c, da);
        st4(du + m * 2 * d + d + c, db);
        if (dub) {
            store_bf16x4(dub + m * 2 * d + c, da);
            store_bf16x4(dub + m * 2 * d + d + c, db);
        }
    }
}
// z = silu(BatchNorm(c)) -> bf16 only (the tensor-core path's operand), d % 4 == 0
__global__ void k_bn_silu_fwd_v4(const float *__restrict__ c, const float *__restrict__ stat, const float *__restrict__ gamma,
                                 const float *__restrict__ beta, int64_t M, int d, __nv_bfloat16 *__restrict__ zb) {
    const int nq = d >> 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * nq; e += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(e % nq) << 2;
        const float4 v = ld4(c + 4 * e), mu = ld4(stat + col), rs = ld4(stat + d + col), ga = ld4(gamma + col), be = ld4(beta + col);
        const float y0 = ga.x * (v.x - mu.x) * rs.x + be.x, y1 = ga.y * (v.y - mu.y) * rs.y + be.y, y2 = ga.z * (v.z - mu.z) * rs.z + be.z,
                    y3 = ga.w * (v.w - mu.w) * rs.w + be.w;
        store_bf16x4(zb + 4 * e, make_float4(y0 * sigmoid_f(y0), y1 * sigmoid_f(y1), y2 * sigmoid_f(y2), y3 * sigmoid_f(y3)));
    }
}

// relative positional table rows k = 0..2T-2 <-> offset T-1-k (embedding.py:42-56)
__global__ void k_rel_pos_table(float *__restrict__ pe, int T, int d) {
    const int R = 2 * T - 1;
    const float neg = -(logf(10000.0f) / (float)d);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < R * d; e += gridDim.x * blockDim.x) {
        const int k = e / d, c = e % d, m = c & ~1;
        const float arg = (float)(T - 1 - k) * expf((float)m * neg);
        pe[e] = (c & 1) ? cosf(arg) : sinf(arg);
    }
}

// ------------------------------------------------------------------------------------------
// convolutional front-end (convolution.py:190-237), channel-last activations: a1 (B,T2,F2,cc), a2 (B,T,F,cc)
// ------------------------------------------------------------------------------------------
__global__ void k_fe_conv1(const float *__restrict__ img, int64_t sb, int64_t sw, int64_t sh, const float *__restrict__ w1,
                           const float *__restrict__ b1, float *__restrict__ a1, int B, int W, int H, int T2, int F2, int cc) {
    const int64_t total = (int64_t)B * T2 * F2 * cc;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cc);
        int64_t r = idx / cc;
        const int f2 = (int)(r % F2);
        r /= F2;
        const int t2 = (int)(r % T2), b = (int)(r / T2);
        float acc = b1[c];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int w = 2 * t2 - 1 + kh;
            if (w < 0 || w >= W) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int y = 2 * f2 - 1 + kw;
                if (y < 0 || y >= H) continue;
                acc = fmaf(w1[c * 9 + kh * 3 + kw], img[b * sb + w * sw + y * sh], acc);
            }
        }
        a1[idx] = fmaxf(acc, 0.0f);
    }
}
__global__ void k_fe_dw(const float *__restrict__ a1, const float *__restrict__ wdw, const float *__restrict__ bdw, float *__restrict__ a2,
                        int B, int T2, int F2, int T, int F, int cc) {
    const int64_t total = (int64_t)B * T * F * cc;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cc);
        int64_t r = idx / cc;
        const int f = (int)(r % F);
        r /= F;
        const int t = (int)(r % T), b = (int)(r / T);
        float acc = bdw[c];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int t2 = 2 * t - 1 + kh;
            if (t2 < 0 || t2 >= T2) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int f2 = 2 * f - 1 + kw;
                if (f2 < 0 || f2 >= F2) continue;
                acc = fmaf(wdw[c * 9 + kh * 3 + kw], a1[(((int64_t)b * T2 + t2) * F2 + f2) * cc + c], acc);
            }
        }
        a2[idx] = acc;
    }
}
// in place: v = (act > 0) ? g : 0     (ReLU backward)
__global__ void k_relu_bwd(const float *__restrict__ act, float *__restrict__ g, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        if (!(act[e] > 0.0f)) g[e] = 0.0f;
}
// dwdw[c,kh,kw] += sum da2[b,t,f,c] * a1[b,2t-1+kh,2f-1+kw,c]; dbdw[c] += sum da2.   blockDim (256) % cc == 0
__global__ void __launch_bounds__(256)
k_fe_dw_wgrad(const float *__restrict__ a1, const float *__restrict__ da2, float *__restrict__ dw, float *__restrict__ db,
              int B, int T2, int F2, int T, int F, int cc) {
    __shared__ float red[256][11];
    const int64_t total = (int64_t)B * T * F * cc;
    float acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.f;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        const int c = (int)(idx % cc);
        int64_t r = idx / cc;
        const int f = (int)(r % F);
        r /= F;
        const int t = (int)(r % T), b = (int)(r / T);
        const float g = da2[idx];
        acc[9] += g;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int t2 = 2 * t - 1 + kh;
            if (t2 < 0 || t2 >= T2) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int f2 = 2 * f - 1 + kw;
                if (f2 < 0 || f2 >= F2) continue;
                acc[kh * 3 + kw] = fmaf(g, a1[(((int64_t)b * T2 + t2) * F2 + f2) * cc + c], acc[kh * 3 + kw]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) red[threadIdx.x][k] = acc[k];
    __syncthreads();
    for (int e = threadIdx.x; e < cc * 10; e += 256) {
        const int c = e / 10, k = e % 10;
        float s = 0.f;
        for (int t = c; t < 256; t += cc) s += red[t][k];
        if (k < 9) atomicAdd(dw + c * 9 + k, s);
        else atomicAdd(db + c, s);
    }
}
// da1 = relu'(a1) * conv_transpose(da2, wdw), written over a1
__global__ void k_fe_dw_dgrad_relu(float *__restrict__ a1, const float *__restrict__ da2, const float *__restrict__ wdw,
                                   int B, int T2, int F2, int T, int F, int cc) {
    const int64_t total = (int64_t)B * T2 * F2 * cc;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cc);
        int64_t r = idx / cc;
        const int f2 = (int)(r % F2);
        r /= F2;
        const int t2 = (int)(r % T2), b = (int)(r / T2);
        float acc = 0.f;
        if (a1[idx] > 0.0f) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int tn = t2 + 1 - kh;
                if (tn < 0 || (tn & 1)) continue;
                const int t = tn >> 1;
                if (t >= T) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int fn = f2 + 1 - kw;
                    if (fn < 0 || (fn & 1)) continue;
                    const int f = fn >> 1;
                    if (f >= F) continue;
                    acc = fmaf(wdw[c * 9 + kh * 3 + kw], da2[(((int64_t)b * T + t) * F + f) * cc + c], acc);
                }
            }
        }
        a1[idx] = acc;
    }
}
// dW1[c,kh,kw] += sum da1[b,t2,f2,c] * img[b,2t2-1+kh,2f2-1+kw]; db1[c] += sum da1
__global__ void __launch_bounds__(256)
k_fe_conv1_wgrad(const float *__restrict__ img, int64_t sb, int64_t sw, int64_t sh, const float *__restrict__ da1, float *__restrict__ dw,
                 float *__restrict__ db, int B, int W, int H, int T2, int F2, int cc) {
    __shared__ float red[256][11];
    const int64_t total = (int64_t)B * T2 * F2 * cc;
    float acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.f;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        int64_t r = idx / cc;
        const int f2 = (int)(r % F2);
        r /= F2;
        const int t2 = (int)(r % T2), b = (int)(r / T2);
        const float g = da1[idx];
        acc[9] += g;
        if (g == 0.0f) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int w = 2 * t2 - 1 + kh;
            if (w < 0 || w >= W) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int y = 2 * f2 - 1 + kw;
                if (y < 0 || y >= H) continue;
                acc[kh * 3 + kw] = fmaf(g, img[b * sb + w * sw + y * sh], acc[kh * 3 + kw]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) red[threadIdx.x][k] = acc[k];
    __syncthreads();
    for (int e = threadIdx.x; e < cc * 10; e += 256) {
        const int c = e / 10, k = e % 10;
        float s = 0.f;
        for (int t = c; t < 256; t += cc) s += red[t][k];
        if (k < 9) atomicAdd(dw + c * 9 + k, s);
        else atomicAdd(db + c, s);
    }
}
// out.0.weight (d, cc*F) column c*F+f  <->  (d, F*cc) column f*cc+c ; to_raw: dst is the raw layout
__global__ void k_permute_wout(const float *__restrict__ src, float *__restrict__ dst, int d, int cc, int F, int to_raw) {
    const int n = d * cc * F;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int row = e / (cc * F), col = e % (cc * F);          // col in the channel-last layout: f*cc + c
        const int f = col / cc, c = col % cc;
        const int raw = row * cc * F + c * F + f;
        if (to_raw) dst[raw] = src[e];
        else dst[e] = src[raw];
    }
}

// ------------------------------------------------------------------------------------------
// AdamW, the arithmetic of torch.optim.AdamW's single-tensor path (model.py:272-274)
// ------------------------------------------------------------------------------------------
__global__ void k_adamw(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int64_t n,
                        float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float gr = g[e];
        float pv = p[e] * (1.0f - lr * wd);
        const float mv = m[e] + (gr - m[e]) * (1.0f - beta1);              // lerp, as torch does
        const float vv = beta2 * v[e] + (1.0f - beta2) * gr * gr;
        m[e] = mv;
        v[e] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pv -= (lr / bc1) * (mv / denom);
        p[e] = pv;
    }
}

}  // namespace train
}  // namespace cocr
