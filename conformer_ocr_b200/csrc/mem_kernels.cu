// Memory-bound kernels of the bf16 path, written for HBM/L2 efficiency: 16-byte vector accesses, shared-memory
// staging with halos, register sliding windows.
//   * k_glu_dw_fast   : GLU -> depthwise conv (K taps) -> folded BatchNorm -> SiLU      (convolution.py:139-142)
//   * k_frontend_fused: Conv2d(1->32,3x3,s2)+ReLU -> depthwise Conv2d(3x3,s2) -> 1x1 conv+ReLU -> (t, f*32+c) bf16
//                       features for the 768->144 projection                              (convolution.py:190-237)
#include <string.h>

#include "common.h"

namespace cocr {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ------------------------------------------------------------------------------------------------------------
// conv module middle.  u (B,T,2d) bf16 [value | gate]  ->  g (B,T,d) bf16.   One CTA = DW_TT frames of one line,
// thread = (channel pair, segment of DW_SEG consecutive frames); the K-tap window slides through registers.
// ------------------------------------------------------------------------------------------------------------
constexpr int DWF_TT = 64;
constexpr int DWF_SEG = 8;

template <int K>
__global__ void __launch_bounds__(576)
k_glu_dw_fast(const bf16 *__restrict__ u, const float *__restrict__ wdw, const float *__restrict__ scale,
              const float *__restrict__ shift, bf16 *__restrict__ g, int T, int d) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const int P = d >> 1;                                   // channel pairs
    constexpr int ROWS = DWF_TT + K - 1;
    uint32_t *xs = reinterpret_cast<uint32_t *>(sm_raw);    // [ROWS][P] bf16x2 (GLU output)
    float2 *ws = reinterpret_cast<float2 *>(xs + ROWS * P); // [K][P]
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * DWF_TT;
    constexpr int pad = (K - 1) / 2;
    const int nthr = blockDim.x;

    // phase 1: GLU of rows [t0-pad, t0+TT+pad) into shared memory (zeros outside the line: conv zero padding)
    const int c8n = d >> 3;                                 // 16-byte chunks per half row
    for (int e = threadIdx.x; e < ROWS * c8n; e += nthr) {
        const int row = e / c8n, c8 = e - row * c8n;
        const int t = t0 - pad + row;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && t < T) {
            const bf16 *src = u + ((int64_t)b * T + t) * 2 * d + c8 * 8;
            const uint4 va = *reinterpret_cast<const uint4 *>(src);
            const uint4 vg = *reinterpret_cast<const uint4 *>(src + d);
            const uint32_t a[4] = {va.x, va.y, va.z, va.w}, gt[4] = {vg.x, vg.y, vg.z, vg.w};
            uint32_t r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 fa = bf2_to_f2(a[i]), fg = bf2_to_f2(gt[i]);
                r[i] = f2_to_bf2(fa.x * sigmoid_fast(fg.x), fa.y * sigmoid_fast(fg.y));
            }
            o = make_uint4(r[0], r[1], r[2], r[3]);
        }
        *reinterpret_cast<uint4 *>(xs + row * P + c8 * 4) = o;
    }
    for (int e = threadIdx.x; e < K * P; e += nthr) {
        const int k = e / P, pr = e - k * P;
        ws[e] = make_float2(wdw[(2 * pr) * K + k], wdw[(2 * pr + 1) * K + k]);
    }
    __syncthreads();

    // phase 2
    const int pr = threadIdx.x % P, seg = threadIdx.x / P;
    if (seg * DWF_SEG >= DWF_TT) return;
    float2 x[DWF_SEG + K - 1];
#pragma unroll
    for (int i = 0; i < DWF_SEG + K - 1; ++i) x[i] = bf2_to_f2(xs[(seg * DWF_SEG + i) * P + pr]);
    float2 acc[DWF_SEG];
#pragma unroll
    for (int i = 0; i < DWF_SEG; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float2 w = ws[k * P + pr];
#pragma unroll
        for (int i = 0; i < DWF_SEG; ++i) {
            acc[i].x = fmaf(w.x, x[i + k].x, acc[i].x);
            acc[i].y = fmaf(w.y, x[i + k].y, acc[i].y);
        }
    }
    const float2 sc = make_float2(scale[2 * pr], scale[2 * pr + 1]);
    const float2 sh = make_float2(shift[2 * pr], shift[2 * pr + 1]);
#pragma unroll
    for (int i = 0; i < DWF_SEG; ++i) {
        const int t = t0 + seg * DWF_SEG + i;
        if (t < T) {
            const float y0 = fmaf(acc[i].x, sc.x, sh.x), y1 = fmaf(acc[i].y, sc.y, sh.y);
            *reinterpret_cast<uint32_t *>(g + ((int64_t)b * T + t) * d + 2 * pr) = f2_to_bf2(y0 * sigmoid_fast(y0), y1 * sigmoid_fast(y1));
        }
    }
}

// returns COCR_ERR_UNSUPPORTED (without touching h->err) when the geometry is not covered; the caller then uses
// the generic kernel of fp32_path.cu
int launch_glu_dw_fast(cocr_handle *h, const bf16 *u, const LayerF32 &lw, bf16 *g, int B, int T, cudaStream_t st) {
    const Dims &dm = h->dm;
    const int P = dm.d / 2;
    const int nthr = P * (DWF_TT / DWF_SEG);
    if (dm.K != 31 || (dm.d & 7) != 0 || nthr > 576) return COCR_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(DWF_TT + 30) * P * 4 + (size_t)31 * P * 8;
    static PerDevice<bool> attr;
    if (!attr.get()) {
        COCR_CUDA(h, cudaFuncSetAttribute(k_glu_dw_fast<31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.get() = true;
    }
    dim3 grid((T + DWF_TT - 1) / DWF_TT, B);
    ProfScope ps(h, "glu_dwconv_bn_silu_bf16", st, 2.0 * 31 * B * (double)T * dm.d, 6.0 * B * T * dm.d);
    k_glu_dw_fast<31><<<grid, nthr, smem, st>>>(u, lw.w_dw, lw.bn_scale, lw.bn_shift, g, T, dm.d);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------------------------
// conv module middle when the GLU already ran in the pw1 GEMM epilogue: x (B,T,d) bf16 -> depthwise conv (K taps,
// zero padding at the line ends) -> folded BatchNorm -> SiLU -> g (B,T,d) bf16.  One CTA = DWF_TT frames of one line;
// the input rows (with halo) are staged in shared memory with 16-byte copies, thread = (channel pair, DWF_SEG frames),
// the K-tap window slides through registers and both channels of a pair share one packed FFMA2 per tap.
// BatchNorm's scale is folded into the taps when they are staged.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 dw_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
        "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float dw_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Persistent: each CTA stages the (pre-scaled, pre-transposed) taps once and walks over (line, time tile) items; the rows
// of the next item stream into the other shared-memory buffer with cp.async (zero fill outside the line) while the
// current item is computed.
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 16 : 0;                            // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int K>
__global__ void __launch_bounds__(576)
k_dw_fast(const bf16 *__restrict__ x, const float2 *__restrict__ taps /* [K][P] w*scale/2 */, const float2 *__restrict__ hshift /* [P] */,
          bf16 *__restrict__ g, int T, int d, int n_tt, int n_items) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const int P = d >> 1;                                   // channel pairs
    constexpr int ROWS = DWF_TT + K - 1;
    constexpr int pad = (K - 1) / 2;
    float2 *ws = reinterpret_cast<float2 *>(sm_raw);        // [K][P]
    uint32_t *xs0 = reinterpret_cast<uint32_t *>(ws + K * P);   // 2 x [ROWS][P] bf16x2
    const int nthr = blockDim.x;
    const int c8n = d >> 3;                                 // 16-byte chunks per row
    // thread -> (first row, 16-byte chunk) of the staging copy, fixed for the whole kernel when the block size is a multiple
    // of the chunks per row (576 = 32 x 18 for d = 144): no divisions inside the copy loop
    const int st_row0 = threadIdx.x / c8n, st_c8 = threadIdx.x - st_row0 * c8n;
    const int st_step = nthr / c8n;                        // rows advanced per pass (nthr % c8n == 0, checked by the launcher)
    auto stage = [&](int item, int buf) {
        const int b = item / n_tt, t0 = (item - b * n_tt) * DWF_TT;
        uint32_t *xs = xs0 + buf * ROWS * P;
        const bf16 *src = x + (int64_t)b * T * d + st_c8 * 8;
        for (int row = st_row0; row < ROWS; row += st_step) {
            const int t = t0 - pad + row;
            const bool ok = t >= 0 && t < T;
            cp_async16(xs + row * P + st_c8 * 4, src + (int64_t)(ok ? t : 0) * d, ok);
        }
        cp_async_commit();
    };
    int item = blockIdx.x;
    if (item < n_items) stage(item, 0);
    for (int e = threadIdx.x; e < K * P; e += nthr) ws[e] = taps[e];
    const int pr = threadIdx.x % P, seg = threadIdx.x / P;
    const float2 hs = hshift[pr];
    int buf = 0;
    for (; item < n_items; item += gridDim.x, buf ^= 1) {
        cp_async_wait<0>();
        __syncthreads();                                    // rows of `item` visible; everyone is done with the other buffer
        if (item + (int)gridDim.x < n_items) stage(item + gridDim.x, buf ^ 1);
        if (seg * DWF_SEG < DWF_TT) {
            const int b = item / n_tt, t0 = (item - b * n_tt) * DWF_TT;
            const uint32_t *xs = xs0 + buf * ROWS * P;
            float2 xw[DWF_SEG + K - 1];
#pragma unroll
            for (int i = 0; i < DWF_SEG + K - 1; ++i) xw[i] = bf2_to_f2(xs[(seg * DWF_SEG + i) * P + pr]);
            float2 acc[DWF_SEG];
#pragma unroll
            for (int i = 0; i < DWF_SEG; ++i) acc[i] = hs;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float2 w = ws[k * P + pr];
#pragma unroll
                for (int i = 0; i < DWF_SEG; ++i) {
#ifndef COCR_DW_SCALAR
                    acc[i] = dw_ffma2(w, xw[i + k], acc[i]);
#else
                    acc[i].x = fmaf(w.x, xw[i + k].x, acc[i].x);
                    acc[i].y = fmaf(w.y, xw[i + k].y, acc[i].y);
#endif
                }
            }
#pragma unroll
            for (int i = 0; i < DWF_SEG; ++i) {
                const int t = t0 + seg * DWF_SEG + i;
                if (t < T) {
                    const float2 hh = acc[i];
                    const float2 r = dw_ffma2(hh, make_float2(dw_tanh(hh.x), dw_tanh(hh.y)), hh);
                    *reinterpret_cast<uint32_t *>(g + ((int64_t)b * T + t) * d + 2 * pr) = f2_to_bf2(r.x, r.y);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Second generation of the same kernel for the default spec (d = 144, K = 31), built to spend its issue slots on FFMA2:
// the first one ran 112 instructions per output pair (31 of them FFMA2; ncu: issue slots 54 % busy, DRAM 8 %) because every
// shared-memory index carried a runtime row pitch and a thread produced only 8 outputs per 38 inputs it unpacked.  Here
// the geometry is compile time (all shared-memory offsets are immediates), a thread owns 16 consecutive frames of a
// channel pair and slides a 16-deep register window over the 46 input rows it needs: per tap one LDS.64 (taps), one LDS.32
// + two ALU ops (new input row) and 16 FFMA2.  Tiles of 96 frames: T = 384 gives 4 tiles per line, 1024 items = 6.92 per SM.
// ------------------------------------------------------------------------------------------------------------
constexpr int DW2_D = 144, DW2_P = DW2_D / 2, DW2_K = 31, DW2_SEG = 16, DW2_TT = 96;
constexpr int DW2_ROWS = DW2_TT + DW2_K - 1, DW2_THREADS = DW2_P * (DW2_TT / DW2_SEG);
constexpr size_t DW2_SMEM = (size_t)2 * DW2_ROWS * DW2_P * 4 + (size_t)DW2_K * DW2_P * 8;

__global__ void __launch_bounds__(DW2_THREADS, 1)
k_dw_v2(const bf16 *__restrict__ x, const float2 *__restrict__ taps /* [K][P] w*scale/2 */, const float2 *__restrict__ hshift /* [P] */,
        bf16 *__restrict__ g, int T, int n_tt, int n_items) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    constexpr int P = DW2_P, K = DW2_K, ROWS = DW2_ROWS, pad = (DW2_K - 1) / 2, C8 = DW2_D / 8;
    float2 *ws = reinterpret_cast<float2 *>(sm_raw);                    // [K][P]
    uint32_t *xs0 = reinterpret_cast<uint32_t *>(ws + K * P);           // 2 x [ROWS][P] bf16x2
    const int st_row0 = threadIdx.x / C8, st_c8 = threadIdx.x - st_row0 * C8;
    constexpr int st_step = DW2_THREADS / C8;                          // 24 rows per pass
    static_assert(DW2_THREADS % C8 == 0, "staging map needs whole rows per pass");
    auto stage = [&](int item, int buf) {
        const int b = item / n_tt, t0 = (item - b * n_tt) * DW2_TT;
        uint32_t *xs = xs0 + buf * ROWS * P;
        const bf16 *src = x + (int64_t)b * T * DW2_D + st_c8 * 8;
#pragma unroll
        for (int i = 0; i < (ROWS + st_step - 1) / st_step; ++i) {
            const int row = st_row0 + i * st_step;
            if (row < ROWS) {
                const int t = t0 - pad + row;
                const bool ok = t >= 0 && t < T;
                cp_async16(xs + row * P + st_c8 * 4, src + (int64_t)(ok ? t : 0) * DW2_D, ok);
            }
        }
        cp_async_commit();
    };
    int item = blockIdx.x;
    if (item < n_items) stage(item, 0);
    for (int e = threadIdx.x; e < K * P; e += DW2_THREADS) ws[e] = taps[e];
    const int pr = threadIdx.x % P, seg = threadIdx.x / P;
    const float2 hs = hshift[pr];
    int buf = 0;
    for (; item < n_items; item += gridDim.x, buf ^= 1) {
        cp_async_wait<0>();
        __syncthreads();                                    // rows of `item` visible; everyone is done with the other buffer
        if (item + (int)gridDim.x < n_items) stage(item + gridDim.x, buf ^ 1);
        const int b = item / n_tt, t0 = (item - b * n_tt) * DW2_TT;
        const uint32_t *xs = xs0 + buf * ROWS * P + (seg * DW2_SEG) * P + pr;      // row 0 of this thread's window
        float2 win[DW2_SEG], acc[DW2_SEG];
#pragma unroll
        for (int i = 0; i < DW2_SEG - 1; ++i) win[i] = bf2_to_f2(xs[i * P]);
#pragma unroll
        for (int i = 0; i < DW2_SEG; ++i) acc[i] = hs;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            win[(k + DW2_SEG - 1) % DW2_SEG] = bf2_to_f2(xs[(k + DW2_SEG - 1) * P]);
            const float2 w = ws[k * P + pr];
#pragma unroll
            for (int i = 0; i < DW2_SEG; ++i) acc[i] = dw_ffma2(w, win[(i + k) % DW2_SEG], acc[i]);
        }
        bf16 *dst = g + ((int64_t)b * T + t0 + seg * DW2_SEG) * DW2_D + 2 * pr;
        const int nvalid = T - (t0 + seg * DW2_SEG);
#pragma unroll
        for (int i = 0; i < DW2_SEG; ++i) {
            if (i < nvalid) {
                const float2 hh = acc[i];
                const float2 r = dw_ffma2(hh, make_float2(dw_tanh(hh.x), dw_tanh(hh.y)), hh);
                *reinterpret_cast<uint32_t *>(dst + (int64_t)i * DW2_D) = f2_to_bf2(r.x, r.y);
            }
        }
    }
}

// taps_dev: [K][P] float2 = w * bn_scale / 2 ; hshift_dev: [P] float2 = bn_shift / 2  (packed once per layer at load)
int launch_dw_fast(cocr_handle *h, const bf16 *x, const float *taps_dev, const float *hshift_dev, bf16 *g, int B, int T, cudaStream_t st) {
    const Dims &dm = h->dm;
    static const bool v1 = getenv("COCR_DW_V1") != nullptr;
    if (dm.K == DW2_K && dm.d == DW2_D && taps_dev && !v1) {
        static PerDevice<bool> attr2;
        if (!attr2.get()) {
            COCR_CUDA(h, cudaFuncSetAttribute(k_dw_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW2_SMEM));
            attr2.get() = true;
        }
        const int n_tt = (T + DW2_TT - 1) / DW2_TT, n_items = n_tt * B;
        const int grid = n_items < h->num_sms ? n_items : h->num_sms;
        ProfScope ps(h, "dwconv_bn_silu_bf16", st, 2.0 * 31 * B * (double)T * dm.d, 4.0 * B * T * dm.d);
        k_dw_v2<<<grid, DW2_THREADS, DW2_SMEM, st>>>(x, reinterpret_cast<const float2 *>(taps_dev), reinterpret_cast<const float2 *>(hshift_dev), g, T,
                                                     n_tt, n_items);
        COCR_LAUNCHED(h);
        return COCR_OK;
    }
    const int P = dm.d / 2;
    const int nthr = P * (DWF_TT / DWF_SEG);
    if (dm.K != 31 || (dm.d & 7) != 0 || nthr > 576 || !taps_dev || nthr % (dm.d / 8) != 0) return COCR_ERR_UNSUPPORTED;
    const size_t smem = (size_t)2 * (DWF_TT + 30) * P * 4 + (size_t)31 * P * 8;
    static PerDevice<bool> attr;
    if (!attr.get()) {
        COCR_CUDA(h, cudaFuncSetAttribute(k_dw_fast<31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.get() = true;
    }
    const int n_tt = (T + DWF_TT - 1) / DWF_TT, n_items = n_tt * B;
    int grid = 2 * h->num_sms;
    if (grid > n_items) grid = n_items;
    ProfScope ps(h, "dwconv_bn_silu_bf16", st, 2.0 * 31 * B * (double)T * dm.d, 4.0 * B * T * dm.d);
    k_dw_fast<31><<<grid, nthr, smem, st>>>(x, reinterpret_cast<const float2 *>(taps_dev), reinterpret_cast<const float2 *>(hshift_dev), g, T, dm.d,
                                            n_tt, n_items);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------------------------
// fused subsampling front-end for 32 conv channels (lane = channel).  One CTA = FE_TT output frames of one line.
//   stage A: input patch (H x (4*TT+3)) -> shared (zero padded)
//   stage B: conv1+ReLU  -> a1 tile [(2*TT+1)][F2][32] bf16 in shared; warp = one t2 row, lane = channel,
//            input row broadcast from shared with 16-byte loads, 9 weights per lane in registers
//   stage C: depthwise 3x3 s2 -> a2 [TT][F][32] fp32 in shared
//   stage D: 1x1 conv 32->32 + ReLU, weights Wpw[lane][0..31] in registers, a2 broadcast; writes bf16 features
// ------------------------------------------------------------------------------------------------------------
constexpr int FE_TT = 8;
constexpr int FE_THREADS = 256;

__global__ void __launch_bounds__(FE_THREADS)
k_frontend_fused(const float *__restrict__ img, int64_t sb, int64_t sw, int64_t sh, const float *__restrict__ w1,
                 const float *__restrict__ b1, const float *__restrict__ wdw, const float *__restrict__ bdw,
                 const float *__restrict__ wpw, const float *__restrict__ bpw, bf16 *__restrict__ feat, int W, int H, int T2,
                 int T, int F2, int F) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    constexpr int NW = 4 * FE_TT + 3;               // input columns (time) needed
    constexpr int NT2 = 2 * FE_TT + 1;              // conv1 rows needed
    const int HP = (H + 2 + 3) & ~3;                // padded input row: y = -1 .. H, rounded to 4 floats
    float *patch = reinterpret_cast<float *>(sm_raw);                     // [NW][HP], patch[wl][y+1]
    bf16 *a1 = reinterpret_cast<bf16 *>(patch + NW * HP);                 // [NT2][F2][32]
    float *a2 = reinterpret_cast<float *>(a1 + (size_t)NT2 * F2 * 32);    // [TT][F][32]
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * FE_TT;
    const int w0 = 4 * t0 - 3;                      // first input column of the patch
    const int t20 = 2 * t0 - 1;                     // first conv1 row of the tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // stage A
    for (int e = threadIdx.x; e < NW * HP; e += FE_THREADS) {
        int wl, y;
        if (sw == 1) { y = e / NW; wl = e - y * NW; }           // (N,1,H,W) layout: time is contiguous
        else { wl = e / HP; y = e - wl * HP; }                  // (B,W,H) layout: height is contiguous
        const int w = w0 + wl, yy = y - 1;
        float v = 0.0f;
        if (w >= 0 && w < W && yy >= 0 && yy < H) v = img[b * sb + w * sw + yy * sh];
        patch[wl * HP + y] = v;
    }
    __syncthreads();

    // stage B: conv1 (convolution.py:192-198); conv "height" = time, conv "width" = image height
    {
        float wr[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) wr[i] = w1[lane * 9 + i];
        const float bias = b1[lane];
        for (int r = warp; r < NT2; r += FE_THREADS / 32) {
            const int t2 = t20 + r;
            bf16 *dst = a1 + (size_t)r * F2 * 32 + lane;
            if (t2 < 0 || t2 >= T2) {                            // outside the conv1 output: zero padding of the dw conv
                for (int f2 = 0; f2 < F2; ++f2) dst[f2 * 32] = __float2bfloat16_rn(0.0f);
                continue;
            }
            // local input columns of this row: wl = 2*r + kh  (w = 2*t2 - 1 + kh)
            const float *p0 = patch + (2 * r) * HP, *p1 = p0 + HP, *p2 = p1 + HP;
            for (int f2 = 0; f2 < F2; f2 += 2) {                 // two outputs per iteration share one 16-byte window
                // y indices (padded by +1): outputs f2, f2+1 use y' = 2*f2 .. 2*f2+4
                const float4 q0 = *reinterpret_cast<const float4 *>(p0 + 2 * f2);
                const float4 q1 = *reinterpret_cast<const float4 *>(p1 + 2 * f2);
                const float4 q2 = *reinterpret_cast<const float4 *>(p2 + 2 * f2);
                const float e0 = p0[2 * f2 + 4], e1 = p1[2 * f2 + 4], e2 = p2[2 * f2 + 4];
                float o0 = bias, o1 = bias;
                o0 = fmaf(wr[0], q0.x, o0); o0 = fmaf(wr[1], q0.y, o0); o0 = fmaf(wr[2], q0.z, o0);
                o0 = fmaf(wr[3], q1.x, o0); o0 = fmaf(wr[4], q1.y, o0); o0 = fmaf(wr[5], q1.z, o0);
                o0 = fmaf(wr[6], q2.x, o0); o0 = fmaf(wr[7], q2.y, o0); o0 = fmaf(wr[8], q2.z, o0);
                o1 = fmaf(wr[0], q0.z, o1); o1 = fmaf(wr[1], q0.w, o1); o1 = fmaf(wr[2], e0, o1);
                o1 = fmaf(wr[3], q1.z, o1); o1 = fmaf(wr[4], q1.w, o1); o1 = fmaf(wr[5], e1, o1);
                o1 = fmaf(wr[6], q2.z, o1); o1 = fmaf(wr[7], q2.w, o1); o1 = fmaf(wr[8], e2, o1);
                dst[f2 * 32] = __float2bfloat16_rn(fmaxf(o0, 0.0f));
                if (f2 + 1 < F2) dst[(f2 + 1) * 32] = __float2bfloat16_rn(fmaxf(o1, 0.0f));
            }
        }
    }
    __syncthreads();

    // stage C: depthwise conv (convolution.py:200-205): a2[t][f][c] = b + sum a1[2t-1+kh][2f-1+kw][c] w[c][kh][kw]
    {
        float wr[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) wr[i] = wdw[lane * 9 + i];
        const float bias = bdw[lane];
        for (int tl = warp; tl < FE_TT; tl += FE_THREADS / 32) {
            for (int f = 0; f < F; ++f) {
                float o = bias;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int r = 2 * tl + kh;                   // local conv1 row (t2 = 2t-1+kh, t20 = 2*t0-1)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int f2 = 2 * f - 1 + kw;
                        if (f2 >= 0 && f2 < F2) o = fmaf(wr[kh * 3 + kw], __bfloat162float(a1[((size_t)r * F2 + f2) * 32 + lane]), o);
                    }
                }
                a2[(tl * F + f) * 32 + lane] = o;
            }
        }
    }
    __syncthreads();

    // stage D: 1x1 conv + ReLU (convolution.py:207-213) and the (t, f*32+c) feature layout the projection consumes
    {
        float wr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) wr[i] = wpw[lane * 32 + i];
        const float bias = bpw[lane];
        for (int e = warp; e < FE_TT * F; e += FE_THREADS / 32) {
            const int tl = e / F, f = e - tl * F;
            const int t = t0 + tl;
            if (t >= T) continue;
            const float4 *src = reinterpret_cast<const float4 *>(a2 + (tl * F + f) * 32);
            float o = bias;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = src[i];
                o = fmaf(wr[4 * i + 0], v.x, o); o = fmaf(wr[4 * i + 1], v.y, o);
                o = fmaf(wr[4 * i + 2], v.z, o); o = fmaf(wr[4 * i + 3], v.w, o);
            }
            feat[((int64_t)b * T + t) * (F * 32) + f * 32 + lane] = __float2bfloat16_rn(fmaxf(o, 0.0f));
        }
    }
}

int launch_frontend_fused(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, int B, int W, bf16 *feat,
                          cudaStream_t st) {
    const Dims &dm = h->dm;
    if (dm.cc != 32 || (dm.F2 & 1) != 0) return COCR_ERR_UNSUPPORTED;
    const int T2 = out_len1(W), T = out_len1(T2);
    const int HP = (dm.H + 2 + 3) & ~3;
    const size_t smem = (size_t)(4 * FE_TT + 3) * HP * 4 + (size_t)(2 * FE_TT + 1) * dm.F2 * 32 * 2 + (size_t)FE_TT * dm.F * 32 * 4;
    if (smem > 200 * 1024) return COCR_ERR_UNSUPPORTED;
    static PerDevice<size_t> attr;
    if (smem > attr.get()) {
        COCR_CUDA(h, cudaFuncSetAttribute(k_frontend_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.get() = smem;
    }
    dim3 grid((T + FE_TT - 1) / FE_TT, B);
    const double n1 = (double)B * T2 * dm.F2 * 32, n2 = (double)B * T * dm.F * 32;
    ProfScope ps(h, "frontend_fused", st, 18.0 * n1 + 18.0 * n2 + 64.0 * n2, 4.0 * B * (double)W * dm.H + 2.0 * n2);
    k_frontend_fused<<<grid, FE_THREADS, smem, st>>>(img, sb, sw, sh, h->sub_w1, h->sub_b1, h->sub_wdw, h->sub_bdw, h->sub_wpw,
                                                    h->sub_bpw, feat, W, dm.H, T2, T, dm.F2, dm.F);
    COCR_LAUNCHED(h);
    return COCR_OK;
}


// ------------------------------------------------------------------------------------------------------------
// Front-end, version 2 (32 conv channels): conv1 3x3 s2 + ReLU and the depthwise 3x3 s2 conv on CUDA cores with
// the weights as immediate/constant operands, the 1x1 conv + ReLU as a tcgen05 GEMM over the (frame, f) rows.
//   * the weights travel in the kernel parameter block (constant bank): every FFMA takes its weight as a constant
//     operand, so the inner loops are pure FFMA streams with no weight loads;
//   * conv1: thread = one (t2, f2) position, all 32 channels in registers (9 input loads, 288 FFMA);
//   * depthwise: thread = one (t, f) position x 8 channels (warp-uniform channel group), 16-byte a1 reads;
//   * a2 leaves as bf16 rows [(b, t, f)][32 channels] - exactly the row-major (t, f*32 + c) feature layout - and the
//     1x1 conv is  feat = relu(a2 * Wpw^T + b)  on tensor cores (k_gemm_tc, EPI_BF16 with relu).
// ------------------------------------------------------------------------------------------------------------
struct FrontendWeights {
    float w1[32 * 9], b1[32], wdw[32 * 9], bdw[32];
};

constexpr int FE2_TT = 12;           // 25 x 48 conv1 positions = 4.7 passes of 256 threads (8 frames: 3.2 passes, the last one 81 % idle)
constexpr int FE2_THREADS = 256;

// 16-byte chunk g (8 channels) of position pos inside a1[pos][32 ch] bf16, xor-swizzled so that lanes writing / reading
// consecutive positions spread over the banks
__device__ __forceinline__ uint32_t a1_offset(int pos, int g) { return (uint32_t)pos * 64u + (uint32_t)((g ^ (pos >> 1)) & 3) * 16u; }

template <int G>
__device__ __forceinline__ void dw_group(const FrontendWeights &fw, const uint8_t *a1, int tl, int f, int F2, float (&o)[8]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = fw.bdw[G * 8 + c];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int r = 2 * tl + kh;                               // local conv1 row (t2 = 2t - 1 + kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int f2 = 2 * f - 1 + kw;
            if (f2 >= 0 && f2 < F2) {
                const uint4 v = *reinterpret_cast<const uint4 *>(a1 + a1_offset(r * F2 + f2, G));
                const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) {
                    const float2 x = bf2_to_f2(u[c2]);
                    o[2 * c2] = fmaf(fw.wdw[(G * 8 + 2 * c2) * 9 + kh * 3 + kw], x.x, o[2 * c2]);
                    o[2 * c2 + 1] = fmaf(fw.wdw[(G * 8 + 2 * c2 + 1) * 9 + kh * 3 + kw], x.y, o[2 * c2 + 1]);
                }
            }
        }
    }
}

// PixT = float: the reference's input (pixels already scaled to [0, 1]); PixT = uint8_t: raw 8-bit pixels, scaled here by the
// correctly rounded p / 255 of torchvision's ToTensor, so both inputs give bit-identical results.
template <typename PixT>
__global__ void __launch_bounds__(FE2_THREADS, 2)
k_frontend_convs2(const PixT *__restrict__ img, int64_t sb, int64_t sw, int64_t sh, const __grid_constant__ FrontendWeights fw,
                  bf16 *__restrict__ a2, int W, int H, int T2, int T, int F2, int F) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    constexpr int NW = 4 * FE2_TT + 3;              // input columns (time) needed
    constexpr int NT2 = 2 * FE2_TT + 1;             // conv1 rows needed
    const int HP = (H + 2 + 3) & ~3;                // padded input row: y = -1 .. H, rounded to 4 floats
    float *patch = reinterpret_cast<float *>(sm_raw);                     // [NW][HP], patch[wl][y+1]
    uint8_t *a1 = reinterpret_cast<uint8_t *>(patch + NW * HP);           // [NT2 * F2 positions][64 B], swizzled chunks
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * FE2_TT;
    const int w0 = 4 * t0 - 3;                      // first input column of the patch
    const int t20 = 2 * t0 - 1;                     // first conv1 row of the tile

    // stage A: input patch (zero padded)
    for (int e = threadIdx.x; e < NW * HP; e += FE2_THREADS) {
        int wl, y;
        if (sw == 1) { y = e / NW; wl = e - y * NW; }           // (N,1,H,W) layout: time is contiguous
        else { wl = e / HP; y = e - wl * HP; }                  // (B,W,H) layout: height is contiguous
        const int w = w0 + wl, yy = y - 1;
        float v = 0.0f;
        if (w >= 0 && w < W && yy >= 0 && yy < H) {
            if constexpr (sizeof(PixT) == 1) v = __fdiv_rn((float)img[b * sb + w * sw + yy * sh], 255.0f);
            else v = img[b * sb + w * sw + yy * sh];
        }
        patch[wl * HP + y] = v;
    }
    __syncthreads();

    // stage B: conv1 + ReLU (convolution.py:192-198); conv "height" = time, conv "width" = image height
    for (int pos = threadIdx.x; pos < NT2 * F2; pos += FE2_THREADS) {
        const int r = pos / F2, f2 = pos - r * F2;
        const int t2 = t20 + r;
        uint32_t pk[16];
        if (t2 < 0 || t2 >= T2) {                               // outside the conv1 output: zero padding of the dw conv
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
        } else {
            float in[9];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) in[kh * 3 + kw] = patch[(2 * r + kh) * HP + 2 * f2 + kw];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                float o0 = fw.b1[c], o1 = fw.b1[c + 1];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    o0 = fmaf(fw.w1[c * 9 + k], in[k], o0);
                    o1 = fmaf(fw.w1[(c + 1) * 9 + k], in[k], o1);
                }
                pk[c >> 1] = f2_to_bf2(fmaxf(o0, 0.0f), fmaxf(o1, 0.0f));
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4 *>(a1 + a1_offset(pos, g)) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
    }
    __syncthreads();

    // stage C: depthwise conv (convolution.py:200-205): a2[t][f][c] = b + sum a1[2t-1+kh][2f-1+kw][c] w[c][kh][kw]
    const int npos = FE2_TT * F;                                  // multiple of 32 for F = 24: a warp stays inside one group
    for (int item = threadIdx.x; item < 4 * npos; item += FE2_THREADS) {
        const int g = item / npos, p2 = item - g * npos;
        const int tl = p2 / F, f = p2 - tl * F;
        const int t = t0 + tl;
        float o[8];
        switch (g) {
            case 0: dw_group<0>(fw, a1, tl, f, F2, o); break;
            case 1: dw_group<1>(fw, a1, tl, f, F2, o); break;
            case 2: dw_group<2>(fw, a1, tl, f, F2, o); break;
            default: dw_group<3>(fw, a1, tl, f, F2, o); break;
        }
        if (t < T) {
            // a2 row = [hi (32 bf16) | lo (32 bf16)], hi + lo = the fp32 value to 16 mantissa bits: the 1x1 conv GEMM runs
            // with K = 64 against [Wpw | Wpw], so the activation side of that GEMM is not rounded to 8 bits
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) {
                hi[c2] = f2_to_bf2(o[2 * c2], o[2 * c2 + 1]);
                const float2 hf = bf2_to_f2(hi[c2]);
                lo[c2] = f2_to_bf2(o[2 * c2] - hf.x, o[2 * c2 + 1] - hf.y);
            }
            bf16 *row = a2 + (((int64_t)b * T + t) * F + f) * 64 + g * 8;
            *reinterpret_cast<uint4 *>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(row + 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// conv1 + depthwise conv -> a2 (B*T*F, 64) bf16 [hi | lo]; the caller runs the 1x1 conv as a K = 64 GEMM.  COCR_ERR_UNSUPPORTED when the
// geometry is not covered.
int launch_frontend_convs2(cocr_handle *h, const void *img_any, bool u8, int64_t sb, int64_t sw, int64_t sh, int B, int W, bf16 *a2, cudaStream_t st) {
    const float *img = static_cast<const float *>(img_any);     // fp32 pixels unless `u8` (then uint8 pixels)
    const Dims &dm = h->dm;
    if (dm.cc != 32 || ((FE2_TT * dm.F) & 31) != 0 || h->fe_weights.size() != sizeof(FrontendWeights) / 4) return COCR_ERR_UNSUPPORTED;
    const int T2 = out_len1(W), T = out_len1(T2);
    const int HP = (dm.H + 2 + 3) & ~3;
    const size_t smem = (size_t)(4 * FE2_TT + 3) * HP * 4 + (size_t)(2 * FE2_TT + 1) * dm.F2 * 64;
    if (smem > 200 * 1024) return COCR_ERR_UNSUPPORTED;
    static PerDevice<size_t> attr;
    if (smem > attr.get()) {
        COCR_CUDA(h, cudaFuncSetAttribute(k_frontend_convs2<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        COCR_CUDA(h, cudaFuncSetAttribute(k_frontend_convs2<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.get() = smem;
    }
    FrontendWeights fw;
    memcpy(&fw, h->fe_weights.data(), sizeof(fw));
    dim3 grid((T + FE2_TT - 1) / FE2_TT, B);
    const double n1 = (double)B * T2 * dm.F2 * 32, n2 = (double)B * T * dm.F * 32;
    ProfScope ps(h, "frontend_convs", st, 18.0 * n1 + 18.0 * n2, (u8 ? 1.0 : 4.0) * B * (double)W * dm.H + 2.0 * n2);
    if (u8) k_frontend_convs2<uint8_t><<<grid, FE2_THREADS, smem, st>>>(static_cast<const uint8_t *>(img_any), sb, sw, sh, fw, a2, W, dm.H, T2, T, dm.F2, dm.F);
    else k_frontend_convs2<float><<<grid, FE2_THREADS, smem, st>>>(img, sb, sw, sh, fw, a2, W, dm.H, T2, T, dm.F2, dm.F);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

}  // namespace cocr
