// Convolutional sub-sampling front-end on tensor cores (convolution.py:190-216,227-236), default geometry (height 96,
// 32 conv channels):  image -> Conv2d(1->32, 3x3, s2, p1) + ReLU -> depthwise Conv2d(3x3, s2, p1) -> Conv2d 1x1 32->32 + ReLU
// -> feature rows (frame, f*32 + c) bf16 for the 768->144 projection.  One kernel; nothing but the image is read from HBM
// and nothing but the feature matrix is written (the second-generation front-end wrote the depthwise output as [hi|lo]
// rows, 302 MB per batch, and a separate GEMM read them back: 1 086 MB moved for 236 MB algorithmic).
//
// A persistent CTA (two per SM) walks over items = (line, 8 output frames).  Per item, in bulk-synchronous phases:
//   1  the zero-padded input patch (35 columns x 98 rows, fp32) -> shared memory
//   2  im2col of conv1: one 64-byte row per conv1 position (17 x 48 = 816 positions): K = 32 bf16 =
//      [hi(9 taps) | lo(9) | hi(9) | 0 x 5], hi + lo = the fp32 pixel to 16 mantissa bits (SWIZZLE_64B K-major A operand)
//   M1 7 x 2 tcgen05.mma (128 x 32 x 16) against B1 = [w_hi | w_hi | w_lo | 0] per channel: hi*w_hi + lo*w_hi + hi*w_lo,
//      i.e. conv1 to ~2^-16 relative accuracy with fp32 accumulation in TMEM (224 columns)
//      (the bias rides in two more columns of K; rows outside the conv1 output are all zero: the depthwise conv's zero padding)
//   3  TMEM -> registers: bf16, packed ReLU -> a1[position][32 channels] in shared memory
//   4  depthwise 3x3 s2 on CUDA cores (weights as constant operands) -> [hi(32) | lo(32)] rows of the 1x1 conv's A operand
//      (192 rows, SWIZZLE_128B)
//   M2 2 x 4 tcgen05.mma (128 x 32 x 16) against B2 = [Wpw | Wpw]
//   5  TMEM -> registers: + bias, ReLU, bf16 -> 64-byte feature rows in HBM (consecutive threads = consecutive rows)
// Shared memory: the im2col rows (56 KB) are dead once M1 has completed and are reused for a1; the patch (14 KB) is dead
// after phase 2 and is reused for the 1x1 conv's A operand (32 KB): 94 KB per CTA, two CTAs per SM.  TMEM: 256 columns.
#include <string.h>

#include "common.h"
#include "tc_utils.cuh"

namespace cocr {

typedef __nv_bfloat16 bf16;

struct FrontendWeights3 {
    float w1[32 * 9], b1[32], wdw[32 * 9], bdw[32];     // same order as mem_kernels.cu: FrontendWeights / cocr_handle::fe_weights
};

constexpr int FE3_TT = 8;                              // output frames per item
constexpr int FE3_H = 96, FE3_F2 = 48, FE3_F = 24;
constexpr int FE3_NW = 4 * FE3_TT + 3;                 // input columns (time) of the patch
constexpr int FE3_NT2 = 2 * FE3_TT + 1;                // conv1 rows
constexpr int FE3_NPOS = FE3_NT2 * FE3_F2;             // 816 conv1 positions
constexpr int FE3_MT1 = (FE3_NPOS + 127) / 128;        // 7 M tiles of conv1
constexpr int FE3_ROWS2 = FE3_TT * FE3_F;              // 192 rows of the 1x1 conv
constexpr int FE3_MT2 = (FE3_ROWS2 + 127) / 128;       // 2 M tiles
constexpr int FE3_HP = 100;                            // patch row: y = -1 .. 98 (96 rows + zero padding), multiple of 4 floats
constexpr int FE3_THREADS = 256;
constexpr uint32_t FE3_A1_BYTES = FE3_MT1 * 128 * 64;  // 57 344: im2col rows, later a1 (816 x 64 B)
constexpr uint32_t FE3_Y_BYTES = FE3_MT2 * 128 * 128;  // 32 768: A operand of the 1x1 conv, earlier the patch (35 x 100 x 4 = 14 000)
constexpr uint32_t FE3_B1_BYTES = 32 * 64, FE3_B2_BYTES = 32 * 128;
constexpr uint32_t FE3_SMEM = FE3_A1_BYTES + FE3_Y_BYTES + FE3_B1_BYTES + FE3_B2_BYTES + 2 * 32 * 4 + 64;
constexpr uint32_t FE3_TMEM_COLS = 256;
static_assert(FE3_NW * FE3_HP * 4 <= FE3_Y_BYTES, "patch must fit the region it shares");
static_assert(FE3_NPOS * 64 <= FE3_A1_BYTES && FE3_MT1 * 32 <= FE3_TMEM_COLS, "conv1 tiles");
static_assert(FE3_ROWS2 % 64 == 0 && FE3_THREADS == 256, "phase 4 deals 64 positions to each pair of warps");

constexpr uint64_t kLayoutSw64 = 4;
__device__ __forceinline__ uint64_t fe3_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint64_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}
// byte offset of (row, 16-byte chunk 0..3) in a [rows][64 B] SWIZZLE_64B tile (512-byte aligned): address bits 4-5 ^= bits 7-8
__device__ __forceinline__ uint32_t sw64_offset(uint32_t row, uint32_t chunk) { return row * 64u + (((chunk ^ (row >> 1)) & 3u) << 4); }
__device__ __forceinline__ uint32_t fe3_a1_offset(int pos, int g) { return (uint32_t)pos * 64u + (uint32_t)((g ^ (pos >> 1)) & 3) * 16u; }
__device__ __forceinline__ float2 fe3_bf2_to_f2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u)); }
__device__ __forceinline__ uint32_t fe3_pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float fe3_bf_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// depthwise conv of NP (frame, f) positions for channel group G (8 channels).  The weights are constant operands fetched
// once per tap and used for all NP positions (with one position per weight fetch the loop spent as many issue slots on
// uniform constant loads as on FFMAs).  The only tap that can leave the conv1 output is kw = 0 at f = 0 (f2 = -1): its
// operand is replaced by zero with a select instead of a branch.
template <int G, int NP>
__device__ __forceinline__ void fe3_dw_group(const FrontendWeights3 &fw, const uint8_t *a1, const int (&tl)[NP], const int (&f)[NP],
                                             float (&o)[NP][8]) {
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) o[p][c] = fw.bdw[G * 8 + c];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            uint4 v[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int f2 = 2 * f[p] - 1 + kw;
                const int pos = (2 * tl[p] + kh) * FE3_F2 + (kw == 0 ? max(f2, 0) : f2);      // local conv1 row 2t - 1 + kh
                v[p] = *reinterpret_cast<const uint4 *>(a1 + fe3_a1_offset(pos, G));
                if (kw == 0 && f2 < 0) v[p] = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) {
                const float w0 = fw.wdw[(G * 8 + 2 * c2) * 9 + kh * 3 + kw], w1 = fw.wdw[(G * 8 + 2 * c2 + 1) * 9 + kh * 3 + kw];
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const uint32_t u = c2 == 0 ? v[p].x : (c2 == 1 ? v[p].y : (c2 == 2 ? v[p].z : v[p].w));
                    const float2 x = fe3_bf2_to_f2(u);
                    o[p][2 * c2] = fmaf(w0, x.x, o[p][2 * c2]);
                    o[p][2 * c2 + 1] = fmaf(w1, x.y, o[p][2 * c2 + 1]);
                }
            }
        }
    }
}

template <typename PixT>
__global__ void __launch_bounds__(FE3_THREADS, 2)
k_frontend_tc(const PixT *__restrict__ img, int64_t sb, int64_t sw, int64_t sh, const __grid_constant__ FrontendWeights3 fw,
              const bf16 *__restrict__ wpw2 /* (32, 64) = [Wpw | Wpw] */, const float *__restrict__ bpw, bf16 *__restrict__ feat, int W,
              int T2, int T, int n_tt, int n_items) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA1 = smem;                                        // im2col rows | a1
    uint8_t *sY = smem + FE3_A1_BYTES;                          // patch | A operand of the 1x1 conv
    uint8_t *sB1 = sY + FE3_Y_BYTES;
    uint8_t *sB2 = sB1 + FE3_B1_BYTES;
    float *sbias = reinterpret_cast<float *>(sB2 + FE3_B2_BYTES);   // [32] conv1 bias | [32] 1x1 conv bias
    uint64_t *bar = reinterpret_cast<uint64_t *>(sbias + 64);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    float *patch = reinterpret_cast<float *>(sY);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0u) __trap();
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, FE3_TMEM_COLS);
    // B operands, written once: B1 row c = [w_hi(9) | w_hi(9) | w_lo(9) | 0 x 5] (SWIZZLE_64B), B2 = [Wpw | Wpw] rows (SWIZZLE_128B)
    if (threadIdx.x < 32) {
        const int c = threadIdx.x;
        __align__(16) bf16 row[32];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float w = fw.w1[c * 9 + k];
            const bf16 hi = __float2bfloat16_rn(w);
            row[k] = hi; row[9 + k] = hi; row[18 + k] = __float2bfloat16_rn(w - __bfloat162float(hi));
        }
        // the conv1 bias rides in the GEMM: im2col columns 27, 28 hold 1.0 and B1 holds the bias split into bf16 hi + lo there
        {
            const float bb = fw.b1[c];
            const bf16 bh = __float2bfloat16_rn(bb);
            row[27] = bh; row[28] = __float2bfloat16_rn(bb - __bfloat162float(bh));
        }
#pragma unroll
        for (int k = 29; k < 32; ++k) row[k] = __float2bfloat16_rn(0.0f);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4 *>(sB1 + sw64_offset(c, ch)) = *reinterpret_cast<const uint4 *>(row + 8 * ch);
        sbias[c] = fw.b1[c];
        sbias[32 + c] = bpw[c];
    }
    {
        const int r = threadIdx.x >> 3, ch = threadIdx.x & 7;                       // 32 rows x 8 chunks of 16 bytes
        *reinterpret_cast<uint4 *>(sB2 + sw128_offset(r, ch)) = *reinterpret_cast<const uint4 *>(wpw2 + r * 64 + ch * 8);
    }
    // im2col rows beyond the last position (816 .. 895) are multiplied like the others; keep them finite
    for (uint32_t o = FE3_NPOS * 64 + threadIdx.x * 16; o < FE3_A1_BYTES; o += FE3_THREADS * 16) *reinterpret_cast<uint4 *>(sA1 + o) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = make_idesc_bf16(128, 32);
    const int q = warp & 3, half = warp >> 2;                  // TMEM lane quarter of this warp, which of two M tiles it drains
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t n_done = 0;                                       // completions of `bar` so far

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / n_tt, t0 = (item - b * n_tt) * FE3_TT;
        const int w0 = 4 * t0 - 3;                             // first input column of the patch
        const int t20 = 2 * t0 - 1;                            // first conv1 row of the item

        // ---- phase 1: input patch (zero padded), patch[wl][y + 1].  All loads of a thread are issued before the first store:
        // the phase is bound by global-load latency, so what counts is how many loads are in flight (ncu: with a
        // load -> store loop this phase held 45 % of the kernel's stall samples) ----
        if (sw == 1) {
            // (N,1,H,W) layout, time contiguous: a warp takes whole image rows (lanes = 32 consecutive columns + 3 more), so the
            // address arithmetic is one multiply per row instead of two divisions per element
            constexpr int NR = (FE3_HP - 2 + 7) / 8;                      // rows y = 0 .. 97 dealt to 8 warps
            float v0[NR], v1[NR];
            const int wa = w0 + lane, wb = w0 + 32 + lane;
            const bool oka = wa >= 0 && wa < W, okb = lane < FE3_NW - 32 && wb >= 0 && wb < W;
            const PixT *pa = img + b * sb + (int64_t)(warp - 1) * sh + wa;      // row yy = warp - 1 + 8 i: one pointer bump per row
            const int64_t step = 8 * sh;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int yy = warp - 1 + 8 * i;
                const bool row_ok = yy >= 0 && yy < FE3_H;
                v0[i] = (row_ok && oka) ? (float)pa[0] : 0.0f;
                v1[i] = (row_ok && okb) ? (float)pa[32] : 0.0f;
                pa += step;
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int y = warp + 8 * i;
                if (y < FE3_HP - 2) {
                    float a = v0[i], c = v1[i];
                    if constexpr (sizeof(PixT) == 1) { a = __fdiv_rn(a, 255.0f); c = __fdiv_rn(c, 255.0f); }
                    patch[lane * FE3_HP + y] = a;
                    if (lane < FE3_NW - 32) patch[(32 + lane) * FE3_HP + y] = c;
                }
            }
        } else {
            constexpr int NE = (FE3_NW * FE3_HP + FE3_THREADS - 1) / FE3_THREADS;      // 14 elements per thread
            float v[NE];
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                const int e = threadIdx.x + i * FE3_THREADS;
                const int wl = e / FE3_HP, y = e - wl * FE3_HP;           // (B,W,H) layout: height is contiguous
                const int w = w0 + wl, yy = y - 1;
                v[i] = 0.0f;
                if (e < FE3_NW * FE3_HP && w >= 0 && w < W && yy >= 0 && yy < FE3_H) v[i] = (float)img[b * sb + w * sw + yy * sh];
            }
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                const int e = threadIdx.x + i * FE3_THREADS;
                if (e < FE3_NW * FE3_HP) {
                    if constexpr (sizeof(PixT) == 1) patch[e] = __fdiv_rn(v[i], 255.0f);
                    else patch[e] = v[i];
                }
            }
        }
        __syncthreads();

        // ---- phase 2: im2col rows of conv1 (convolution.py:192-196; conv "height" = time, conv "width" = image height) ----
        for (int pos = threadIdx.x; pos < FE3_NPOS; pos += FE3_THREADS) {
            const int r = pos / FE3_F2, f2 = pos - r * FE3_F2;
            uint8_t *row = sA1 + (uint32_t)(pos >> 7) * (128 * 64);
            if (t20 + r < 0 || t20 + r >= T2) {
                // outside the conv1 output = zero padding of the depthwise conv: an all-zero row (bias columns included) gives 0
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4 *>(row + sw64_offset(pos & 127, ch)) = make_uint4(0u, 0u, 0u, 0u);
                continue;
            }
            float hi[9], lo[9];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float v = patch[(2 * r + kh) * FE3_HP + 2 * f2 + kw];
                    const float h = fe3_bf_round(v);
                    hi[kh * 3 + kw] = h;
                    lo[kh * 3 + kw] = v - h;
                }
            uint32_t pk[16];                                   // 32 bf16: hi[0..8] lo[0..8] hi[0..8] 0 x 5
            pk[0] = fe3_pack(hi[0], hi[1]); pk[1] = fe3_pack(hi[2], hi[3]); pk[2] = fe3_pack(hi[4], hi[5]); pk[3] = fe3_pack(hi[6], hi[7]);
            pk[4] = fe3_pack(hi[8], lo[0]); pk[5] = fe3_pack(lo[1], lo[2]); pk[6] = fe3_pack(lo[3], lo[4]); pk[7] = fe3_pack(lo[5], lo[6]);
            pk[8] = fe3_pack(lo[7], lo[8]); pk[9] = pk[0]; pk[10] = pk[1]; pk[11] = pk[2];
            pk[12] = pk[3]; pk[13] = fe3_pack(hi[8], 1.0f); pk[14] = fe3_pack(1.0f, 0.0f); pk[15] = 0u;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
                *reinterpret_cast<uint4 *>(row + sw64_offset(pos & 127, ch)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
        fence_proxy_async_smem();
        __syncthreads();

        // ---- M1: conv1 as 7 M tiles x 2 k-steps ----
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(sA1), b0 = smem_u32(sB1);
#pragma unroll
                for (int mt = 0; mt < FE3_MT1; ++mt)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        umma_bf16_ss(tmem_base + (uint32_t)(mt * 32), fe3_desc(a0 + mt * (128 * 64) + ks * 32, 512, kLayoutSw64),
                                     fe3_desc(b0 + ks * 32, 512, kLayoutSw64), idesc, ks ? 1u : 0u);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, n_done & 1u);
        ++n_done;
        tc_fence_after();

        // ---- phase 3: conv1 epilogue -> a1 (aliases the im2col rows: every MMA reading them has completed) ----
#pragma unroll 1
        for (int it = 0; it < (FE3_MT1 + 1) / 2; ++it) {
            const int mt = 2 * it + half;
            if (mt < FE3_MT1) {                                // warp-uniform
                uint32_t r32[32];
                tmem_ld_32x32b_x32(tmem_base + lane_base + (uint32_t)(mt * 32), r32);
                tmem_ld_wait_dep(r32);
                const int pos = mt * 128 + q * 32 + lane;
                if (pos < FE3_NPOS) {
                    uint32_t pk[16];
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r32[c]), __uint_as_float(r32[c + 1]));
                        h2 = __hmax2(h2, __floats2bfloat162_rn(0.0f, 0.0f));           // ReLU on the packed pair
                        pk[c >> 1] = *reinterpret_cast<uint32_t *>(&h2);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        *reinterpret_cast<uint4 *>(sA1 + fe3_a1_offset(pos, g)) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();

        // ---- phase 4: depthwise conv (convolution.py:200-205) -> [hi | lo] rows of the 1x1 conv's A operand (aliases the patch).
        // Warps 2g, 2g+1 own channel group g; a thread takes positions j, j + 64, j + 128 of the 192 ----
        {
            constexpr int NP = FE3_ROWS2 / 64;
            const int g = warp >> 1, j = threadIdx.x & 63;
            int tl[NP], f[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) { const int p2 = j + 64 * p; tl[p] = p2 / FE3_F; f[p] = p2 - tl[p] * FE3_F; }
            float o[NP][8];
            switch (g) {
                case 0: fe3_dw_group<0, NP>(fw, sA1, tl, f, o); break;
                case 1: fe3_dw_group<1, NP>(fw, sA1, tl, f, o); break;
                case 2: fe3_dw_group<2, NP>(fw, sA1, tl, f, o); break;
                default: fe3_dw_group<3, NP>(fw, sA1, tl, f, o); break;
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int p2 = j + 64 * p;
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) {
                    hi[c2] = fe3_pack(o[p][2 * c2], o[p][2 * c2 + 1]);
                    const float2 hf = fe3_bf2_to_f2(hi[c2]);
                    lo[c2] = fe3_pack(o[p][2 * c2] - hf.x, o[p][2 * c2 + 1] - hf.y);
                }
                uint8_t *row = sY + (uint32_t)(p2 >> 7) * (128 * 128);
                *reinterpret_cast<uint4 *>(row + sw128_offset(p2 & 127, g)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(row + sw128_offset(p2 & 127, 4 + g)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        fence_proxy_async_smem();
        __syncthreads();

        // ---- M2: 1x1 conv, 2 M tiles x 4 k-steps (K = 64 = [hi | lo] against [Wpw | Wpw]) ----
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(sY), b0 = smem_u32(sB2);
#pragma unroll
                for (int mt = 0; mt < FE3_MT2; ++mt)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16_ss(tmem_base + (uint32_t)(mt * 32), fe3_desc(a0 + mt * (128 * 128) + ks * 32, 1024, kLayoutSw128),
                                     fe3_desc(b0 + ks * 32, 1024, kLayoutSw128), idesc, ks ? 1u : 0u);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, n_done & 1u);
        ++n_done;
        tc_fence_after();

        // ---- phase 5: + bias, ReLU (convolution.py:207-213) -> feature rows (b, t, f*32 + c) ----
        {
            uint32_t r32[32];
            tmem_ld_32x32b_x32(tmem_base + lane_base + (uint32_t)(half * 32), r32);
            tmem_ld_wait_dep(r32);
            const int p2 = half * 128 + q * 32 + lane;
            const int tl = p2 / FE3_F, f = p2 - tl * FE3_F;
            if (p2 < FE3_ROWS2 && t0 + tl < T) {
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 32; c += 2)
                    pk[c >> 1] = fe3_pack(fmaxf(__uint_as_float(r32[c]) + sbias[32 + c], 0.0f), fmaxf(__uint_as_float(r32[c + 1]) + sbias[32 + c + 1], 0.0f));
                uint4 *dst = reinterpret_cast<uint4 *>(feat + (((int64_t)b * T + t0 + tl) * FE3_F + f) * 32);
#pragma unroll
                for (int g = 0; g < 4; ++g) dst[g] = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
            }
        }
        tc_fence_before();
        __syncthreads();                                       // TMEM and both shared regions are free for the next item
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FE3_TMEM_COLS);
    }
}

// COCR_ERR_UNSUPPORTED when the geometry is not the default one (the caller falls back to the CUDA-core front-end)
int launch_frontend_tc(cocr_handle *h, const void *img_any, bool u8, int64_t sb, int64_t sw, int64_t sh, int B, int W, const bf16 *wpw2,
                       bf16 *feat, cudaStream_t st) {
    const Dims &dm = h->dm;
    if (dm.cc != 32 || dm.H != FE3_H || dm.F2 != FE3_F2 || dm.F != FE3_F || h->fe_weights.size() != sizeof(FrontendWeights3) / 4 || !wpw2)
        return COCR_ERR_UNSUPPORTED;
    const int T2 = out_len1(W), T = out_len1(T2);
    static PerDevice<bool> attr;
    if (!attr.get()) {
        COCR_CUDA(h, cudaFuncSetAttribute(k_frontend_tc<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FE3_SMEM));
        COCR_CUDA(h, cudaFuncSetAttribute(k_frontend_tc<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FE3_SMEM));
        attr.get() = true;
    }
    FrontendWeights3 fw;
    memcpy(&fw, h->fe_weights.data(), sizeof(fw));
    const int n_tt = (T + FE3_TT - 1) / FE3_TT, n_items = n_tt * B;
    const int grid = n_items < 2 * h->num_sms ? n_items : 2 * h->num_sms;
    const double n1 = (double)B * T2 * dm.F2 * 32, n2 = (double)B * T * dm.F * 32;
    ProfScope ps(h, "frontend_tc", st, 18.0 * n1 + 18.0 * n2 + 64.0 * n2, (u8 ? 1.0 : 4.0) * B * (double)W * dm.H + 2.0 * n2);
    if (u8)
        k_frontend_tc<uint8_t><<<grid, FE3_THREADS, FE3_SMEM, st>>>(static_cast<const uint8_t *>(img_any), sb, sw, sh, fw, wpw2, h->sub_bpw, feat, W, T2, T,
                                                                    n_tt, n_items);
    else
        k_frontend_tc<float><<<grid, FE3_THREADS, FE3_SMEM, st>>>(static_cast<const float *>(img_any), sb, sw, sh, fw, wpw2, h->sub_bpw, feat, W, T2, T,
                                                                  n_tt, n_items);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

}  // namespace cocr
