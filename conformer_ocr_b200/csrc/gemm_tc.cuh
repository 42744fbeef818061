// Persistent warp-specialised bf16 GEMM on tcgen05 tensor cores:  C[M,N] = epi(A[M,K] * W[N,K]^T + bias)
//
//   * A (activations, row-major bf16) streams through a ring of 128x64 SWIZZLE_128B tiles loaded by TMA;
//   * W (weights, row-major [N,K] bf16, i.e. K-major "B" operand) is either RESIDENT in shared memory for
//     the whole kernel (one N-slice per CTA; all K=144 layers) or streamed through the same ring (K=576/768);
//   * one elected thread issues tcgen05.mma (UMMA 128 x block_n x 16) into a double-buffered TMEM accumulator;
//   * NE (8) epilogue warps - two per TMEM lane quarter, splitting the columns in 32-wide pieces - read TMEM with
//     tcgen05.ld (thread = row), apply the fused epilogue, transpose through a per-warp swizzled staging tile
//     and move 64/128-byte row segments with coalesced 16-byte accesses.
//
// Epilogues: bias (+SiLU) -> bf16; bias -> fp32; bias -> running arg-max (CTC decode, logits never stored);
// bias + rotation by the frame angle (attention operands, see attn_tc.cuh); and the residual epilogue
//   x <- x + alpha*(acc + bias) [-> LayerNorm_final(x)] ; xn = bf16(LayerNorm_next(x))
// which keeps the fp32 residual stream and both LayerNorms of a conformer block inside the GEMM
// (modules.py:31-32, encoder.py:99, feed_forward.py:46, attention.py:148, convolution.py:136).
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..2+NE-1 = epilogue.
#pragma once

#include <cuda_fp16.h>

#include "tc_utils.cuh"

namespace cocr {

enum GemmEpi : int {
    EPI_BF16 = 0,        // out bf16 = acc + bias
    EPI_BF16_SILU = 1,   // out bf16 = silu(acc + bias)
    EPI_F32 = 2,         // out fp32 = acc + bias
    EPI_ARGMAX = 3,      // best[row] = max over columns of (acc + bias), lowest column wins ties (packed u64)
    EPI_BF16_POS = 4,    // out bf16 = acc + bias, with per-16-column-group position handling (attention operands):
                         //   kind 0 plain; kind 1 rotate pairs by the row's angle (relative-position term);
                         //   kind 2 fill with (cos, sin) of the row's angle
    EPI_RESID = 5,       // fp32 residual stream update + LayerNorm(s); requires N == block_n (one N tile)
};

struct GemmArgs {
    int M, N, K;             // K % 16 == 0
    int block_n;             // UMMA N: multiple of 16, 16..256
    int n_tiles, m_tiles;
    int nkc;                 // ceil(K / 64) k-chunks
    int last_ksteps;         // k-steps (of 16) in the last chunk
    int resident;            // 1: W slice resident in smem (grid is a multiple of n_tiles)
    int stages;              // ring slots
    int tmem_cols;           // power of two >= 2*block_n
    const float *bias;       // (N) or null
    void *out;               // bf16 / fp32 (M, ldc) row-major; EPI_RESID: bf16 xn (M, ldc)
    int ldc;
    unsigned long long *best;   // EPI_ARGMAX: (M) packed keys, zero-initialised by the caller
    // EPI_BF16_POS
    const int16_t *col_meta;    // (N/16): kind << 12 | first pair index
    const float2 *trig;         // (T, trig_ld): (sin, cos)(t * omega_m)
    int trig_ld, T;             // row r of the GEMM is frame t = r % T of its line
    // EPI_RESID
    float *x;                   // (M, N) fp32 residual stream, updated in place
    float alpha;                // module factor (0.5 for the half-step FFNs)
    int init;                   // 1: x = acc + bias (start of the stream)
    const float *gf, *bf;       // block-final LayerNorm applied to x itself (or null)
    const float *g1, *b1;       // LayerNorm of the next module -> xn (null: xn = bf16(x))
};

constexpr int GEMM_NE = 8;                          // epilogue warps
constexpr int GEMM_THREADS = 64 + 32 * GEMM_NE;
constexpr int GEMM_BM = 128;
constexpr int GEMM_KC = 64;                       // k-chunk (elements) = one 128-byte swizzle span of bf16
constexpr uint32_t GEMM_A_CHUNK = GEMM_BM * 128;  // bytes of one A k-chunk tile
constexpr uint32_t GEMM_STAGE_BYTES = 4096;       // per-warp staging buffer (32 rows x 128 B), two per warp
constexpr uint32_t GEMM_FIXED_SMEM = GEMM_NE * 2 * GEMM_STAGE_BYTES + 7 * 256 * 4 + 2 * 128 * 2 * 4 * 2 + 512 + 1024;

__host__ __device__ inline uint32_t gemm_b_chunk_bytes(int block_n) { return (uint32_t)block_n * 128u; }

// orderable key: larger float -> larger key; among equal floats the LOWER column index wins
__device__ __forceinline__ unsigned long long pack_best(float v, int col) {
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)col);
}
__device__ __forceinline__ void unpack_best(unsigned long long k, float *v, int *col) {
    uint32_t b = (uint32_t)(k >> 32);
    b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
    *v = __uint_as_float(b);
    *col = (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
}

// silu(a), silu(b) -> packed bf16x2.  x*sigmoid(x) = h + h*tanh(h), h = x/2, evaluated as f16x2 (one MUFU op for
// two elements; 11-bit mantissa, then rounded once to bf16).  |x| is clamped to the f16 range.
__device__ __forceinline__ uint32_t silu2_bf16(float a, float b) {
    a = fminf(fmaxf(a, -60000.0f), 60000.0f);
    b = fminf(fmaxf(b, -60000.0f), 60000.0f);
    __half2 h = __floats2half2_rn(0.5f * a, 0.5f * b);
    uint32_t hu = *reinterpret_cast<uint32_t *>(&h), tu;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(hu));
    __half2 t = *reinterpret_cast<__half2 *>(&tu);
    __half2 r = __hfma2(h, t, h);
    float2 f = __half22float2(r);
    __nv_bfloat162 o = __floats2bfloat162_rn(f.x, f.y);
    return *reinterpret_cast<uint32_t *>(&o);
}

// byte offset of (row, 16-byte chunk 0..3) in a [32 rows][64 B] staging tile, conflict-free for row-wise 16-byte
// writes (thread = row) and for 4-lanes-per-row reads
__device__ __forceinline__ uint32_t stg64_offset(uint32_t row, uint32_t chunk) {
    return row * 64u + (((chunk ^ (row >> 1)) & 3u) << 4);
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B tiles
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const uint32_t b_chunk = gemm_b_chunk_bytes(p.block_n);
    const uint32_t bres_bytes = p.resident ? (uint32_t)p.nkc * b_chunk : 0u;
    const uint32_t slot_bytes = GEMM_A_CHUNK + (p.resident ? 0u : b_chunk);
    uint8_t *bres = smem;
    uint8_t *ring = smem + bres_bytes;
    uint8_t *staging = ring + (uint32_t)p.stages * slot_bytes;                        // NE warps x 2 x 4 KB
    float *sbias = reinterpret_cast<float *>(staging + GEMM_NE * 2 * GEMM_STAGE_BYTES); // [2][256] bias (per acc stage)
    float *sln = sbias + 512;                                                           // [4][256] gf, bf, g1, b1
    float *spad = sln + 1024;                                                           // [256] spare
    float2 *sred = reinterpret_cast<float2 *>(spad + 256);                              // [2 rounds][2 halves][128] partial sums
    uint64_t *bars = reinterpret_cast<uint64_t *>(sred + 2 * 2 * 128);
    uint64_t *full = bars;                       // [stages]
    uint64_t *empty = bars + p.stages;           // [stages]
    uint64_t *bfull = bars + 2 * p.stages;       // [1]
    uint64_t *tfull = bfull + 1;                 // [2]
    uint64_t *tempty = tfull + 2;                // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile schedule shared by all roles
    int n_fixed = 0, my_first, my_stride, my_count;
    if (p.resident) {
        const int groups = gridDim.x / p.n_tiles;
        n_fixed = blockIdx.x % p.n_tiles;
        my_first = blockIdx.x / p.n_tiles;
        my_stride = groups;
        my_count = (my_first < p.m_tiles) ? (p.m_tiles - my_first + groups - 1) / groups : 0;
    } else {
        const int total = p.m_tiles * p.n_tiles;
        my_first = blockIdx.x;
        my_stride = gridDim.x;
        my_count = (my_first < total) ? (total - my_first + my_stride - 1) / my_stride : 0;
    }
    auto tile_mn = [&](int j, int &mt, int &nt) {
        int t = my_first + j * my_stride;
        if (p.resident) { mt = t; nt = n_fixed; }
        else { mt = t / p.n_tiles; nt = t % p.n_tiles; }
    };

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(bfull, 1);
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], GEMM_NE); mbar_init(&tempty[1], GEMM_NE);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if constexpr (EPI == EPI_RESID) {
        // LayerNorm parameters are tile independent (single N tile)
        for (int c = threadIdx.x; c < 256; c += GEMM_THREADS) {
            const bool in = c < p.N;
            sln[c] = (in && p.gf) ? p.gf[c] : 1.0f;
            sln[256 + c] = (in && p.bf) ? p.bf[c] : 0.0f;
            sln[512 + c] = (in && p.g1) ? p.g1[c] : 1.0f;
            sln[768 + c] = (in && p.b1) ? p.b1[c] : 0.0f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            if (p.resident && my_count > 0) {
                mbar_expect_tx(bfull, bres_bytes);
                for (int kc = 0; kc < p.nkc; ++kc) tma_load_2d(bres + (uint32_t)kc * b_chunk, &tmB, bfull, kc * GEMM_KC, n_fixed * p.block_n);
            }
            int slot = 0;
            uint32_t phase = 0;
            for (int j = 0; j < my_count; ++j) {
                int mt, nt;
                tile_mn(j, mt, nt);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&empty[slot], phase ^ 1u);
                    mbar_expect_tx(&full[slot], slot_bytes);
                    uint8_t *dst = ring + (uint32_t)slot * slot_bytes;
                    tma_load_2d(dst, &tmA, &full[slot], kc * GEMM_KC, mt * GEMM_BM);
                    if (!p.resident) tma_load_2d(dst + GEMM_A_CHUNK, &tmB, &full[slot], kc * GEMM_KC, nt * p.block_n);
                    if (++slot == p.stages) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(GEMM_BM, p.block_n);
            if (p.resident && my_count > 0) mbar_wait(bfull, 0);
            int slot = 0;
            uint32_t phase = 0;
            for (int j = 0; j < my_count; ++j) {
                const int acc = j & 1;
                const uint32_t acc_phase = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&full[slot], phase);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(ring + (uint32_t)slot * slot_bytes);
                    const uint32_t b_base = p.resident ? smem_u32(bres + (uint32_t)kc * b_chunk) : a_base + GEMM_A_CHUNK;
                    const int nks = (kc == p.nkc - 1) ? p.last_ksteps : 4;
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t ad = make_smem_desc_sw128(a_base + ks * 32, 1024);
                        const uint64_t bd = make_smem_desc_sw128(b_base + ks * 32, 1024);
                        umma_bf16_ss(d_tmem, ad, bd, idesc, (kc | ks) ? 1u : 0u);
                    }
                    umma_commit(&empty[slot]);
                    if (++slot == p.stages) { slot = 0; phase ^= 1u; }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ================= epilogue =================
        const int ew = warp - 2;                         // 0 .. NE-1
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = ew >> 2;                        // which of the NE/4 column shares
        constexpr int NSH = GEMM_NE / 4;                 // warps sharing one lane quarter
        const int et = ew * 32 + lane;                   // 0 .. 32*NE-1
        uint8_t *stage_base = staging + (uint32_t)ew * 2 * GEMM_STAGE_BYTES;
        const int npieces = (p.block_n + 31) / 32;       // 32-column pieces
        for (int j = 0; j < my_count; ++j) {
            int mt, nt;
            tile_mn(j, mt, nt);
            const int acc = j & 1;
            const uint32_t acc_phase = (uint32_t)(j >> 1) & 1u;
            const int n0 = nt * p.block_n;
            float *sb = sbias + acc * 256;
            if (et < p.block_n) sb[et] = (p.bias && n0 + et < p.N) ? p.bias[n0 + et] : 0.0f;
            named_bar_sync(1, 32 * GEMM_NE);
            const uint32_t t_row = tmem_base + (uint32_t)(acc * p.block_n) + ((uint32_t)(q * 32) << 16);
            const int64_t grow0 = (int64_t)mt * GEMM_BM + q * 32;    // first global row of this warp

            if constexpr (EPI == EPI_RESID) {
                // ---- fp32 residual update + LayerNorm(s); this warp owns rows [grow0, +32) x its column pieces ----
                constexpr int MAXP = 3;                    // pieces per warp: block_n <= 192 -> 6 pieces / 2 warps
                float v[MAXP][32];
                // (1) pull the x pieces in through the staging tile (coalesced 128-byte row segments)
                if (!p.init) {
#pragma unroll
                    for (int k = 0; k < MAXP; ++k) {
                        const int pi = half + NSH * k;
                        if (pi < npieces) {
                            const int c0 = pi * 32, ncol = min(32, p.block_n - c0);
                            uint8_t *sbuf = stage_base + (k & 1) * GEMM_STAGE_BYTES;
                            const int cj = lane & 7;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int row = i * 4 + (lane >> 3);
                                const int64_t grow = grow0 + row;
                                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (cj * 4 < ncol && grow < p.M) t = *reinterpret_cast<const float4 *>(p.x + grow * p.N + c0 + cj * 4);
                                *reinterpret_cast<float4 *>(sbuf + sw128_offset(row, cj)) = t;
                            }
                            __syncwarp();
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                const float4 t = *reinterpret_cast<const float4 *>(sbuf + sw128_offset(lane, jj));
                                v[k][jj * 4 + 0] = t.x; v[k][jj * 4 + 1] = t.y; v[k][jj * 4 + 2] = t.z; v[k][jj * 4 + 3] = t.w;
                            }
                            __syncwarp();
                        }
                    }
                }
                // (2) accumulator -> x_new = x + alpha * (acc + bias); partial LayerNorm sums
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int k = 0; k < MAXP; ++k) {
                    const int pi = half + NSH * k;
                    if (pi < npieces) {
                        const int c0 = pi * 32, ncol = min(32, p.block_n - c0);
                        uint32_t r[32];
                        if (ncol == 32) {
                            tmem_ld_32x32b_x32(t_row + (uint32_t)c0, r);
                        } else {
                            uint32_t t16[16];
                            tmem_ld_32x32b_x16(t_row + (uint32_t)c0, t16);
#pragma unroll
                            for (int i = 0; i < 16; ++i) { r[i] = t16[i]; r[16 + i] = 0u; }
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const float a = __uint_as_float(r[e]) + sb[c0 + e];
                            const float t = p.init ? a : fmaf(p.alpha, a, v[k][e]);
                            const float tv = (e < ncol) ? t : 0.0f;
                            v[k][e] = tv;
                            s1 += tv;
                            s2 = fmaf(tv, tv, s2);
                        }
                    }
                }
                // the accumulator stage can be reused by the MMA warp from here on
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);

                const float inv_n = 1.0f / (float)p.N;
                const int rix = q * 32 + lane;                               // row inside the tile
                float mean, rstd;
                auto exchange = [&](int round) {
                    sred[(round * 2 + half) * 128 + rix] = make_float2(s1, s2);
                    named_bar_sync(2 + q, 32 * NSH);
                    const float2 o = sred[(round * 2 + (half ^ 1)) * 128 + rix];
                    const float t1 = s1 + o.x, t2 = s2 + o.y;
                    mean = t1 * inv_n;
                    rstd = rsqrtf(fmaxf(t2 * inv_n - mean * mean, 0.0f) + 1e-5f);
                };
                if (p.gf) {   // block-final LayerNorm replaces the residual stream (encoder.py:99)
                    exchange(0);
                    s1 = 0.f; s2 = 0.f;
#pragma unroll
                    for (int k = 0; k < MAXP; ++k) {
                        const int pi = half + NSH * k;
                        if (pi < npieces) {
                            const int c0 = pi * 32, ncol = min(32, p.block_n - c0);
#pragma unroll
                            for (int e = 0; e < 32; ++e) {
                                const float t = (e < ncol) ? (v[k][e] - mean) * rstd * sln[c0 + e] + sln[256 + c0 + e] : 0.0f;
                                v[k][e] = t;
                                s1 += t;
                                s2 = fmaf(t, t, s2);
                            }
                        }
                    }
                }
                // (3) write x_new (fp32, coalesced through staging)
#pragma unroll
                for (int k = 0; k < MAXP; ++k) {
                    const int pi = half + NSH * k;
                    if (pi < npieces) {
                        const int c0 = pi * 32, ncol = min(32, p.block_n - c0);
                        uint8_t *sbuf = stage_base + (k & 1) * GEMM_STAGE_BYTES;
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj)
                            *reinterpret_cast<float4 *>(sbuf + sw128_offset(lane, jj)) =
                                make_float4(v[k][jj * 4], v[k][jj * 4 + 1], v[k][jj * 4 + 2], v[k][jj * 4 + 3]);
                        __syncwarp();
                        const int cj = lane & 7;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = i * 4 + (lane >> 3);
                            const int64_t grow = grow0 + row;
                            if (cj * 4 < ncol && grow < p.M)
                                *reinterpret_cast<float4 *>(p.x + grow * p.N + c0 + cj * 4) =
                                    *reinterpret_cast<const float4 *>(sbuf + sw128_offset(row, cj));
                        }
                        __syncwarp();
                    }
                }
                // (4) xn = bf16(LayerNorm_next(x_new)) (or bf16(x_new))
                if (p.g1) exchange(1);
#pragma unroll
                for (int k = 0; k < MAXP; ++k) {
                    const int pi = half + NSH * k;
                    if (pi < npieces) {
                        const int c0 = pi * 32, ncol = min(32, p.block_n - c0);
                        uint8_t *sbuf = stage_base + (k & 1) * GEMM_STAGE_BYTES;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            uint32_t pk[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int c = jj * 8 + 2 * e;
                                float a = v[k][c], b = v[k][c + 1];
                                if (p.g1) {
                                    a = (a - mean) * rstd * sln[512 + c0 + c] + sln[768 + c0 + c];
                                    b = (b - mean) * rstd * sln[512 + c0 + c + 1] + sln[768 + c0 + c + 1];
                                }
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                                pk[e] = *reinterpret_cast<uint32_t *>(&h2);
                            }
                            *reinterpret_cast<uint4 *>(sbuf + stg64_offset(lane, jj)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                        __syncwarp();
                        const int cj = lane & 3;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int row = i * 8 + (lane >> 2);
                            const int64_t grow = grow0 + row;
                            if (cj * 8 < ncol && grow < p.M)
                                *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + grow * p.ldc + c0 + cj * 8) =
                                    *reinterpret_cast<const uint4 *>(sbuf + stg64_offset(row, cj));
                        }
                        __syncwarp();
                    }
                }
            } else {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                float best_v = -INFINITY;
                int best_c = 0x7fffffff;
                int buf = 0;
                int trow = 0;
                if constexpr (EPI == EPI_BF16_POS) trow = (int)((grow0 + lane) % p.T);
                for (int pi = half; pi < npieces; pi += NSH) {
                    const int c0 = pi * 32;
                    const int ncol = min(32, p.block_n - c0);            // 32 or 16
                    uint32_t r[32];
                    if (ncol == 32) {
                        tmem_ld_32x32b_x32(t_row + (uint32_t)c0, r);
                    } else {
                        uint32_t t16[16];
                        tmem_ld_32x32b_x16(t_row + (uint32_t)c0, t16);
#pragma unroll
                        for (int i = 0; i < 16; ++i) { r[i] = t16[i]; r[16 + i] = 0u; }
                    }
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(sb + c0 + e4 * 4);
                        f[e4 * 4 + 0] = __uint_as_float(r[e4 * 4 + 0]) + b4.x;
                        f[e4 * 4 + 1] = __uint_as_float(r[e4 * 4 + 1]) + b4.y;
                        f[e4 * 4 + 2] = __uint_as_float(r[e4 * 4 + 2]) + b4.z;
                        f[e4 * 4 + 3] = __uint_as_float(r[e4 * 4 + 3]) + b4.w;
                    }
                    if constexpr (EPI == EPI_ARGMAX) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < ncol && n0 + c0 + i < p.N && f[i] > best_v) { best_v = f[i]; best_c = n0 + c0 + i; }
                        }
                    } else if constexpr (EPI == EPI_F32) {
                        // two 16-column halves, each a [32 rows][64 B] staging tile
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            if (hh * 16 < ncol) {
                                uint8_t *sbuf = stage_base + buf * GEMM_STAGE_BYTES;
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj)
                                    *reinterpret_cast<float4 *>(sbuf + stg64_offset(lane, jj)) =
                                        make_float4(f[hh * 16 + jj * 4], f[hh * 16 + jj * 4 + 1], f[hh * 16 + jj * 4 + 2], f[hh * 16 + jj * 4 + 3]);
                                __syncwarp();
                                const int cj = lane & 3;
                                const int gcol = n0 + c0 + hh * 16 + cj * 4;
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int row = i * 8 + (lane >> 2);
                                    const int64_t grow = grow0 + row;
                                    if (grow < p.M && gcol < p.N) {
                                        const float4 t = *reinterpret_cast<const float4 *>(sbuf + stg64_offset(row, cj));
                                        float *dst = reinterpret_cast<float *>(p.out) + grow * p.ldc + gcol;
                                        if (gcol + 3 < p.N && (p.ldc & 3) == 0) {
                                            *reinterpret_cast<float4 *>(dst) = t;
                                        } else {
                                            const float tv[4] = {t.x, t.y, t.z, t.w};
                                            for (int e = 0; e < 4; ++e)
                                                if (gcol + e < p.N) dst[e] = tv[e];
                                        }
                                    }
                                }
                                buf ^= 1;
                            }
                        }
                    } else {
                        uint8_t *sbuf = stage_base + buf * GEMM_STAGE_BYTES;
                        if constexpr (EPI == EPI_BF16_POS) {
                            // rotation / fill of the positional columns; the (sin, cos) rows of this warp's 32 frames are
                            // pulled in coalesced and transposed through the other staging buffer
                            uint8_t *tbuf = stage_base + (buf ^ 1) * GEMM_STAGE_BYTES;
                            __syncwarp();     // readers of the previous piece are done with that buffer
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                if (g * 16 < ncol) {
                                    const int meta = p.col_meta[(n0 + c0) / 16 + g];
                                    const int kind = meta >> 12;
                                    if (kind != 0) {   // warp-uniform
                                        const int pbase = meta & 0xFFF;
                                        // 32 rows x 8 pairs x 8 B = 32 x 64 B ; lane -> (row = i*8 + lane/4, 16-byte chunk lane%4)
#pragma unroll
                                        for (int i = 0; i < 4; ++i) {
                                            const int row = i * 8 + (lane >> 2);
                                            const int tr = (int)((grow0 + row) % p.T);
                                            const float4 t = *reinterpret_cast<const float4 *>(p.trig + (size_t)tr * p.trig_ld + pbase + (lane & 3) * 2);
                                            *reinterpret_cast<float4 *>(tbuf + stg64_offset(row, lane & 3)) = t;
                                        }
                                        __syncwarp();
#pragma unroll
                                        for (int q4 = 0; q4 < 4; ++q4) {
                                            const float4 sc = *reinterpret_cast<const float4 *>(tbuf + stg64_offset(lane, q4));
                                            const float sn[2] = {sc.x, sc.z}, cs[2] = {sc.y, sc.w};
#pragma unroll
                                            for (int u = 0; u < 2; ++u) {
                                                const int e0 = g * 16 + (q4 * 2 + u) * 2;
                                                const float as = f[e0], ac = f[e0 + 1];
                                                if (kind == 1) {
                                                    f[e0] = as * sn[u] + ac * cs[u];
                                                    f[e0 + 1] = ac * sn[u] - as * cs[u];
                                                } else {
                                                    f[e0] = cs[u];
                                                    f[e0 + 1] = sn[u];
                                                }
                                            }
                                        }
                                        __syncwarp();
                                    }
                                }
                            }
                            (void)trow;
                        }
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            if (jj * 8 < ncol) {
                                uint32_t pk[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float a = f[jj * 8 + 2 * e], b = f[jj * 8 + 2 * e + 1];
                                    if constexpr (EPI == EPI_BF16_SILU) {
                                        pk[e] = silu2_bf16(a, b);
                                    } else {
                                        __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                                        pk[e] = *reinterpret_cast<uint32_t *>(&h2);
                                    }
                                }
                                *reinterpret_cast<uint4 *>(sbuf + stg64_offset(lane, jj)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            }
                        }
                        __syncwarp();
                        // coalesced write-out: 4 lanes cover one 64-byte row segment, 8 rows per instruction
                        const int cj = lane & 3;
                        const int gcol = n0 + c0 + cj * 8;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int row = i * 8 + (lane >> 2);
                            const int64_t grow = grow0 + row;
                            if (cj * 8 < ncol && grow < p.M && gcol < p.N) {
                                const uint4 t = *reinterpret_cast<const uint4 *>(sbuf + stg64_offset(row, cj));
                                __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(p.out) + grow * p.ldc + gcol;
                                if (gcol + 7 < p.N && (p.ldc & 7) == 0) {
                                    *reinterpret_cast<uint4 *>(dst) = t;
                                } else {
                                    const __nv_bfloat16 *vh = reinterpret_cast<const __nv_bfloat16 *>(&t);
                                    for (int e = 0; e < 8; ++e)
                                        if (gcol + e < p.N) dst[e] = vh[e];
                                }
                            }
                        }
                        buf ^= 1;     // the other buffer is free: its readers finished before the last __syncwarp
                    }
                }
                if constexpr (EPI == EPI_ARGMAX) {
                    const int64_t grow = grow0 + lane;
                    if (grow < p.M && best_c != 0x7fffffff) atomicMax(p.best + grow, pack_best(best_v, best_c));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

}  // namespace cocr
