// Persistent warp-specialised bf16 GEMM on tcgen05 tensor cores:  C[M,N] = epi(A[M,K] * W[N,K]^T + bias)
//
//   * A (activations, row-major bf16) streams through a ring of 128x64 SWIZZLE_128B tiles loaded by TMA;
//   * W (weights, row-major [N,K] bf16, i.e. K-major "B" operand) is either RESIDENT in shared memory for
//     the whole kernel (one N-slice per CTA; all K=144 layers) or streamed through the same ring (K=576/768);
//   * one elected thread issues tcgen05.mma (UMMA 128 x block_n x 16) into a double-buffered TMEM accumulator;
//   * four epilogue warps read TMEM with tcgen05.ld (thread = row), apply bias / SiLU / arg-max, transpose
//     through a per-warp swizzled staging tile and write 128-byte row segments with coalesced 16-byte stores.
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#pragma once

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
    void *out;               // bf16 / fp32 (M, ldc) row-major, or unused for EPI_ARGMAX
    int ldc;
    unsigned long long *best;   // EPI_ARGMAX: (M) packed keys, zero-initialised by the caller
    // EPI_BF16_POS
    const int16_t *col_meta;    // (N/16): kind << 12 | first pair index
    const float2 *trig;         // (T, trig_ld): (sin, cos)(t * omega_m)
    int trig_ld, T;             // row r of the GEMM is frame t = r % T of its line
};

constexpr int GEMM_THREADS = 192;
constexpr int GEMM_BM = 128;
constexpr int GEMM_KC = 64;                       // k-chunk (elements) = one 128-byte swizzle span of bf16
constexpr uint32_t GEMM_A_CHUNK = GEMM_BM * 128;  // bytes of one A k-chunk tile

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

__device__ __forceinline__ float silu_fast(float v) {
    // x*sigmoid(x) = 0.5*x*(1 + tanh(0.5*x)); one MUFU op
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
    return 0.5f * v * (1.0f + t);
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
    uint8_t *staging = ring + (uint32_t)p.stages * slot_bytes;           // 4 warps x 2 x 4 KB
    float *sbias = reinterpret_cast<float *>(staging + 4 * 2 * 4096);     // 2 x 256 floats
    uint64_t *bars = reinterpret_cast<uint64_t *>(sbias + 512);
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
        mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
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
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int et = (warp - 2) * 32 + lane;           // 0..127
        uint8_t *stage_base = staging + (uint32_t)(warp - 2) * 2 * 4096;
        constexpr int CH = (EPI == EPI_F32) ? 32 : 64;   // columns per 128-byte row segment
        for (int j = 0; j < my_count; ++j) {
            int mt, nt;
            tile_mn(j, mt, nt);
            const int acc = j & 1;
            const uint32_t acc_phase = (uint32_t)(j >> 1) & 1u;
            const int n0 = nt * p.block_n;
            float *sb = sbias + acc * 256;
            for (int c = et; c < p.block_n; c += 128) sb[c] = (p.bias && n0 + c < p.N) ? p.bias[n0 + c] : 0.0f;
            named_bar_sync(1, 128);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t)(acc * p.block_n) + ((uint32_t)(q * 32) << 16);
            const int64_t grow0 = (int64_t)mt * GEMM_BM + q * 32;    // first global row of this warp
            float best_v = -INFINITY;
            int best_c = 0x7fffffff;
            int buf = 0;
            for (int c0 = 0; c0 < p.block_n; c0 += CH) {
                const int ncol = min(CH, p.block_n - c0);            // multiple of 16
                uint32_t r[CH];
#pragma unroll
                for (int pc = 0; pc < CH / 16; ++pc) {
                    if (pc * 16 < ncol) {
                        uint32_t t16[16];
                        tmem_ld_32x32b_x16(t_row + (uint32_t)(c0 + pc * 16), t16);
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[pc * 16 + i] = t16[i];
                    }
                }
                tmem_ld_wait();
                if constexpr (EPI == EPI_ARGMAX) {
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        if (i < ncol && n0 + c0 + i < p.N) {
                            float v = __uint_as_float(r[i]) + sb[c0 + i];
                            if (v > best_v) { best_v = v; best_c = n0 + c0 + i; }
                        }
                    }
                } else {
                    uint8_t *sbuf = stage_base + buf * 4096;
                    if constexpr (EPI == EPI_F32) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            if (jj * 4 < ncol) {
                                float4 o;
                                o.x = __uint_as_float(r[jj * 4 + 0]) + sb[c0 + jj * 4 + 0];
                                o.y = __uint_as_float(r[jj * 4 + 1]) + sb[c0 + jj * 4 + 1];
                                o.z = __uint_as_float(r[jj * 4 + 2]) + sb[c0 + jj * 4 + 2];
                                o.w = __uint_as_float(r[jj * 4 + 3]) + sb[c0 + jj * 4 + 3];
                                *reinterpret_cast<float4 *>(sbuf + sw128_offset(lane, jj)) = o;
                            }
                        }
                    } else {
                        int trow = 0;
                        if constexpr (EPI == EPI_BF16_POS) trow = (int)((grow0 + lane) % p.T);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (g * 16 < ncol) {
                                float f[16];
#pragma unroll
                                for (int e = 0; e < 16; ++e) {
                                    f[e] = __uint_as_float(r[g * 16 + e]) + sb[c0 + g * 16 + e];
                                    if constexpr (EPI == EPI_BF16_SILU) f[e] = silu_fast(f[e]);
                                }
                                if constexpr (EPI == EPI_BF16_POS) {
                                    const int meta = p.col_meta[(n0 + c0) / 16 + g];
                                    const int kind = meta >> 12;
                                    if (kind != 0) {
                                        const float4 *tp = reinterpret_cast<const float4 *>(p.trig + (size_t)trow * p.trig_ld + (meta & 0xFFF));
#pragma unroll
                                        for (int q4 = 0; q4 < 4; ++q4) {
                                            const float4 sc = tp[q4];     // (sin, cos) of pairs 2*q4, 2*q4+1
                                            const float sn[2] = {sc.x, sc.z}, cs[2] = {sc.y, sc.w};
#pragma unroll
                                            for (int u = 0; u < 2; ++u) {
                                                const int e0 = (q4 * 2 + u) * 2;
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
                                    }
                                }
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    uint32_t pk[4];
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[hh * 8 + 2 * e], f[hh * 8 + 2 * e + 1]);
                                        pk[e] = *reinterpret_cast<uint32_t *>(&h2);
                                    }
                                    *reinterpret_cast<uint4 *>(sbuf + sw128_offset(lane, g * 2 + hh)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    // coalesced write-out: 8 lanes cover one 128-byte row segment, 4 rows per instruction
                    constexpr int EPC = (EPI == EPI_F32) ? 4 : 8;     // elements per 16-byte piece
                    const int cj = lane & 7;
                    const int gcol = n0 + c0 + cj * EPC;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + (lane >> 3);
                        const int64_t grow = grow0 + row;
                        if (cj * EPC < ncol && grow < p.M && gcol < p.N) {
                            uint4 v = *reinterpret_cast<const uint4 *>(sbuf + sw128_offset(row, cj));
                            if constexpr (EPI == EPI_F32) {
                                float *dst = reinterpret_cast<float *>(p.out) + grow * p.ldc + gcol;
                                if (gcol + 3 < p.N && (p.ldc & 3) == 0) {
                                    *reinterpret_cast<uint4 *>(dst) = v;
                                } else {
                                    const float *vf = reinterpret_cast<const float *>(&v);
                                    for (int e = 0; e < 4; ++e)
                                        if (gcol + e < p.N) dst[e] = vf[e];
                                }
                            } else {
                                __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(p.out) + grow * p.ldc + gcol;
                                if (gcol + 7 < p.N && (p.ldc & 7) == 0) {
                                    *reinterpret_cast<uint4 *>(dst) = v;
                                } else {
                                    const __nv_bfloat16 *vh = reinterpret_cast<const __nv_bfloat16 *>(&v);
                                    for (int e = 0; e < 8; ++e)
                                        if (gcol + e < p.N) dst[e] = vh[e];
                                }
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
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

}  // namespace cocr
