// Fused relative-position multi-head attention on tcgen05 tensor cores (attention.py:72-103).
//
// The reference adds a Transformer-XL positional score (q_i+v).P[i-j] through a (T, 2T-1) GEMM followed by
// an index shift.  Because P[r] = W_p * PE(r) with sinusoidal PE, PE(i-j) = R(i) * E(j) for a block-diagonal
// rotation R(i); so   (q_i+v).P[i-j] = a~_i . E(j)   with a~_i = R(i)^T W_p,h^T (q_i+v)  and
// E(j) = (cos(j w_m), sin(j w_m))_m.   The whole score is ONE contraction of length 48 + d:
//        S[i,j] = [q_i+u | a~_i] . [k_j | E(j)]
// The projection GEMM (EPI_BF16_POS) emits, per head,
//        Qe = [q+u (dh) | 0 | a~ tail (16) | a~ main (d-16)]     (48+d columns, 64-column chunks)
//        Kx = [k (dh)   | 0 | E tail (16)]                       (64 columns)
//        Vx = [v (dh)   | 0]                                      (64 columns)
// and E main (d-16 columns) comes from a table shared by all lines / heads / layers.
//
// One CTA = 128 queries of one (line, head).  warps 0-3: online softmax (thread = query row, TMEM lane),
// warp 4: TMA producer, warp 5: TMEM allocator + MMA issuer.  S is double-buffered in TMEM so that
// Q K^T of tile j+1 overlaps the softmax of tile j; P goes through shared memory (bf16, SWIZZLE_128B).
#pragma once

#include "tc_utils.cuh"

namespace cocr {

struct AttnArgs {
    int B, T, nh, dh, d;
    int nchunk;                 // 64-column chunks of the score contraction = (48 + d) / 64  (3 for d = 144)
    const int32_t *lens;        // ragged mode: valid frames per line, else null
    __nv_bfloat16 *ctx;         // (B*T, d) head-major context
    float scale_log2;           // log2(e) / sqrt(dh)
};

constexpr int ATT_THREADS = 192;
constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr uint32_t ATT_CHUNK = 128 * 128;      // bytes of a 128-row x 64-col bf16 SWIZZLE_128B tile
constexpr int ATT_MAXCHUNK = 3;
constexpr int ATT_STAGES = 2;
constexpr int ATT_DV = 64;                     // padded head dim of the V operand / O accumulator
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_S_COL0 = 0, ATT_S_COL1 = 128, ATT_O_COL = 256;

__host__ __device__ inline uint32_t attn_stage_bytes(int nchunk) { return (uint32_t)(nchunk + 1) * ATT_CHUNK; }   // K chunks + V
__host__ __device__ inline uint32_t attn_smem_bytes(int nchunk) {
    return (uint32_t)nchunk * ATT_CHUNK + ATT_STAGES * attn_stage_bytes(nchunk) + 2 * ATT_CHUNK + 256 + 1024;
}

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
k_attn_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE, const AttnArgs p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t stage_bytes = attn_stage_bytes(p.nchunk);
    uint8_t *sQ = smem;                                                  // nchunk tiles
    uint8_t *sKV = sQ + (uint32_t)p.nchunk * ATT_CHUNK;                  // stages x [K chunk0 | E chunks | V]
    uint8_t *sP = sKV + ATT_STAGES * stage_bytes;                        // 2 tiles (keys 0-63, 64-127)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + 2 * ATT_CHUNK);
    uint64_t *qfull = bars;            // [1]
    uint64_t *kfull = bars + 1;        // [2]
    uint64_t *kempty = bars + 3;       // [2]
    uint64_t *sfull = bars + 5;        // [2]
    uint64_t *sempty = bars + 7;       // [2]
    uint64_t *pfull = bars + 9;        // [1]
    uint64_t *pvdone = bars + 10;      // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, hd = blockIdx.y, b = blockIdx.z;
    const int Tk = p.lens ? min(max(p.lens[b], 0), p.T) : p.T;           // keys beyond are masked
    const int ntiles = (Tk + ATT_BK - 1) / ATT_BK;
    const int row0 = b * p.T;                                             // first frame of this line in the (B*T) matrices
    const int qe_cols = 48 + p.d;                                         // columns of one head's Qe block

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmE);
        mbar_init(qfull, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kfull[s], 1); mbar_init(&kempty[s], 1);
            mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 4);
        }
        mbar_init(pfull, 4);
        mbar_init(pvdone, 1);
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ================= TMA producer =================
        if (lane == 0 && ntiles > 0) {
            mbar_expect_tx(qfull, (uint32_t)p.nchunk * ATT_CHUNK);
            for (int c = 0; c < p.nchunk; ++c)
                tma_load_2d(sQ + (uint32_t)c * ATT_CHUNK, &tmQ, qfull, hd * qe_cols + c * 64, row0 + qt * ATT_BQ);
            for (int j = 0; j < ntiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&kempty[st], ph ^ 1u);
                mbar_expect_tx(&kfull[st], stage_bytes);
                uint8_t *dst = sKV + (uint32_t)st * stage_bytes;
                tma_load_2d(dst, &tmK, &kfull[st], hd * 64, row0 + j * ATT_BK);
                for (int c = 1; c < p.nchunk; ++c)
                    tma_load_2d(dst + (uint32_t)c * ATT_CHUNK, &tmE, &kfull[st], (c - 1) * 64, j * ATT_BK);
                tma_load_2d(dst + (uint32_t)p.nchunk * ATT_CHUNK, &tmV, &kfull[st], hd * 64, row0 + j * ATT_BK);
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0 && ntiles > 0) {
            const uint32_t idesc_s = make_idesc_bf16(ATT_BQ, ATT_BK);                 // S = Qe Ke^T  (both K-major)
            const uint32_t idesc_o = make_idesc_bf16(ATT_BQ, ATT_DV, 0, 1);           // O += P V     (V is MN-major)
            mbar_wait(qfull, 0);
            auto issue_pv = [&](int j) {
                const int st = j & 1;
                mbar_wait(pfull, (uint32_t)j & 1u);
                tc_fence_after();
                const uint32_t v_base = smem_u32(sKV + (uint32_t)st * stage_bytes + (uint32_t)p.nchunk * ATT_CHUNK);
                const uint32_t p_base = smem_u32(sP);
                for (int ks = 0; ks < ATT_BK / 16; ++ks) {
                    const uint64_t ad = make_smem_desc_sw128(p_base + (uint32_t)(ks >> 2) * ATT_CHUNK + (ks & 3) * 32, 1024);
                    const uint64_t bd = make_smem_desc_sw128(v_base + ks * 2048, 1024);
                    umma_bf16_ss(tmem_base + ATT_O_COL, ad, bd, idesc_o, (j | ks) ? 1u : 0u);
                }
                umma_commit(&kempty[st]);     // K/E/V stage may be refilled
                umma_commit(pvdone);          // P buffer free, O up to date
            };
            for (int j = 0; j < ntiles; ++j) {
                const int st = j & 1, sb = j & 1;
                const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&kfull[st], ph);
                mbar_wait(&sempty[sb], ph ^ 1u);
                tc_fence_after();
                const uint32_t k_base = smem_u32(sKV + (uint32_t)st * stage_bytes);
                const uint32_t q_base = smem_u32(sQ);
                const uint32_t d_s = tmem_base + (sb ? ATT_S_COL1 : ATT_S_COL0);
                for (int c = 0; c < p.nchunk; ++c)
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = make_smem_desc_sw128(q_base + (uint32_t)c * ATT_CHUNK + ks * 32, 1024);
                        const uint64_t bd = make_smem_desc_sw128(k_base + (uint32_t)c * ATT_CHUNK + ks * 32, 1024);
                        umma_bf16_ss(d_s, ad, bd, idesc_s, (c | ks) ? 1u : 0u);
                    }
                umma_commit(&sfull[sb]);
                if (j > 0) issue_pv(j - 1);
            }
            issue_pv(ntiles - 1);
        }
    } else {
        // ================= softmax / output (thread = query row) =================
        const int r = warp * 32 + lane;                         // row in the tile == TMEM lane
        const int i = qt * ATT_BQ + r;                          // frame index of this query inside its line
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float m_run = -INFINITY, l_run = 0.0f;
        for (int j = 0; j < ntiles; ++j) {
            const int sb = j & 1;
            const uint32_t ph = (uint32_t)(j >> 1) & 1u;
            mbar_wait(&sfull[sb], ph);
            tc_fence_after();
            const uint32_t s_addr = tmem_base + (sb ? ATT_S_COL1 : ATT_S_COL0) + lane_base;
            const int kv0 = j * ATT_BK;
            // pass 1: row maximum over the valid keys of this tile
            float tmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(s_addr + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (kv0 + c * 32 + e < Tk) tmax = fmaxf(tmax, __uint_as_float(v[e]));
            }
            const float m_new = fmaxf(m_run, tmax);
            const float corr = (m_run == -INFINITY) ? 0.0f : ex2_fast((m_run - m_new) * p.scale_log2);
            // P buffer and O are free once P V of the previous tile has completed
            if (j > 0) {
                mbar_wait(pvdone, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
                // rescale the running output (TMEM round trip), 64 columns
#pragma unroll
                for (int c = 0; c < ATT_DV / 16; ++c) {
                    uint32_t o[16];
                    tmem_ld_32x32b_x16(tmem_base + ATT_O_COL + lane_base + c * 16, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
                    tmem_st_32x32b_x16(tmem_base + ATT_O_COL + lane_base + c * 16, o);
                }
                tmem_st_wait();
            }
            // pass 2: probabilities -> bf16 P tile in shared memory (SWIZZLE_128B, K-major)
            float psum = 0.0f;
            const float mb = m_new * p.scale_log2;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(s_addr + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k0 = c * 32 + g * 8 + 2 * e;
                        float p0 = (kv0 + k0 < Tk) ? ex2_fast(fmaf(__uint_as_float(v[g * 8 + 2 * e]), p.scale_log2, -mb)) : 0.0f;
                        float p1 = (kv0 + k0 + 1 < Tk) ? ex2_fast(fmaf(__uint_as_float(v[g * 8 + 2 * e + 1]), p.scale_log2, -mb)) : 0.0f;
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                        // accumulate the row sum from the ROUNDED probabilities so that sum and P V agree
                        psum += __bfloat162float(h2.x) + __bfloat162float(h2.y);
                        pk[e] = *reinterpret_cast<uint32_t *>(&h2);
                    }
                    const int piece = c * 4 + g;                 // 16-byte piece = 8 keys
                    *reinterpret_cast<uint4 *>(sP + (uint32_t)(piece >> 3) * ATT_CHUNK + sw128_offset(r, piece & 7)) =
                        make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            l_run = l_run * corr + psum;
            m_run = m_new;
            // S buffer consumed; P written; O rescaled
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sempty[sb]);
                mbar_arrive(pfull);
            }
        }
        // epilogue: O / l -> ctx
        if (ntiles > 0) {
            mbar_wait(pvdone, (uint32_t)(ntiles - 1) & 1u);
            tc_fence_after();
        }
        const float inv = (l_run > 0.0f) ? 1.0f / l_run : 0.0f;
        __nv_bfloat16 *dst = p.ctx + ((int64_t)row0 + i) * p.d + hd * p.dh;
#pragma unroll
        for (int c = 0; c < ATT_DV / 16; ++c) {
            uint32_t o[16];
            if (ntiles > 0) {
                tmem_ld_32x32b_x16(tmem_base + ATT_O_COL + lane_base + c * 16, o);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) o[e] = 0u;
            }
            if (i < p.T) {
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const int col = c * 16 + e;
                    if (col + 1 < p.dh) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(o[e]) * inv, __uint_as_float(o[e + 1]) * inv);
                        *reinterpret_cast<__nv_bfloat162 *>(dst + col) = h2;
                    } else if (col < p.dh) {
                        dst[col] = __float2bfloat16_rn(__uint_as_float(o[e]) * inv);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

}  // namespace cocr
