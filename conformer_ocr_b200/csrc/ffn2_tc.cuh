// Second generation of the fused feed-forward kernel (ffn_tc.cuh): same data flow, two changes aimed at what bounds the first
// one, the shared-memory bandwidth its SS MMAs share with the TMA writes and with the final epilogue's L1 traffic
// (knock-out runs: no MMAs -11 us, no final-epilogue traffic -11 us, no weight / xn loads -8 us each; G1 takes 1 700 instead
// of 880 cycles while the previous tile's final epilogue runs).
//   * the xn tile is copied ONCE per tile from its TMA slot into tensor memory (tcgen05.cp 128x256b issued by the MMA thread,
//     so it is ordered with the MMAs by the pipe itself) and G1 takes A from TMEM: the first generation re-read the 40 KB tile
//     from shared memory for every hidden chunk (240 KB per tile, 41 % of the MMAs' operand traffic), and an SS MMA costs
//     N/2 + 50 cycles where a TS MMA costs N/2 + 14;
//   * hidden chunks of 64 units (nine for ff = 576) make room for the 80 TMEM columns of that operand:
//     H0 0, H1 64, Y0 128, Y1 272, A 416 (496 of 512).  A G1 is 10 TS MMAs of N = 64, a G2 4 TS MMAs of N = 144 (K = 64).
// Weight chunks: W1_c = 64 rows x (64 + 64 + 32) columns (24 KB of a slot), W2_c = 144 rows x 64 columns (18 KB); four slots.
// Roles, barriers, E1 and the final epilogue are those of ffn_tc.cuh.
#pragma once

#include "ffn_tc.cuh"

namespace cocr {

constexpr int FFN2_HC = 64;                                        // hidden units per chunk
constexpr int FFN2_MAX_NC = 12;                                    // ff <= 768
constexpr int FFN2_WSTAGES = 4;
constexpr uint32_t FFN2_W1_CHUNK = FFN2_HC * 128;                  // one 64-column k-chunk of W1_c (64 rows)
constexpr uint32_t FFN2_W1_BYTES = 3 * FFN2_W1_CHUNK;              // 24 KB
constexpr uint32_t FFN2_W2_BYTES = FFN_D * 128;                    // W2_c: 144 rows x 64 columns, 18 KB
constexpr uint32_t FFN2_SLOT_BYTES = FFN2_W1_BYTES;
constexpr uint32_t FFN2_H_COL0 = 0, FFN2_H_COL1 = 64, FFN2_Y_COL0 = 128, FFN2_Y_COL1 = 272, FFN2_A_COL = 416;
static_assert(FFN2_A_COL + FFN_K1 / 2 <= FFN_TMEM_COLS, "tensor memory budget");
constexpr uint32_t FFN2_SMEM_BYTES = 2 * FFN_A_BYTES + FFN2_WSTAGES * FFN2_SLOT_BYTES + FFN_FLOATS * 4 + 256;

// 128 rows x 32 bytes of a K-major SWIZZLE_128B shared-memory tile -> 8 TMEM columns (the layout a TMEM A operand has)
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__global__ void __launch_bounds__(FFN_THREADS, 1)
k_ffn_fused2(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
             const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ FfnArgs p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                            // [2] xn tiles
    uint8_t *sW = smem + 2 * FFN_A_BYTES;                          // [FFN2_WSTAGES] weight chunks
    float *sb2 = reinterpret_cast<float *>(sW + FFN2_WSTAGES * FFN2_SLOT_BYTES);   // [144]
    float *sln = sb2 + FFN_D;                                      // [4][144] gf, bf, g1, b1n
    uint64_t *bars = reinterpret_cast<uint64_t *>(sln + 4 * FFN_D);
    uint64_t *afull = bars;              // [2]  xn tile loaded
    uint64_t *aempty = bars + 2;         // [2]  xn tile copied into tensor memory
    uint64_t *wfull = bars + 4;          // [FFN2_WSTAGES]
    uint64_t *wempty = bars + 8;         // [FFN2_WSTAGES]
    uint64_t *h1full = bars + 12;        // [2]  G1 done: fp32 hidden chunk in TMEM
    uint64_t *hbfull = bars + 14;        // [2]  bf16 hidden chunk written (E1 warps)
    uint64_t *yfull = bars + 16;         // [2]  last G2 of a tile done
    uint64_t *yempty = bars + 18;        // [2]  Y drained (final warps)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.NC;
    const int my_tiles = ((int)blockIdx.x < p.m_tiles) ? (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_total = my_tiles * NC;                             // hidden chunks this CTA walks through

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0u) __trap();
        prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1);
            mbar_init(&h1full[s], 1); mbar_init(&hbfull[s], FFN_EPI_WARPS);
            mbar_init(&yfull[s], 1); mbar_init(&yempty[s], FFN_FIN_WARPS);
        }
        for (int s = 0; s < FFN2_WSTAGES; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, FFN_TMEM_COLS);
    for (int c = threadIdx.x; c < FFN_D; c += FFN_THREADS) {
        sb2[c] = p.b2 ? p.b2[c] : 0.0f;
        sln[c] = p.gf ? p.gf[c] : 1.0f;
        sln[FFN_D + c] = p.bf ? p.bf[c] : 0.0f;
        sln[2 * FFN_D + c] = p.g1 ? p.g1[c] : 1.0f;
        sln[3 * FFN_D + c] = p.b1n ? p.b1n[c] : 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: xn tiles and weight chunks in the MMA's consumption order =================
        if (lane == 0) {
            uint32_t wq = 0;                                       // weight chunks issued
            auto load_w = [&](bool first, int c) {
                const uint32_t sl = wq % FFN2_WSTAGES;
                mbar_wait(&wempty[sl], ((wq / FFN2_WSTAGES) & 1u) ^ 1u);
                uint8_t *dst = sW + sl * FFN2_SLOT_BYTES;
                if (first) {
                    mbar_expect_tx(&wfull[sl], FFN2_W1_BYTES);
                    for (int kc = 0; kc < 3; ++kc) tma_load_2d(dst + (uint32_t)kc * FFN2_W1_CHUNK, &tmW1, &wfull[sl], kc * 64, c * FFN2_HC);
                } else {
                    mbar_expect_tx(&wfull[sl], FFN2_W2_BYTES);
                    tma_load_2d(dst, &tmW2, &wfull[sl], 0, c * FFN_D);
                }
                ++wq;
            };
            auto load_a = [&](int t) {                             // t = local tile index
                const int sl = t & 1;
                mbar_wait(&aempty[sl], (((uint32_t)t >> 1) & 1u) ^ 1u);
                mbar_expect_tx(&afull[sl], FFN_A_BYTES);
                const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * 128;
                for (int kc = 0; kc < 3; ++kc) tma_load_2d(sA + (uint32_t)sl * FFN_A_BYTES + (uint32_t)kc * FFN_A_CHUNK, &tmX, &afull[sl], kc * 64, row0);
            };
            if (n_total > 0) { load_a(0); load_w(true, 0); }
            if (my_tiles > 1) load_a(1);
            for (int n = 0; n < n_total; ++n) {
                if (n + 1 < n_total) {
                    const int c1 = (n + 1) % NC;
                    load_w(true, c1);
                    if (c1 == 0 && (n + 1) / NC + 1 < my_tiles) load_a((n + 1) / NC + 1);
                }
                load_w(false, n % NC);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc1 = make_idesc_bf16(128, FFN2_HC), idesc2 = make_idesc_bf16(128, FFN_D);
        uint32_t wq = 0;
        int t1 = 0, c1 = 0, t2 = 0, c2 = 0;                        // (tile, chunk) of G1 and of G2 advance separately
        auto g1 = [&](int n) {                                     // H(n & 1) = xn_tile (TMEM) * W1_c^T
            const int t = t1, c = c1;
            if (++c1 == NC) { c1 = 0; ++t1; }
            const int sa = t & 1;
            if (c == 0) {
                // the tile's xn rows -> tensor memory, once: every G1 of the previous tile has been issued, and the pipe
                // executes tcgen05.cp and tcgen05.mma in issue order, so the operand they read is not overwritten early
                mbar_wait(&afull[sa], ((uint32_t)t >> 1) & 1u);
                tc_fence_after();
                const uint32_t a_base = smem_u32(sA + (uint32_t)sa * FFN_A_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < FFN_K1 / 16; ++k)
                        tmem_cp_128x256b(tmem_base + FFN2_A_COL + (uint32_t)k * 8u,
                                         make_smem_desc_sw128(a_base + (uint32_t)(k >> 2) * FFN_A_CHUNK + (uint32_t)(k & 3) * 32, 1024));
                    umma_commit(&aempty[sa]);                      // the TMA slot may be refilled once the copies have run
                }
                __syncwarp();
            }
            const uint32_t sl = wq % FFN2_WSTAGES;
            mbar_wait(&wfull[sl], (wq / FFN2_WSTAGES) & 1u);
            ++wq;
            tc_fence_after();
            const uint32_t b_base = smem_u32(sW + sl * FFN2_SLOT_BYTES);
            const uint32_t d_tmem = tmem_base + ((n & 1) ? FFN2_H_COL1 : FFN2_H_COL0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < FFN_K1 / 16; ++k) {
                    if (p.dbg & 8) break;
                    umma_bf16_ts(d_tmem, tmem_base + FFN2_A_COL + (uint32_t)k * 8u,
                                 make_smem_desc_sw128(b_base + (uint32_t)(k >> 2) * FFN2_W1_CHUNK + (uint32_t)(k & 3) * 32, 1024), idesc1, k ? 1u : 0u);
                }
                umma_commit(&wempty[sl]);
                umma_commit(&h1full[n & 1]);
            }
            __syncwarp();
        };
        auto g2 = [&](int n) {                                     // Y(t & 1) += bf16 H(n & 1) (TMEM) * W2_c^T
            const int t = t2, c = c2;
            if (++c2 == NC) { c2 = 0; ++t2; }
            const uint32_t sl = wq % FFN2_WSTAGES;
            mbar_wait(&wfull[sl], (wq / FFN2_WSTAGES) & 1u);
            ++wq;
            mbar_wait(&hbfull[n & 1], ((uint32_t)n >> 1) & 1u);
            if (c == 0) mbar_wait(&yempty[t & 1], (((uint32_t)t >> 1) & 1u) ^ 1u);   // Y of tile t-2 has been drained
            tc_fence_after();
            const uint32_t b_base = smem_u32(sW + sl * FFN2_SLOT_BYTES);
            const uint32_t a_tmem = tmem_base + ((n & 1) ? FFN2_H_COL1 : FFN2_H_COL0);
            const uint32_t y_tmem = tmem_base + ((t & 1) ? FFN2_Y_COL1 : FFN2_Y_COL0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < FFN2_HC / 16; ++k) {
                    if (p.dbg & 8) break;
                    // packed hidden pairs: k-steps 0, 1 in columns 0..15 (lower half), 2, 3 in columns 48..63 (upper half)
                    umma_bf16_ts(y_tmem, a_tmem + (uint32_t)(k < 2 ? k * 8 : 48 + (k - 2) * 8), make_smem_desc_sw128(b_base + (uint32_t)k * 32, 1024), idesc2,
                                 (c | k) ? 1u : 0u);
                }
                umma_commit(&wempty[sl]);
                if (c == NC - 1) umma_commit(&yfull[t & 1]);
            }
            __syncwarp();
        };
        if (n_total > 0) g1(0);
        for (int n = 0; n < n_total; ++n) {
            if (n + 1 < n_total) g1(n + 1);
            g2(n);
        }
    } else if (warp < 2 + FFN_EPI_WARPS) {
        // ================= E1 warps: H <- bf16(silu(H + b1)), thread = row, 32 columns =================
        const int q = warp & 3;                                    // TMEM lane quarter
        const int hf = (warp - 2) >> 2;                            // half of the hidden columns
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int t = 0, c = 0;
        for (int n = 0; n < n_total; ++n) {
            if (c == 0) {
                // request this tile's rows of the residual stream into L2: the final epilogue reads them a tile from now
                const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
                const int64_t rows = min((int64_t)128, (int64_t)p.M - tile * 128);
                const char *xb = reinterpret_cast<const char *>(p.x + tile * 128 * FFN_D);
                for (int64_t o = (int64_t)(threadIdx.x - 64) * 128; o < rows * FFN_D * 4; o += (int64_t)FFN_EPI_WARPS * 32 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + o));
            }
            const uint32_t hcol = tmem_base + ((n & 1) ? FFN2_H_COL1 : FFN2_H_COL0) + lane_base;
            mbar_wait(&h1full[n & 1], ((uint32_t)n >> 1) & 1u);
            tc_fence_after();
            uint32_t r0[32];
            tmem_ld_32x32b_x32(hcol + (uint32_t)(hf * 32), r0);
            tmem_ld_wait_dep(r0);
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float2 sv = silu_from_half(make_float2(__uint_as_float(r0[2 * e]), __uint_as_float(r0[2 * e + 1])));
                pk[e] = pack_bf16x2(sv.x, sv.y);
            }
            // each warp packs into columns it has read itself (lower half -> 0..15, upper half -> 48..63)
            tmem_st_32x32b_x16n(hcol + (uint32_t)(hf * 48), pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hbfull[n & 1]);
            if (++c == NC) { c = 0; ++t; }
        }
    } else {
        // ================= final-epilogue warps =================
        // x <- x + alpha * (Y + b2) [-> LN_final]; xn = bf16(LN_next(x)); fragment layout (see EPI_RESID in gemm_tc.cuh).
        // Their own warps: the x / xn traffic of a tile (184 KB through the LSU) and the LayerNorm arithmetic overlap the
        // next tile's GEMMs and E1 instead of stalling them; Y is double-buffered, so they have a whole tile period for it.
        // Two warps per lane quarter, 16 rows each, ONE row per lane quad at a time (two passes over the warp's Y rows: the
        // 16x256b load returns rows r and r+8 together and the pass keeps one of them): half the registers of the two-row
        // form, which is what lets eight of these warps sit next to the E1 warps.
        const int q = warp & 3;
        const int hf = (warp - 2 - FFN_EPI_WARPS) >> 2;
        int tr_n = GEMM_TRACE_EVENTS;
        for (int t = 0; t < my_tiles; ++t) {
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                constexpr int NG = FFN_NG;
                const int r_in = lane >> 2, qd = lane & 3;
                const float inv_n = 1.0f / (float)FFN_D;
                const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
                const int64_t row = tile * 128 + q * 32 + hf * 16 + pass * 8 + r_in;
                const bool ok = row < p.M && !(p.dbg & 4);
                float2 v[NG];
                float *xr = p.x + row * FFN_D + 4 * qd;
#pragma unroll
                for (int gp = 0; gp < NG / 2; ++gp) {
                    const float4 ta = ok ? *reinterpret_cast<const float4 *>(xr + 16 * gp) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[2 * gp] = make_float2(ta.x, ta.y); v[2 * gp + 1] = make_float2(ta.z, ta.w);
                }
                if (pass == 0) {
                    mbar_wait(&yfull[t & 1], ((uint32_t)t >> 1) & 1u);
                    tc_fence_after();
                }
                const uint32_t taddr = tmem_base + ((t & 1) ? FFN2_Y_COL1 : FFN2_Y_COL0) + ((uint32_t)(q * 32 + hf * 16) << 16);
#pragma unroll
                for (int blk = 0; blk < NG / 2; ++blk) {
                    uint32_t t1[8];
                    tmem_ld_16x256b_x2(taddr + 16u * blk, t1);
                    tmem_ld_wait_dep(t1);
                    if (blk == NG / 2 - 1 && pass == 1) {          // Y is in registers: tile t+2 may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&yempty[t & 1]);
                    }
#pragma unroll
                    for (int gg = 0; gg < 2; ++gg) {
                        const int g = 2 * blk + gg;
                        const float2 b2 = *reinterpret_cast<const float2 *>(sb2 + 8 * g + 2 * qd);
                        const float y0 = __uint_as_float(pass ? t1[4 * gg + 2] : t1[4 * gg]), y1 = __uint_as_float(pass ? t1[4 * gg + 3] : t1[4 * gg + 1]);
                        v[g].x = fmaf(p.alpha, y0 + b2.x, v[g].x);
                        v[g].y = fmaf(p.alpha, y1 + b2.y, v[g].y);
                    }
                }
                // LayerNorm statistics (two passes in registers; the row lives in one quad of lanes)
                auto stats = [&](float &mean, float &rstd) {
                    float sA = 0.f;
#pragma unroll
                    for (int g = 0; g < NG; ++g) sA += v[g].x + v[g].y;
                    sA += __shfl_xor_sync(0xffffffffu, sA, 1);
                    sA += __shfl_xor_sync(0xffffffffu, sA, 2);
                    mean = sA * inv_n;
                    float qA = 0.f;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const float d0 = v[g].x - mean, d1 = v[g].y - mean;
                        qA = fmaf(d0, d0, qA); qA = fmaf(d1, d1, qA);
                    }
                    qA += __shfl_xor_sync(0xffffffffu, qA, 1);
                    qA += __shfl_xor_sync(0xffffffffu, qA, 2);
                    rstd = rsqrtf(qA * inv_n + 1e-5f);
                };
                if (p.gf) {   // block-final LayerNorm replaces the residual stream (encoder.py:99)
                    float m, r;
                    stats(m, r);
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const float2 gg = *reinterpret_cast<const float2 *>(sln + 8 * g + 2 * qd);
                        const float2 bb = *reinterpret_cast<const float2 *>(sln + FFN_D + 8 * g + 2 * qd);
                        v[g].x = (v[g].x - m) * r * gg.x + bb.x; v[g].y = (v[g].y - m) * r * gg.y + bb.y;
                    }
                }
#pragma unroll
                for (int gp = 0; gp < NG / 2; ++gp)
                    if (ok) *reinterpret_cast<float4 *>(xr + 16 * gp) = make_float4(v[2 * gp].x, v[2 * gp].y, v[2 * gp + 1].x, v[2 * gp + 1].y);
                float m = 0.f, r = 1.f;
                if (p.g1) stats(m, r);
                uint32_t pa[NG];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float a0 = v[g].x, a1 = v[g].y;
                    if (p.g1) {
                        const float2 gg = *reinterpret_cast<const float2 *>(sln + 2 * FFN_D + 8 * g + 2 * qd);
                        const float2 bb = *reinterpret_cast<const float2 *>(sln + 3 * FFN_D + 8 * g + 2 * qd);
                        a0 = (a0 - m) * r * gg.x + bb.x; a1 = (a1 - m) * r * gg.y + bb.y;
                    }
                    pa[g] = pack_bf16x2(a0, a1);
                }
                // neighbouring lanes swap one packed pair: every lane stores 8 bytes, a quad the 32 contiguous bytes of two groups
                __nv_bfloat16 *na = p.xn_out + row * p.ldo;
                const bool odd = qd & 1;
#pragma unroll
                for (int g = 0; g < NG; g += 2) {
                    const uint32_t send = odd ? pa[g] : pa[g + 1];
                    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
                    const int col = odd ? 8 * (g + 1) + 2 * (qd - 1) : 8 * g + 2 * qd;
                    const uint2 o = odd ? make_uint2(recv, pa[g + 1]) : make_uint2(pa[g], recv);
                    if (ok) *reinterpret_cast<uint2 *>(na + col) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FFN_TMEM_COLS);
    }
}

}  // namespace cocr
