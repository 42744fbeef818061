// Fused macaron feed-forward module on tcgen05 tensor cores (feed_forward.py:30-46, encoder.py:91-99):
//
//     x  <- x + alpha * (W2 * silu(W1 * xn + b1) + b2)      [-> LayerNorm_final(x)]
//     xn <- bf16(LayerNorm_next(x))
//
// in ONE kernel: the hidden activation (rows x ff, 113 MB per module at the benchmark size) never leaves the SM.  The two
// separate GEMMs wrote and re-read it through HBM and were bound by exactly that traffic (97 us per module against
// 26 us of unavoidable x / xn traffic).
//
// Built for encoder_dim 144 and ff = 144 * NC.  A persistent CTA takes 128-row tiles; per tile and hidden chunk c (144
// hidden units):
//     G1(c):  H_c  = xn_tile (128 x 144, smem)   * W1_c^T (144 x 144, smem)  -> TMEM accumulator (fp32, double-buffered)
//     E1(c):  H_c <- bf16(silu(H_c + b1_c)) written IN PLACE as packed pairs over the first 72 columns of the accumulator
//     G2(c):  Y   += H_c (TMEM A operand)         * W2_c^T (144 x 144, smem)  -> TMEM accumulator Y
// and after the last chunk the residual / LayerNorm epilogue of EPI_RESID (gemm_tc.cuh) drains Y.
// The weight chunks (2 * NC * 41 KB per tile) stream from L2 through a two-slot ring in the order the MMAs consume
// them; the xn tile is double-buffered so the next tile's first G1 overlaps the current tile's tail.
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (G1 runs one chunk ahead of G2), warps 2..9 = E1
// (two per TMEM lane quarter, 72 of the 144 hidden columns each), warps 10..17 = final epilogue (two per lane quarter).
// TMEM (512 columns): H0 at 0, H1 at 144, Y at 288.
#pragma once

#include "gemm_tc.cuh"

namespace cocr {

struct FfnArgs {
    int M;                      // rows
    int NC;                     // hidden chunks of 144 (ff = 144 * NC)
    int m_tiles;
    const float *b1;            // (ff) FFN-up bias, pre-scaled by 1/2 like the weights (silu from half, gemm_tc.cuh)
    const float *b2;            // (144)
    float *x;                   // (M, 144) fp32 residual stream in the private layout of GemmArgs::x, updated in place
    __nv_bfloat16 *xn_out;      // (M, ldo) bf16 LayerNorm_next(x)
    int ldo;
    float alpha;                // 0.5 for the half-step modules
    const float *gf, *bf;       // block-final LayerNorm applied to x itself (or null)
    const float *g1, *b1n;      // LayerNorm of the next module -> xn_out (null: xn_out = bf16(x))
    long long *trace;           // debug: (tag, clock) pairs of CTA 0 (COCR_FFN_TRACE=<file>), else null
    int dbg;                    // experiments (COCR_FFN_DBG, results are garbage): 1 no weight loads, 2 no silu math, 4 no final
                                //    epilogue memory traffic, 8 no MMAs, 16 no xn tile loads
};

constexpr int FFN_D = 144;
constexpr int FFN_NG = FFN_D / 8;                                 // 8-column groups per row
constexpr int FFN_MAX_NC = 8;
constexpr int FFN_EPI_WARPS = 8;                                  // E1 warps (two per TMEM lane quarter)
constexpr int FFN_FIN_WARPS = 8;                                  // final-epilogue warps (two per lane quarter, 16 rows each)
constexpr int FFN_THREADS = 64 + 32 * (FFN_EPI_WARPS + FFN_FIN_WARPS);
constexpr int FFN_WSTAGES = 2;
constexpr uint32_t FFN_A_CHUNK = 128 * 128;                       // one 64-column k-chunk of the xn tile (SWIZZLE_128B)
constexpr uint32_t FFN_A_BYTES = 3 * FFN_A_CHUNK;                 // k = 64 + 64 + 16 (the last chunk is mostly padding)
constexpr uint32_t FFN_W_CHUNK = FFN_D * 128;                     // one 64-column k-chunk of a 144-row weight chunk
constexpr uint32_t FFN_W_BYTES = 3 * FFN_W_CHUNK;
constexpr uint32_t FFN_H_COL0 = 0, FFN_H_COL1 = 144, FFN_Y_COL = 288;
constexpr uint32_t FFN_TMEM_COLS = 512;
constexpr uint32_t FFN_FLOATS = FFN_MAX_NC * FFN_D + FFN_D + 4 * FFN_D;   // b1 | b2 | gf bf g1 b1n
constexpr uint32_t FFN_SMEM_BYTES = 2 * FFN_A_BYTES + FFN_WSTAGES * FFN_W_BYTES + FFN_FLOATS * 4 + 256;

__global__ void __launch_bounds__(FFN_THREADS, 1)
k_ffn_fused(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
            const __grid_constant__ CUtensorMap tmW2, const FfnArgs p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                            // [2] xn tiles
    uint8_t *sW = smem + 2 * FFN_A_BYTES;                          // [FFN_WSTAGES] weight chunks
    float *sb1 = reinterpret_cast<float *>(sW + FFN_WSTAGES * FFN_W_BYTES);   // [NC * 144]
    float *sb2 = sb1 + FFN_MAX_NC * FFN_D;                         // [144]
    float *sln = sb2 + FFN_D;                                      // [4][144] gf, bf, g1, b1n
    uint64_t *bars = reinterpret_cast<uint64_t *>(sln + 4 * FFN_D);
    uint64_t *afull = bars;              // [2]  xn tile loaded
    uint64_t *aempty = bars + 2;         // [2]  every G1 of the tile has run
    uint64_t *wfull = bars + 4;          // [FFN_WSTAGES]
    uint64_t *wempty = bars + 6;         // [FFN_WSTAGES]
    uint64_t *h1full = bars + 8;         // [2]  G1 done: fp32 hidden chunk in TMEM
    uint64_t *hbfull = bars + 10;        // [2]  bf16 hidden chunk written (epilogue warps)
    uint64_t *yfull = bars + 12;         // [1]  last G2 of a tile done
    uint64_t *yempty = bars + 13;        // [1]  Y drained (epilogue warps)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 14);

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
        }
        for (int s = 0; s < FFN_WSTAGES; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        mbar_init(yfull, 1); mbar_init(yempty, FFN_FIN_WARPS);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, FFN_TMEM_COLS);
    for (int c = threadIdx.x; c < NC * FFN_D; c += FFN_THREADS) sb1[c] = p.b1[c];
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
            int tr_n = 0;
            auto load_w = [&](const CUtensorMap *tm, int c) {
                const uint32_t sl = wq % FFN_WSTAGES;
                GEMM_TRACE(2, 100 + c);
                mbar_wait(&wempty[sl], ((wq / FFN_WSTAGES) & 1u) ^ 1u);
                GEMM_TRACE(2, 110 + c);
                if (p.dbg & 1) { mbar_arrive(&wfull[sl]); ++wq; return; }
                mbar_expect_tx(&wfull[sl], FFN_W_BYTES);
                for (int kc = 0; kc < 3; ++kc) tma_load_2d(sW + sl * FFN_W_BYTES + (uint32_t)kc * FFN_W_CHUNK, tm, &wfull[sl], kc * 64, c * FFN_D);
                ++wq;
            };
            auto load_a = [&](int t) {                             // t = local tile index
                const int sl = t & 1;
                mbar_wait(&aempty[sl], (((uint32_t)t >> 1) & 1u) ^ 1u);
                GEMM_TRACE(2, 150);
                if (p.dbg & 16) { mbar_arrive(&afull[sl]); return; }
                mbar_expect_tx(&afull[sl], FFN_A_BYTES);
                const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * 128;
                for (int kc = 0; kc < 3; ++kc) tma_load_2d(sA + (uint32_t)sl * FFN_A_BYTES + (uint32_t)kc * FFN_A_CHUNK, &tmX, &afull[sl], kc * 64, row0);
            };
            // the xn tiles run one tile ahead of the weights (their slot frees a whole tile earlier than it is needed, and
            // a tile requested only when its first G1 is due would expose the HBM latency)
            if (n_total > 0) { load_a(0); load_w(&tmW1, 0); }
            if (my_tiles > 1) load_a(1);
            for (int n = 0; n < n_total; ++n) {
                if (n + 1 < n_total) {
                    const int c1 = (n + 1) % NC;
                    load_w(&tmW1, c1);
                    if (c1 == 0 && (n + 1) / NC + 1 < my_tiles) load_a((n + 1) / NC + 1);
                }
                load_w(&tmW2, n % NC);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_bf16(128, FFN_D);
        uint32_t wq = 0;
        int tr_n = (lane == 0) ? 0 : GEMM_TRACE_EVENTS;
        auto g1 = [&](int n) {                                     // H(n & 1) = xn_tile * W1_c^T
            const int t = n / NC, c = n - t * NC;
            const int sa = t & 1;
            GEMM_TRACE(1, 200 + c);
            if (c == 0) mbar_wait(&afull[sa], ((uint32_t)t >> 1) & 1u);
            const uint32_t sl = wq % FFN_WSTAGES;
            mbar_wait(&wfull[sl], (wq / FFN_WSTAGES) & 1u);
            ++wq;
            GEMM_TRACE(1, 210 + c);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sA + (uint32_t)sa * FFN_A_BYTES), b_base = smem_u32(sW + sl * FFN_W_BYTES);
            const uint32_t d_tmem = tmem_base + ((n & 1) ? FFN_H_COL1 : FFN_H_COL0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < FFN_D / 16; ++k) {
                    if (p.dbg & 8) break;
                    const uint32_t off = (uint32_t)(k >> 2), ks = (uint32_t)(k & 3);
                    umma_bf16_ss(d_tmem, make_smem_desc_sw128(a_base + off * FFN_A_CHUNK + ks * 32, 1024),
                                 make_smem_desc_sw128(b_base + off * FFN_W_CHUNK + ks * 32, 1024), idesc, k ? 1u : 0u);
                }
                umma_commit(&wempty[sl]);
                umma_commit(&h1full[n & 1]);
                if (c == NC - 1) umma_commit(&aempty[sa]);
            }
            __syncwarp();
            GEMM_TRACE(1, 220 + c);
        };
        auto g2 = [&](int n) {                                     // Y += bf16 H(n & 1) (TMEM) * W2_c^T
            const int t = n / NC, c = n - t * NC;
            const uint32_t sl = wq % FFN_WSTAGES;
            GEMM_TRACE(1, 230 + c);
            mbar_wait(&wfull[sl], (wq / FFN_WSTAGES) & 1u);
            ++wq;
            GEMM_TRACE(1, 240 + c);
            mbar_wait(&hbfull[n & 1], ((uint32_t)n >> 1) & 1u);
            GEMM_TRACE(1, 250 + c);
            if (c == 0) mbar_wait(yempty, ((uint32_t)t & 1u) ^ 1u);   // the previous tile's Y has been drained
            GEMM_TRACE(1, 260 + c);
            tc_fence_after();
            const uint32_t b_base = smem_u32(sW + sl * FFN_W_BYTES);
            const uint32_t a_tmem = tmem_base + ((n & 1) ? FFN_H_COL1 : FFN_H_COL0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < FFN_D / 16; ++k) {
                    if (p.dbg & 8) break;
                    const uint32_t off = (uint32_t)(k >> 2), ks = (uint32_t)(k & 3);
                    umma_bf16_ts(tmem_base + FFN_Y_COL, a_tmem + (uint32_t)k * 8u, make_smem_desc_sw128(b_base + off * FFN_W_CHUNK + ks * 32, 1024), idesc,
                                 (c | k) ? 1u : 0u);
                }
                umma_commit(&wempty[sl]);
                if (c == NC - 1) umma_commit(yfull);
            }
            __syncwarp();
            GEMM_TRACE(1, 270 + c);
        };
        if (n_total > 0) g1(0);
        for (int n = 0; n < n_total; ++n) {
            if (n + 1 < n_total) g1(n + 1);
            g2(n);
        }
    } else if (warp < 2 + FFN_EPI_WARPS) {
        // ================= E1 warps =================
        const int q = warp & 3;                                    // TMEM lane quarter
        const int hf = (warp - 2) >> 2;                            // half of the hidden columns
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int tr_n = (threadIdx.x == 64) ? 0 : GEMM_TRACE_EVENTS;
        for (int n = 0; n < n_total; ++n) {
            const int t = n / NC, c = n - t * NC;
            GEMM_TRACE(0, 300 + c);
            if (c == 0) {
                // request this tile's rows of the residual stream into L2: the final epilogue reads them ~4 chunks from now
                const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
                const int64_t rows = min((int64_t)128, (int64_t)p.M - tile * 128);
                const char *xb = reinterpret_cast<const char *>(p.x + tile * 128 * FFN_D);
                for (int64_t o = (int64_t)(threadIdx.x - 64) * 128; o < rows * FFN_D * 4; o += (int64_t)FFN_EPI_WARPS * 32 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + o));
            }
            // ---- E1: H <- bf16(silu(H + b1)), thread = row, 72 columns ----
            {
                const uint32_t hcol = tmem_base + ((n & 1) ? FFN_H_COL1 : FFN_H_COL0) + lane_base;
                mbar_wait(&h1full[n & 1], ((uint32_t)n >> 1) & 1u);
                GEMM_TRACE(0, 310 + c);
                tc_fence_after();
                uint32_t r0[32], r1[32], r2[8];
                tmem_ld_32x32b_x32(hcol + (uint32_t)(hf * 72), r0);
                tmem_ld_32x32b_x32(hcol + (uint32_t)(hf * 72 + 32), r1);
                tmem_ld_32x32b_x8(hcol + (uint32_t)(hf * 72 + 64), r2);
                tmem_ld_wait_dep(r0); tmem_ld_wait_dep(r1); tmem_ld_wait_dep(r2);
                // the packed output of the upper half lands on columns 36..71, which the lower half has just read as fp32
                named_bar_sync(1 + q, 64);
                GEMM_TRACE(0, 320 + c);
                const float *bb = sb1 + c * FFN_D + hf * 72;
                uint32_t pk[36];
#pragma unroll
                for (int e4 = 0; e4 < 18; ++e4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(bb + 4 * e4);
                    float2 v0, v1;
                    if (e4 < 8) { v0 = make_float2(__uint_as_float(r0[4 * e4]), __uint_as_float(r0[4 * e4 + 1])); v1 = make_float2(__uint_as_float(r0[4 * e4 + 2]), __uint_as_float(r0[4 * e4 + 3])); }
                    else if (e4 < 16) { v0 = make_float2(__uint_as_float(r1[4 * e4 - 32]), __uint_as_float(r1[4 * e4 - 31])); v1 = make_float2(__uint_as_float(r1[4 * e4 - 30]), __uint_as_float(r1[4 * e4 - 29])); }
                    else { v0 = make_float2(__uint_as_float(r2[4 * e4 - 64]), __uint_as_float(r2[4 * e4 - 63])); v1 = make_float2(__uint_as_float(r2[4 * e4 - 62]), __uint_as_float(r2[4 * e4 - 61])); }
                    float2 s0 = fadd2(v0, make_float2(b4.x, b4.y)), s1 = fadd2(v1, make_float2(b4.z, b4.w));
                    if (!(p.dbg & 2)) { s0 = silu_from_half(s0); s1 = silu_from_half(s1); }
                    pk[2 * e4] = pack_bf16x2(s0.x, s0.y);
                    pk[2 * e4 + 1] = pack_bf16x2(s1.x, s1.y);
                }
                GEMM_TRACE(0, 340 + c);
                tmem_st_32x32b_x32(hcol + (uint32_t)(hf * 36), pk);
                tmem_st_32x32b_x4_at<32>(hcol + (uint32_t)(hf * 36 + 32), pk);
                tmem_st_wait();
                GEMM_TRACE(0, 350 + c);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hbfull[n & 1]);
                GEMM_TRACE(0, 330 + c);
            }
        }
    } else {
        // ================= final-epilogue warps =================
        // x <- x + alpha * (Y + b2) [-> LN_final]; xn = bf16(LN_next(x)); fragment layout (see EPI_RESID in gemm_tc.cuh),
        // two warps per lane quarter, 16 rows each.  Their own warps: the x / xn traffic of a tile (184 KB through the LSU)
        // and the LayerNorm arithmetic overlap the next tile's GEMMs and E1 instead of stalling them, and Y is released
        // as soon as it sits in registers (x is loaded before Y is complete, so nothing but the TMEM loads is in between).
        const int q = warp & 3;
        const int hf = (warp - 2 - FFN_EPI_WARPS) >> 2;
        int tr_n = GEMM_TRACE_EVENTS;
        for (int t = 0; t < my_tiles; ++t) {
            {
                constexpr int NG = FFN_NG;
                const int r_in = lane >> 2, qd = lane & 3;
                const float inv_n = 1.0f / (float)FFN_D;
                const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
                const int64_t rowA = tile * 128 + q * 32 + hf * 16 + r_in, rowB = rowA + 8;
                const bool okA = rowA < p.M && !(p.dbg & 4), okB = rowB < p.M && !(p.dbg & 4);
                float2 va[NG], vb[NG];
                float *xa = p.x + rowA * FFN_D + 4 * qd, *xb = p.x + rowB * FFN_D + 4 * qd;
#pragma unroll
                for (int gp = 0; gp < NG / 2; ++gp) {
                    const float4 ta = okA ? *reinterpret_cast<const float4 *>(xa + 16 * gp) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 tb = okB ? *reinterpret_cast<const float4 *>(xb + 16 * gp) : make_float4(0.f, 0.f, 0.f, 0.f);
                    va[2 * gp] = make_float2(ta.x, ta.y); va[2 * gp + 1] = make_float2(ta.z, ta.w);
                    vb[2 * gp] = make_float2(tb.x, tb.y); vb[2 * gp + 1] = make_float2(tb.z, tb.w);
                }
                mbar_wait(yfull, (uint32_t)t & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + FFN_Y_COL + ((uint32_t)(q * 32 + hf * 16) << 16);
#pragma unroll
                for (int blk = 0; blk < NG / 2; ++blk) {
                    uint32_t t1[8];
                    tmem_ld_16x256b_x2(taddr + 16u * blk, t1);
                    tmem_ld_wait_dep(t1);
                    if (blk == NG / 2 - 1) {                       // Y is in registers: the next tile's first G2 may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(yempty);
                    }
#pragma unroll
                    for (int gg = 0; gg < 2; ++gg) {
                        const int g = 2 * blk + gg;
                        const float2 b2 = *reinterpret_cast<const float2 *>(sb2 + 8 * g + 2 * qd);
                        va[g].x = fmaf(p.alpha, __uint_as_float(t1[4 * gg]) + b2.x, va[g].x);
                        va[g].y = fmaf(p.alpha, __uint_as_float(t1[4 * gg + 1]) + b2.y, va[g].y);
                        vb[g].x = fmaf(p.alpha, __uint_as_float(t1[4 * gg + 2]) + b2.x, vb[g].x);
                        vb[g].y = fmaf(p.alpha, __uint_as_float(t1[4 * gg + 3]) + b2.y, vb[g].y);
                    }
                }
                // LayerNorm statistics of rows A and B (two passes in registers; a row lives in one quad of lanes)
                auto stats = [&](float &meanA, float &rstdA, float &meanB, float &rstdB) {
                    float sA = 0.f, sB = 0.f;
#pragma unroll
                    for (int g = 0; g < NG; ++g) { sA += va[g].x + va[g].y; sB += vb[g].x + vb[g].y; }
                    sA += __shfl_xor_sync(0xffffffffu, sA, 1); sB += __shfl_xor_sync(0xffffffffu, sB, 1);
                    sA += __shfl_xor_sync(0xffffffffu, sA, 2); sB += __shfl_xor_sync(0xffffffffu, sB, 2);
                    meanA = sA * inv_n; meanB = sB * inv_n;
                    float qA = 0.f, qB = 0.f;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        float d0 = va[g].x - meanA, d1 = va[g].y - meanA;
                        qA = fmaf(d0, d0, qA); qA = fmaf(d1, d1, qA);
                        d0 = vb[g].x - meanB; d1 = vb[g].y - meanB;
                        qB = fmaf(d0, d0, qB); qB = fmaf(d1, d1, qB);
                    }
                    qA += __shfl_xor_sync(0xffffffffu, qA, 1); qB += __shfl_xor_sync(0xffffffffu, qB, 1);
                    qA += __shfl_xor_sync(0xffffffffu, qA, 2); qB += __shfl_xor_sync(0xffffffffu, qB, 2);
                    rstdA = rsqrtf(qA * inv_n + 1e-5f); rstdB = rsqrtf(qB * inv_n + 1e-5f);
                };
                if (p.gf) {   // block-final LayerNorm replaces the residual stream (encoder.py:99)
                    float mA, rA, mB, rB;
                    stats(mA, rA, mB, rB);
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const float2 gg = *reinterpret_cast<const float2 *>(sln + 8 * g + 2 * qd);
                        const float2 bb = *reinterpret_cast<const float2 *>(sln + FFN_D + 8 * g + 2 * qd);
                        va[g].x = (va[g].x - mA) * rA * gg.x + bb.x; va[g].y = (va[g].y - mA) * rA * gg.y + bb.y;
                        vb[g].x = (vb[g].x - mB) * rB * gg.x + bb.x; vb[g].y = (vb[g].y - mB) * rB * gg.y + bb.y;
                    }
                }
#pragma unroll
                for (int gp = 0; gp < NG / 2; ++gp) {
                    if (okA) *reinterpret_cast<float4 *>(xa + 16 * gp) = make_float4(va[2 * gp].x, va[2 * gp].y, va[2 * gp + 1].x, va[2 * gp + 1].y);
                    if (okB) *reinterpret_cast<float4 *>(xb + 16 * gp) = make_float4(vb[2 * gp].x, vb[2 * gp].y, vb[2 * gp + 1].x, vb[2 * gp + 1].y);
                }
                float mA = 0.f, rA = 1.f, mB = 0.f, rB = 1.f;
                if (p.g1) stats(mA, rA, mB, rB);
                uint32_t pa[NG], pb[NG];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float a0 = va[g].x, a1 = va[g].y, c0 = vb[g].x, c1 = vb[g].y;
                    if (p.g1) {
                        const float2 gg = *reinterpret_cast<const float2 *>(sln + 2 * FFN_D + 8 * g + 2 * qd);
                        const float2 bb = *reinterpret_cast<const float2 *>(sln + 3 * FFN_D + 8 * g + 2 * qd);
                        a0 = (a0 - mA) * rA * gg.x + bb.x; a1 = (a1 - mA) * rA * gg.y + bb.y;
                        c0 = (c0 - mB) * rB * gg.x + bb.x; c1 = (c1 - mB) * rB * gg.y + bb.y;
                    }
                    pa[g] = pack_bf16x2(a0, a1);
                    pb[g] = pack_bf16x2(c0, c1);
                }
                __nv_bfloat16 *na = p.xn_out + rowA * p.ldo, *nb = p.xn_out + rowB * p.ldo;
                const bool odd = qd & 1;
#pragma unroll
                for (int g = 0; g < NG; g += 2) {
                    const uint32_t sendA = odd ? pa[g] : pa[g + 1], sendB = odd ? pb[g] : pb[g + 1];
                    const uint32_t recvA = __shfl_xor_sync(0xffffffffu, sendA, 1), recvB = __shfl_xor_sync(0xffffffffu, sendB, 1);
                    const int col = odd ? 8 * (g + 1) + 2 * (qd - 1) : 8 * g + 2 * qd;
                    const uint2 oA = odd ? make_uint2(recvA, pa[g + 1]) : make_uint2(pa[g], recvA);
                    const uint2 oB = odd ? make_uint2(recvB, pb[g + 1]) : make_uint2(pb[g], recvB);
                    if (okA) *reinterpret_cast<uint2 *>(na + col) = oA;
                    if (okB) *reinterpret_cast<uint2 *>(nb + col) = oB;
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
