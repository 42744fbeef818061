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
// Persistent kernel: one CTA per SM walks over work items (line, head, 128-query tile), q-tile fastest so that
// the CTAs working at the same time share K/V in L2.  Roles: warps 0-3 online softmax + output (thread = query row
// = TMEM lane), warp 4 TMA producer of Q and K(+E) (runs ahead across items), warp 5 TMEM allocator + MMA issuer,
// warp 6 TMA producer of V.  K and V have separate rings: a K slot is free as soon as its S MMA has run, a V slot
// when its P V has run, so the K loads run several tiles ahead of the softmax.
//   * S = Qe Ke^T is double-buffered in TMEM; the MMA of the next tile (also across items) is issued before the
//     P V of the current one, so the tensor pipe works under the softmax;
//   * the softmax reads its whole 128-key row into registers with one wait, rescales the running output only when
//     the row maximum grew by more than 2^8 (lazy rescale; exact because the final division uses the same reference
//     point), and writes P as packed bf16 back into TMEM over the S columns it has just consumed;
//   * P V takes its A operand straight from TMEM (tcgen05.mma with A in tensor memory): P never touches shared memory;
//   * the E table stays resident in shared memory when T <= ATT_ERES_TILES*128, else it streams with K/V.
#pragma once

#include "tc_utils.cuh"

namespace cocr {

struct AttnArgs {
    int B, T, nh, dh, d;
    int nchunk;                 // 64-column chunks of the score contraction = (48 + d) / 64  (3 for d = 144)
    int nqt;                    // query tiles per line
    int n_items;                // B * nh * nqt
    int e_resident;             // 1: E table resident in shared memory
    int k_stages, v_stages;     // ring slots of the K(+E) tiles and of the V tiles
    const int32_t *lens;        // ragged mode: valid frames per line, else null
    __nv_bfloat16 *ctx;         // (B*T, d) head-major context
    float scale_log2;           // log2(e) / sqrt(dh)
    long long *trace;           // debug: (tag, clock) pairs of CTA 0 (COCR_ATTN_TRACE), else null
};

constexpr int ATT_TRACE_EVENTS = 2048;     // per role (0 softmax warp 0, 1 MMA thread, 2 K loader)
#define ATT_TRACE(role, tag)                                                                      \
    do {                                                                                          \
        if (p.trace && blockIdx.x == 0 && tr_n < ATT_TRACE_EVENTS) {                              \
            p.trace[((role) * ATT_TRACE_EVENTS + tr_n) * 2] = (tag);                              \
            p.trace[((role) * ATT_TRACE_EVENTS + tr_n) * 2 + 1] = clock64();                      \
            ++tr_n;                                                                               \
        }                                                                                         \
    } while (0)

constexpr int ATT_THREADS = 224;
constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr uint32_t ATT_CHUNK = 128 * 128;      // bytes of a 128-row x 64-col bf16 SWIZZLE_128B tile
constexpr int ATT_MAXCHUNK = 3;
constexpr int ATT_MAXSTAGES = 4;
constexpr int ATT_ERES_TILES = 3;              // E resident for T <= 384
constexpr int ATT_DV = 64;                     // padded head dim of the V operand / O accumulator
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_S_COL0 = 0, ATT_S_COL1 = 128, ATT_O_COL = 256;
constexpr float ATT_LAZY_LOG2 = 8.0f;          // rescale only when the maximum grew by more than 2^8

__host__ __device__ inline uint32_t attn_kstage_bytes(int nchunk, int e_resident) {
    return (uint32_t)(e_resident ? 1 : nchunk) * ATT_CHUNK;              // K chunk 0 [+ E chunks]
}
__host__ __device__ inline uint32_t attn_eres_bytes(int nchunk, int T) {
    return (uint32_t)((T + ATT_BK - 1) / ATT_BK) * (uint32_t)(nchunk - 1) * ATT_CHUNK;
}
__host__ __device__ inline uint32_t attn_smem_bytes(int nchunk, int e_resident, int k_stages, int v_stages, int T) {
    return (uint32_t)nchunk * ATT_CHUNK + (e_resident ? attn_eres_bytes(nchunk, T) : 0u) +
           (uint32_t)k_stages * attn_kstage_bytes(nchunk, e_resident) + (uint32_t)v_stages * ATT_CHUNK + 256 + 1024;
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
    const uint32_t kstage_bytes = attn_kstage_bytes(p.nchunk, p.e_resident);
    const int ne = p.nchunk - 1;                                         // E chunks per key tile
    uint8_t *sQ = smem;                                                  // nchunk tiles
    uint8_t *sE = sQ + (uint32_t)p.nchunk * ATT_CHUNK;                   // resident E: [key tile][chunk]
    uint8_t *sK = sE + (p.e_resident ? attn_eres_bytes(p.nchunk, p.T) : 0u);
    uint8_t *sV = sK + (uint32_t)p.k_stages * kstage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sV + (uint32_t)p.v_stages * ATT_CHUNK);
    uint64_t *qfull = bars;             // [1]
    uint64_t *qempty = bars + 1;        // [1]
    uint64_t *efull = bars + 2;         // [1]
    uint64_t *sfull = bars + 3;         // [2]
    uint64_t *pfull = bars + 5;         // [2]
    uint64_t *pvdone = bars + 7;        // [2]
    uint64_t *kfull = bars + 9;         // [k_stages]
    uint64_t *kempty = kfull + ATT_MAXSTAGES;
    uint64_t *vfull = kempty + ATT_MAXSTAGES;   // [v_stages]
    uint64_t *vempty = vfull + ATT_MAXSTAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(vempty + ATT_MAXSTAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qe_cols = 48 + p.d;                                         // columns of one head's Qe block

    auto decode = [&](int item, int &b, int &hd, int &qt) {
        qt = item % p.nqt;
        const int bh = item / p.nqt;
        hd = bh % p.nh;
        b = bh / p.nh;
    };
    auto key_tiles = [&](int b) {
        const int Tk = p.lens ? min(max(p.lens[b], 0), p.T) : p.T;       // keys beyond are masked
        return (Tk + ATT_BK - 1) / ATT_BK;
    };

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmE);
        mbar_init(qfull, 1); mbar_init(qempty, 1); mbar_init(efull, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&sfull[s], 1); mbar_init(&pfull[s], 4); mbar_init(&pvdone[s], 1); }
        for (int s = 0; s < p.k_stages; ++s) { mbar_init(&kfull[s], 1); mbar_init(&kempty[s], 1); }
        for (int s = 0; s < p.v_stages; ++s) { mbar_init(&vfull[s], 1); mbar_init(&vempty[s], 1); }
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ================= TMA producer =================
        if (lane == 0) {
            if (p.e_resident) {
                const int nt = (p.T + ATT_BK - 1) / ATT_BK;
                mbar_expect_tx(efull, attn_eres_bytes(p.nchunk, p.T));
                for (int j = 0; j < nt; ++j)
                    for (int c = 0; c < ne; ++c)
                        tma_load_2d(sE + (uint32_t)(j * ne + c) * ATT_CHUNK, &tmE, efull, c * 64, j * ATT_BK);
            }
            uint32_t g = 0, nq = 0;                                       // global key-tile counter, Q loads issued
            int tr_n = 0;
            auto load_kv = [&](int b, int hd, int j) {
                const int st = (int)(g % (uint32_t)p.k_stages);
                const uint32_t ph = (g / (uint32_t)p.k_stages) & 1u;
                ATT_TRACE(2, 100 + j);
                mbar_wait(&kempty[st], ph ^ 1u);
                ATT_TRACE(2, 110 + j);
                mbar_expect_tx(&kfull[st], kstage_bytes);
                uint8_t *dst = sK + (uint32_t)st * kstage_bytes;
                tma_load_2d(dst, &tmK, &kfull[st], hd * 64, b * p.T + j * ATT_BK);
                if (!p.e_resident)
                    for (int c = 0; c < ne; ++c) tma_load_2d(dst + (uint32_t)(c + 1) * ATT_CHUNK, &tmE, &kfull[st], c * 64, j * ATT_BK);
                ++g;
            };
            // the tiles of the NEXT item are requested into L2 while the current one is processed: the activations do
            // not fit L2 between kernels, so a tile fetched only when its ring slot frees would pay HBM latency
            auto prefetch_item = [&](int item) {
                if (item >= p.n_items) return;
                int b, hd, qt;
                decode(item, b, hd, qt);
                const int ntiles = key_tiles(b);
                if (ntiles == 0) return;
                for (int c = 0; c < p.nchunk; ++c) tma_prefetch_2d(&tmQ, hd * qe_cols + c * 64, b * p.T + qt * ATT_BQ);
                for (int j = 0; j < ntiles; ++j) tma_prefetch_2d(&tmK, hd * 64, b * p.T + j * ATT_BK);
            };
            prefetch_item(blockIdx.x);
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int b, hd, qt;
                decode(item, b, hd, qt);
                const int ntiles = key_tiles(b);
                prefetch_item(item + gridDim.x);
                if (ntiles == 0) continue;
                load_kv(b, hd, 0);                                        // does not depend on the Q buffer
                ATT_TRACE(2, 120);
                mbar_wait(qempty, (nq & 1u) ^ 1u);                        // last S MMA of the previous item has read Q
                ATT_TRACE(2, 121);
                mbar_expect_tx(qfull, (uint32_t)p.nchunk * ATT_CHUNK);
                for (int c = 0; c < p.nchunk; ++c)
                    tma_load_2d(sQ + (uint32_t)c * ATT_CHUNK, &tmQ, qfull, hd * qe_cols + c * 64, b * p.T + qt * ATT_BQ);
                ++nq;
                for (int j = 1; j < ntiles; ++j) load_kv(b, hd, j);
            }
        }
    } else if (warp == 6) {
        // ================= TMA producer of V =================
        if (lane == 0) {
            uint32_t g = 0;
            auto prefetch_item = [&](int item) {
                if (item >= p.n_items) return;
                int b, hd, qt;
                decode(item, b, hd, qt);
                const int ntiles = key_tiles(b);
                for (int j = 0; j < ntiles; ++j) tma_prefetch_2d(&tmV, hd * 64, b * p.T + j * ATT_BK);
            };
            prefetch_item(blockIdx.x);
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int b, hd, qt;
                decode(item, b, hd, qt);
                const int ntiles = key_tiles(b);
                prefetch_item(item + gridDim.x);
                for (int j = 0; j < ntiles; ++j, ++g) {
                    const int st = (int)(g % (uint32_t)p.v_stages);
                    mbar_wait(&vempty[st], ((g / (uint32_t)p.v_stages) & 1u) ^ 1u);
                    mbar_expect_tx(&vfull[st], ATT_CHUNK);
                    tma_load_2d(sV + (uint32_t)st * ATT_CHUNK, &tmV, &vfull[st], hd * 64, b * p.T + j * ATT_BK);
                }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(ATT_BQ, ATT_BK);                 // S = Qe Ke^T  (both K-major)
            const uint32_t idesc_o = make_idesc_bf16(ATT_BQ, ATT_DV, 0, 1);           // O += P V     (P in TMEM, V MN-major)
            if (p.e_resident) mbar_wait(efull, 0);
            // flat walk over (item, key tile) with one tile of look-ahead
            struct Cur { int item, j, ntiles; bool valid; };
            auto first_from = [&](int item) {
                Cur c{item, 0, 0, false};
                for (; c.item < p.n_items; c.item += gridDim.x) {
                    int b, hd, qt;
                    decode(c.item, b, hd, qt);
                    c.ntiles = key_tiles(b);
                    if (c.ntiles > 0) { c.valid = true; break; }
                }
                return c;
            };
            auto next_of = [&](const Cur &c) {
                if (c.j + 1 < c.ntiles) return Cur{c.item, c.j + 1, c.ntiles, true};
                return first_from(c.item + gridDim.x);
            };
            uint32_t nq = 0;                                                          // Q buffers consumed
            int tr_n = 0;
            auto issue_s = [&](const Cur &c, uint32_t g) {
                ATT_TRACE(1, 200 + c.j);
                if (c.j == 0) { mbar_wait(qfull, nq & 1u); ++nq; }
                ATT_TRACE(1, 210 + c.j);
                const int st = (int)(g % (uint32_t)p.k_stages);
                mbar_wait(&kfull[st], (g / (uint32_t)p.k_stages) & 1u);
                ATT_TRACE(1, 220 + c.j);
                tc_fence_after();
                const uint32_t k_base = smem_u32(sK + (uint32_t)st * kstage_bytes);
                const uint32_t q_base = smem_u32(sQ);
                const uint32_t d_s = tmem_base + ((g & 1u) ? ATT_S_COL1 : ATT_S_COL0);
                for (int cc = 0; cc < p.nchunk; ++cc) {
                    uint32_t b_base;
                    if (cc == 0) b_base = k_base;
                    else if (p.e_resident) b_base = smem_u32(sE + (uint32_t)(c.j * ne + cc - 1) * ATT_CHUNK);
                    else b_base = k_base + (uint32_t)cc * ATT_CHUNK;
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = make_smem_desc_sw128(q_base + (uint32_t)cc * ATT_CHUNK + ks * 32, 1024);
                        const uint64_t bd = make_smem_desc_sw128(b_base + ks * 32, 1024);
                        umma_bf16_ss(d_s, ad, bd, idesc_s, (cc | ks) ? 1u : 0u);
                    }
                }
                umma_commit(&sfull[g & 1u]);
                umma_commit(&kempty[st]);                                             // K(+E) slot may be refilled
                if (c.j == c.ntiles - 1) umma_commit(qempty);                         // Q buffer free once this MMA has run
            };
            uint32_t g = 0;
            Cur cur = first_from(blockIdx.x);
            if (cur.valid) issue_s(cur, 0);
            while (cur.valid) {
                const Cur nxt = next_of(cur);
                if (nxt.valid) issue_s(nxt, g + 1);
                // O (+)= P V for the current tile
                const int st = (int)(g % (uint32_t)p.v_stages);
                ATT_TRACE(1, 230 + cur.j);
                mbar_wait(&vfull[st], (g / (uint32_t)p.v_stages) & 1u);
                ATT_TRACE(1, 240 + cur.j);
                mbar_wait(&pfull[g & 1u], (g >> 1) & 1u);
                ATT_TRACE(1, 250 + cur.j);
                tc_fence_after();
                const uint32_t v_base = smem_u32(sV + (uint32_t)st * ATT_CHUNK);
                const uint32_t p_tmem = tmem_base + ((g & 1u) ? ATT_S_COL1 : ATT_S_COL0);
                for (int ks = 0; ks < ATT_BK / 16; ++ks) {
                    const uint64_t bd = make_smem_desc_sw128(v_base + ks * 2048, 1024);
                    umma_bf16_ts(tmem_base + ATT_O_COL, p_tmem + (uint32_t)ks * 8u, bd, idesc_o, (cur.j | ks) ? 1u : 0u);
                }
                umma_commit(&vempty[st]);          // V slot may be refilled
                umma_commit(&pvdone[g & 1u]);      // O up to date
                cur = nxt;
                ++g;
            }
        }
    } else {
        // ================= softmax / output (thread = query row) =================
        const int r = warp * 32 + lane;                         // row in the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        uint32_t g = 0;
        int tr_n = (threadIdx.x == 0) ? 0 : ATT_TRACE_EVENTS;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int b, hd, qt;
            decode(item, b, hd, qt);
            const int Tk = p.lens ? min(max(p.lens[b], 0), p.T) : p.T;
            const int ntiles = (Tk + ATT_BK - 1) / ATT_BK;
            const int i = qt * ATT_BQ + r;                          // frame index of this query inside its line
            float m_run = -INFINITY, l_run = 0.0f;
            for (int j = 0; j < ntiles; ++j, ++g) {
                const uint32_t sb = g & 1u;
                ATT_TRACE(0, 300 + j);
                mbar_wait(&sfull[sb], (g >> 1) & 1u);
                ATT_TRACE(0, 310 + j);
                tc_fence_after();
                const uint32_t s_addr = tmem_base + (sb ? ATT_S_COL1 : ATT_S_COL0) + lane_base;
                uint32_t v0[32], v1[32], v2[32], v3[32];
                tmem_ld_32x32b_x32(s_addr, v0);
                tmem_ld_32x32b_x32(s_addr + 32, v1);
                tmem_ld_32x32b_x32(s_addr + 64, v2);
                tmem_ld_32x32b_x32(s_addr + 96, v3);
                tmem_ld_wait_dep(v0); tmem_ld_wait_dep(v1); tmem_ld_wait_dep(v2); tmem_ld_wait_dep(v3);
                ATT_TRACE(0, 320 + j);
                const int kv0 = j * ATT_BK;
                const int nvalid = Tk - kv0;                         // >= 1; keys >= nvalid of this tile are masked
                if (nvalid < ATT_BK) {
                    const uint32_t ninf = __float_as_uint(-INFINITY);
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (e >= nvalid) v0[e] = ninf;
                        if (32 + e >= nvalid) v1[e] = ninf;
                        if (64 + e >= nvalid) v2[e] = ninf;
                        if (96 + e >= nvalid) v3[e] = ninf;
                    }
                }
                float t0 = -INFINITY, t1 = -INFINITY, t2 = -INFINITY, t3 = -INFINITY;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    t0 = fmaxf(t0, __uint_as_float(v0[e])); t1 = fmaxf(t1, __uint_as_float(v1[e]));
                    t2 = fmaxf(t2, __uint_as_float(v2[e])); t3 = fmaxf(t3, __uint_as_float(v3[e]));
                }
                const float m_new = fmaxf(m_run, fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)));
                if (j == 0) {
                    m_run = m_new;
                } else {
                    const bool need = (m_new - m_run) * p.scale_log2 > ATT_LAZY_LOG2;
                    if (__any_sync(0xffffffffu, need)) {
                        // rescale the running output of every row of this warp to its new maximum (TMEM round trip)
                        const float corr = ex2_fast((m_run - m_new) * p.scale_log2);
                        m_run = m_new;
                        l_run *= corr;
                        mbar_wait(&pvdone[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < 3; ++c) {                  // 48 columns cover dh <= 48
                            uint32_t o[16];
                            tmem_ld_32x32b_x16(tmem_base + ATT_O_COL + lane_base + c * 16, o);
                            tmem_ld_wait_dep(o);
#pragma unroll
                            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
                            tmem_st_32x32b_x16(tmem_base + ATT_O_COL + lane_base + c * 16, o);
                        }
                    }
                }
                // probabilities -> packed bf16 in TMEM over the first 64 columns of this S buffer
                const float mb = m_run * p.scale_log2;
                float ps0 = 0.f, ps1 = 0.f;
                auto emit = [&](uint32_t (&v)[32], int cidx) {
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float p0 = ex2_fast(fmaf(__uint_as_float(v[2 * e]), p.scale_log2, -mb));
                        const float p1 = ex2_fast(fmaf(__uint_as_float(v[2 * e + 1]), p.scale_log2, -mb));
                        ps0 += p0; ps1 += p1;
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                        pk[e] = *reinterpret_cast<uint32_t *>(&h2);
                    }
                    tmem_st_32x32b_x16(s_addr + (uint32_t)cidx * 16u, pk);
                };
                emit(v0, 0); emit(v1, 1); emit(v2, 2); emit(v3, 3);
                l_run += ps0 + ps1;
                ATT_TRACE(0, 330 + j);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pfull[sb]);
                ATT_TRACE(0, 340 + j);
            }
            // output: O / l -> ctx (head-major columns hd*dh .. +dh)
            const float inv = (l_run > 0.0f) ? 1.0f / l_run : 0.0f;
            uint32_t o0[32], o1[16];
            if (ntiles > 0) {
                ATT_TRACE(0, 350);
                mbar_wait(&pvdone[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
                ATT_TRACE(0, 351);
                tc_fence_after();
                tmem_ld_32x32b_x32(tmem_base + ATT_O_COL + lane_base, o0);
                tmem_ld_32x32b_x16(tmem_base + ATT_O_COL + lane_base + 32, o1);
                tmem_ld_wait_dep(o0); tmem_ld_wait_dep(o1);
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) o0[e] = 0u;
#pragma unroll
                for (int e = 0; e < 16; ++e) o1[e] = 0u;
            }
            if (i < p.T) {
                __nv_bfloat16 *dst = p.ctx + ((int64_t)b * p.T + i) * p.d + hd * p.dh;
#pragma unroll
                for (int e = 0; e < 48; e += 2) {
                    if (e + 1 < p.dh) {
                        const float a = __uint_as_float(e < 32 ? o0[e & 31] : o1[e & 15]) * inv;
                        const float c = __uint_as_float(e + 1 < 32 ? o0[(e + 1) & 31] : o1[(e + 1) & 15]) * inv;
                        *reinterpret_cast<__nv_bfloat162 *>(dst + e) = __floats2bfloat162_rn(a, c);
                    } else if (e < p.dh) {
                        dst[e] = __float2bfloat16_rn(__uint_as_float(e < 32 ? o0[e & 31] : o1[e & 15]) * inv);
                    }
                }
            }
            ATT_TRACE(0, 360);
            // the O reads above are complete (wait::ld); the next item's first P V is ordered behind them by pfull
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
