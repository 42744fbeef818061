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
// the CTAs working at the same time share K/V in L2.
//   warps 0-7   online softmax: thread = (query row = TMEM lane, half of the 128 keys of a tile); the two warps of a
//               lane quarter exchange their row maxima through shared memory.  They do nothing else: the softmax is the
//               critical path (MUFU bound), everything around it runs on other warps
//   warps 8-11  auxiliary, one per lane quarter: copy the staged Qe tile of the NEXT item into TMEM and, when the last
//               P V of an item has run, scale O by 1/l and write the context rows (the softmax warps are already on the
//               next item by then)
//   warp 12     TMA producer of Q and K(+E), runs ahead across items and requests the next item into L2
//   warp 13     TMEM allocator + MMA issuer
//   warp 14     TMA producer of V
// Tensor memory (512 columns):  S0 | S1 (2 x 128, fp32 scores, double-buffered)  O (64)  Q0 | Q1 (2 x 96, bf16 pairs).
//   * Both MMAs take their A operand from TENSOR MEMORY: the softmax warps copy each Qe tile from its TMA staging
//     buffer into TMEM once per item (one item ahead), and write P as packed bf16 over the S columns they have just
//     consumed.  An SS-mode 128x128x16 MMA would read 8 KB of shared memory per 64 tensor cycles - exactly the
//     128 B/clk shared-memory limit - and measured 110 cycles instead of 64; with A in TMEM only K/E/V stream from smem.
//   * S of the next tile (also across items) is issued while the softmax of the current one runs; the P V of the
//     current tile is slipped between the k-chunks of that S as soon as P is complete (it unblocks the output).
//   * lazy rescale: the running output is rescaled only when a row maximum grew by more than 2^8 (exact, because
//     the final division uses the same reference point).
//   * K and V have separate rings (a K slot frees when its S MMA has run, a V slot when its P V has run); the E table
//     stays resident in shared memory when T <= ATT_ERES_TILES*128, else it streams with K.
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
    int kv_resident;            // 1: a CTA takes whole (line, head) groups; the group's K and V tiles are loaded once into
                                //    slots 0..tiles-1 and reused by all of its query tiles (needs e_resident, tiles <= stages)
    const int32_t *lens;        // ragged mode: valid frames per line, else null
    __nv_bfloat16 *ctx;         // (B*T, d) head-major context
    float scale_log2;           // log2(e) / sqrt(dh)
    int dv_mma;                 // N of the P V MMA: ATT_DV_MMA (48) covers dh <= 40; ATT_DV (64) = the whole V tile (COCR_ATT_DV64)
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

constexpr int ATT_SM_WARPS = 8;                   // softmax warps
constexpr int ATT_AUX_WARPS = 4;                  // Q staging + output warps (one per TMEM lane quarter)
constexpr int ATT_W_LOADK = ATT_SM_WARPS + ATT_AUX_WARPS, ATT_W_MMA = ATT_W_LOADK + 1, ATT_W_LOADV = ATT_W_LOADK + 2;
constexpr int ATT_THREADS = 32 * (ATT_SM_WARPS + ATT_AUX_WARPS + 3);
constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr uint32_t ATT_CHUNK = 128 * 128;      // bytes of a 128-row x 64-col bf16 SWIZZLE_128B tile
constexpr int ATT_MAXCHUNK = 3;
constexpr int ATT_MAXSTAGES = 4;
constexpr uint32_t ATT_QSLOTS = 2;             // Qe staging: 64-column chunks pass through two slots on their way to TMEM
constexpr int ATT_ERES_TILES = 3;              // E resident for T <= 384
constexpr int ATT_DV = 64;                     // padded head dim of the V tile in shared memory
constexpr int ATT_DV_MMA = 48;                 // N of the P V MMA / columns of O actually computed (dh <= 40 on this path)
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_S_COL0 = 0, ATT_S_COL1 = 128, ATT_O_COL = 256, ATT_Q_COL0 = 320, ATT_Q_COL1 = 416;
constexpr int ATT_POLY_EVERY = 64;
constexpr float ATT_LAZY_LOG2 = 8.0f;          // rescale only when the maximum grew by more than 2^8
constexpr uint32_t ATT_XCH_BYTES = 2 * 2 * 128 * 2 + 2 * 128 * 4;   // row maxima [parity][half][row] bf16 + row sums [half][row] fp32
constexpr uint32_t ATT_BAR_BYTES = 288;          // 32 mbarriers + the TMEM base slot

__host__ __device__ inline uint32_t attn_kstage_bytes(int nchunk, int e_resident) {
    return (uint32_t)(e_resident ? 1 : nchunk) * ATT_CHUNK;              // K chunk 0 [+ E chunks]
}
__host__ __device__ inline uint32_t attn_eres_bytes(int nchunk, int T) {
    return (uint32_t)((T + ATT_BK - 1) / ATT_BK) * (uint32_t)(nchunk - 1) * ATT_CHUNK;
}
__host__ __device__ inline uint32_t attn_smem_bytes(int nchunk, int e_resident, int k_stages, int v_stages, int T) {
    return ATT_QSLOTS * ATT_CHUNK + (e_resident ? attn_eres_bytes(nchunk, T) : 0u) +
           (uint32_t)k_stages * attn_kstage_bytes(nchunk, e_resident) + (uint32_t)v_stages * ATT_CHUNK + ATT_XCH_BYTES + ATT_BAR_BYTES;
}

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed fp32 pairs (FFMA2 / FADD2): one issue slot for two elements
__device__ __forceinline__ float2 attn_fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
        "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float2 attn_add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}
// 2^x for a pair on the FMA pipe (the MUFU unit is the softmax bottleneck: 8 cycles per warp instruction): round-to-nearest
// split x = n + f, f in [-0.5, 0.5], by the 1.5*2^23 trick, cubic minimax polynomial for 2^f (relative error 1.0e-4, far
// below the bf16 rounding of P), exponent added as an integer.  x <= 0 up to the lazy-rescale slack; -inf (masked) -> 0.
__device__ __forceinline__ float2 attn_ex2_poly2(float2 x) {
    const float magic = 12582912.0f;                                    // 1.5 * 2^23
    x.x = fmaxf(x.x, -125.0f); x.y = fmaxf(x.y, -125.0f);
    const float2 t = attn_add2(x, make_float2(magic, magic));           // integer part in the low mantissa bits
    const float2 n = attn_add2(t, make_float2(-magic, -magic));
    const float2 f = attn_add2(x, make_float2(-n.x, -n.y));
    float2 q = attn_fma2(f, make_float2(0.05500843f, 0.05500843f), make_float2(0.24220999f, 0.24220999f));
    q = attn_fma2(q, f, make_float2(0.69328287f, 0.69328287f));
    q = attn_fma2(q, f, make_float2(1.0f, 1.0f));
    q.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
    q.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
    return q;
}
__device__ __forceinline__ float attn_max3(float a, float b, float c) {       // FMNMX3: one issue slot for two comparisons
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ uint32_t attn_pack_bf16x2(float a, float b) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h2);
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
k_attn_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE, const AttnArgs p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];          // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t kstage_bytes = attn_kstage_bytes(p.nchunk, p.e_resident);
    const int ne = p.nchunk - 1;                                         // E chunks per key tile
    uint8_t *sQ = smem;                                                  // Q staging: ATT_QSLOTS chunk slots
    uint8_t *sE = sQ + ATT_QSLOTS * ATT_CHUNK;                           // resident E: [key tile][chunk]
    uint8_t *sK = sE + (p.e_resident ? attn_eres_bytes(p.nchunk, p.T) : 0u);
    uint8_t *sV = sK + (uint32_t)p.k_stages * kstage_bytes;
    __nv_bfloat16 *xmax = reinterpret_cast<__nv_bfloat16 *>(sV + (uint32_t)p.v_stages * ATT_CHUNK);   // [2][2][128] row maxima
    float *xl = reinterpret_cast<float *>(xmax + 2 * 2 * 128);                                           // [2][128] row sums
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(xmax) + ATT_XCH_BYTES);
    uint64_t *qcfull = bars;            // [2]  Q chunk slot filled by TMA
    uint64_t *qcempty = bars + 2;       // [2]  Q chunk slot copied out (aux warps)
    uint64_t *efull = bars + 4;         // [1]
    uint64_t *qtfull = bars + 5;        // [2]  Q copy in TMEM buffer complete (aux warps)
    uint64_t *sfull = bars + 7;         // [2]
    uint64_t *pfull = bars + 9;         // [2]  (8 warps)
    uint64_t *pvdone = bars + 11;       // [2]
    uint64_t *kfull = bars + 13;        // [k_stages]
    uint64_t *kempty = kfull + ATT_MAXSTAGES;
    uint64_t *vfull = kempty + ATT_MAXSTAGES;   // [v_stages]
    uint64_t *vempty = vfull + ATT_MAXSTAGES;
    uint64_t *lfull = vempty + ATT_MAXSTAGES;    // [1]  row sums of an item written (softmax warps)
    uint64_t *lempty = lfull + 1;                // [1]  row sums read (aux warps)
    uint64_t *oempty = lfull + 2;                // [1]  O read out (aux warps): the next item's first P V may overwrite it
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(lfull + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qe_cols = 48 + p.d;                                         // columns of one head's Qe block

    auto valid_keys = [&](int b) { return p.lens ? min(max(p.lens[b], 0), p.T) : p.T; };   // keys beyond are masked
    const int full_tiles = (p.T + ATT_BK - 1) / ATT_BK;
    // The CTA's work sequence.  Streaming mode deals single items (line, head, query tile) round-robin, q-tile fastest, so
    // that concurrent CTAs share K/V in L2; kv_resident mode deals whole (line, head) groups and walks the group's query
    // tiles.  A position is kept in mixed radix (b, hd, qt) and advanced by the CTA stride without any division (a runtime
    // integer division costs ~100 cycles, and every role walks the sequence on its critical path).
    struct Pos { int b, hd, qt; };
    int dqt, dhd, db;                                                    // stride between two units of this CTA
    Pos pos0;
    if (p.kv_resident) {
        dqt = 0; dhd = (int)gridDim.x % p.nh; db = (int)gridDim.x / p.nh;
        pos0.qt = 0; pos0.hd = (int)blockIdx.x % p.nh; pos0.b = (int)blockIdx.x / p.nh;
    } else {
        const int gq = (int)gridDim.x / p.nqt, bq = (int)blockIdx.x / p.nqt;
        dqt = (int)gridDim.x % p.nqt; dhd = gq % p.nh; db = gq / p.nh;
        pos0.qt = (int)blockIdx.x % p.nqt; pos0.hd = bq % p.nh; pos0.b = bq / p.nh;
    }
    auto seq_valid = [&](const Pos &s) { return s.b < p.B; };
    auto seq_next = [&](Pos s) {
        if (p.kv_resident) {
            if (++s.qt < p.nqt) return s;
            s.qt = 0;
        } else {
            s.qt += dqt;
            if (s.qt >= p.nqt) { s.qt -= p.nqt; ++s.hd; }
        }
        s.hd += dhd;
        if (s.hd >= p.nh) { s.hd -= p.nh; ++s.b; }
        s.b += db;
        return s;
    };
    auto key_tiles = [&](const Pos &s) { return p.lens ? (valid_keys(s.b) + ATT_BK - 1) / ATT_BK : full_tiles; };
    // first position from s on whose item has at least one key tile (or the end)
    auto next_active = [&](Pos s) {
        if (p.lens)
            while (seq_valid(s) && key_tiles(s) == 0) s = seq_next(s);
        return s;
    };

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0u) __trap();
        prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmE);
        for (int s = 0; s < (int)ATT_QSLOTS; ++s) { mbar_init(&qcfull[s], 1); mbar_init(&qcempty[s], ATT_AUX_WARPS); }
        mbar_init(efull, 1);
        mbar_init(lfull, ATT_SM_WARPS); mbar_init(lempty, ATT_AUX_WARPS); mbar_init(oempty, ATT_AUX_WARPS);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&qtfull[s], ATT_AUX_WARPS); mbar_init(&sfull[s], 1); mbar_init(&pfull[s], ATT_SM_WARPS); mbar_init(&pvdone[s], 1);
        }
        for (int s = 0; s < p.k_stages; ++s) { mbar_init(&kfull[s], 1); mbar_init(&kempty[s], 1); }
        for (int s = 0; s < p.v_stages; ++s) { mbar_init(&vfull[s], 1); mbar_init(&vempty[s], 1); }
        mbar_fence_init();
    }
    if (warp == ATT_W_MMA) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == ATT_W_LOADK) {
        // ================= TMA producer of Q and K(+E) =================
        if (lane == 0) {
            if (p.e_resident) {
                const int nt = (p.T + ATT_BK - 1) / ATT_BK;
                mbar_expect_tx(efull, attn_eres_bytes(p.nchunk, p.T));
                for (int j = 0; j < nt; ++j)
                    for (int c = 0; c < ne; ++c)
                        tma_load_2d(sE + (uint32_t)(j * ne + c) * ATT_CHUNK, &tmE, efull, c * 64, j * ATT_BK);
            }
            uint32_t g = 0, qc = 0, kpar = 0;                             // key tiles loaded, Q chunks loaded, slot parities
            int tr_n = 0;
            auto load_k = [&](int b, int hd, int j) {
                int st;
                uint32_t ph;
                if (p.kv_resident) { st = j; ph = (kpar >> j) & 1u; kpar ^= 1u << j; }
                else { st = (int)(g % (uint32_t)p.k_stages); ph = (g / (uint32_t)p.k_stages) & 1u; }
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
            // not fit L2 between kernels, so a tile fetched only when its slot frees would pay HBM latency
            auto prefetch_item = [&](const Pos &s) {
                if (!seq_valid(s)) return;
                const int ntiles = key_tiles(s);
                if (ntiles == 0) return;
                for (int c = 0; c < p.nchunk; ++c) tma_prefetch_2d(&tmQ, s.hd * qe_cols + c * 64, s.b * p.T + s.qt * ATT_BQ);
                if (!p.kv_resident || s.qt == 0)
                    for (int j = 0; j < ntiles; ++j) tma_prefetch_2d(&tmK, s.hd * 64, s.b * p.T + j * ATT_BK);
            };
            prefetch_item(pos0);
            for (Pos s = pos0; seq_valid(s); s = seq_next(s)) {
                const int b = s.b, hd = s.hd, qt = s.qt;
                const int ntiles = key_tiles(s);
                prefetch_item(seq_next(s));
                if (ntiles == 0) continue;
                const bool need_k = !p.kv_resident || qt == 0;            // resident mode: the group's K tiles are loaded once
                if (need_k) load_k(b, hd, 0);                             // does not depend on the Q staging slots
                for (int c = 0; c < p.nchunk; ++c, ++qc) {
                    const uint32_t sl = qc % ATT_QSLOTS;
                    mbar_wait(&qcempty[sl], ((qc / ATT_QSLOTS) & 1u) ^ 1u);   // the slot's previous chunk has been copied into TMEM
                    ATT_TRACE(2, 121 + c);
                    mbar_expect_tx(&qcfull[sl], ATT_CHUNK);
                    tma_load_2d(sQ + sl * ATT_CHUNK, &tmQ, &qcfull[sl], hd * qe_cols + c * 64, b * p.T + qt * ATT_BQ);
                }
                if (need_k)
                    for (int j = 1; j < ntiles; ++j) load_k(b, hd, j);
            }
        }
    } else if (warp == ATT_W_LOADV) {
        // ================= TMA producer of V =================
        if (lane == 0) {
            uint32_t g = 0, vpar = 0;
            auto prefetch_item = [&](const Pos &s) {
                if (!seq_valid(s) || (p.kv_resident && s.qt != 0)) return;
                const int ntiles = key_tiles(s);
                for (int j = 0; j < ntiles; ++j) tma_prefetch_2d(&tmV, s.hd * 64, s.b * p.T + j * ATT_BK);
            };
            prefetch_item(pos0);
            for (Pos s = pos0; seq_valid(s); s = seq_next(s)) {
                const int b = s.b, hd = s.hd;
                const int ntiles = key_tiles(s);
                prefetch_item(seq_next(s));
                if (p.kv_resident && s.qt != 0) continue;
                for (int j = 0; j < ntiles; ++j, ++g) {
                    int st;
                    uint32_t ph;
                    if (p.kv_resident) { st = j; ph = (vpar >> j) & 1u; vpar ^= 1u << j; }
                    else { st = (int)(g % (uint32_t)p.v_stages); ph = (g / (uint32_t)p.v_stages) & 1u; }
                    mbar_wait(&vempty[st], ph ^ 1u);
                    mbar_expect_tx(&vfull[st], ATT_CHUNK);
                    tma_load_2d(sV + (uint32_t)st * ATT_CHUNK, &tmV, &vfull[st], hd * 64, b * p.T + j * ATT_BK);
                }
            }
        }
    } else if (warp == ATT_W_MMA) {
        // ================= MMA issuer =================
        // All 32 lanes walk the schedule (identical values in every lane, so they live in uniform registers) and one
        // elected lane issues the tcgen05 instructions: under a divergent `lane == 0` branch the compiler wraps every
        // UTCHMMA in an elect/branch loop that costs more than a 64-cycle MMA.
        {
            const uint32_t idesc_s = make_idesc_bf16(ATT_BQ, ATT_BK);                 // S = Qe Ke^T  (Q in TMEM, K/E K-major)
            const uint32_t idesc_o = make_idesc_bf16(ATT_BQ, p.dv_mma, 0, 1);   // O += P V  (P in TMEM, V MN-major)
            if (p.e_resident) mbar_wait(efull, 0);
            // flat walk over (item, key tile) with one tile of look-ahead
            // s = position in the CTA's sequence, a = index among the active items; first/last = first/last query tile
            // of a K/V-resident group (both true in streaming mode: every item loads and releases its own tiles)
            struct Cur { Pos s; int j, ntiles; uint32_t a; bool first, last, valid; };
            auto first_from = [&](const Pos &s, uint32_t a) {
                Cur c{next_active(s), 0, 0, a, true, true, false};
                if (seq_valid(c.s)) {
                    c.ntiles = key_tiles(c.s);
                    c.valid = true;
                    if (p.kv_resident) { c.first = c.s.qt == 0; c.last = c.s.qt == p.nqt - 1; }
                }
                return c;
            };
            auto next_of = [&](const Cur &c) {
                if (c.j + 1 < c.ntiles) { Cur n = c; ++n.j; return n; }
                return first_from(seq_next(c.s), c.a + 1);
            };
            int tr_n = (lane == 0) ? 0 : ATT_TRACE_EVENTS;
            uint32_t g = 0, kpar = 0, vpar = 0;                                       // resident mode: per-slot barrier parities
            Cur cur = first_from(pos0, 0);
            // P V of the current tile (global index g)
            auto issue_pv = [&]() {
                int st;
                if (p.kv_resident) {
                    st = cur.j;
                    if (cur.first) { mbar_wait(&vfull[st], (vpar >> st) & 1u); vpar ^= 1u << st; }
                } else {
                    st = (int)(g % (uint32_t)p.v_stages);
                    mbar_wait(&vfull[st], (g / (uint32_t)p.v_stages) & 1u);
                }
                mbar_wait(&pfull[g & 1u], (g >> 1) & 1u);
                if (cur.j == 0) mbar_wait(oempty, (cur.a & 1u) ^ 1u);                 // overwrites O: the aux warps must have read it
                ATT_TRACE(1, 250 + cur.j);
                tc_fence_after();
                const uint32_t v_base = smem_u32(sV + (uint32_t)st * ATT_CHUNK);
                const uint32_t p_tmem = tmem_base + ((g & 1u) ? ATT_S_COL1 : ATT_S_COL0);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < ATT_BK / 16; ++ks) {
                        const uint64_t bd = make_smem_desc_sw128(v_base + ks * 2048, 1024);
                        umma_bf16_ts(tmem_base + ATT_O_COL, p_tmem + (uint32_t)ks * 8u, bd, idesc_o, (cur.j | ks) ? 1u : 0u);
                    }
                    if (cur.last) umma_commit(&vempty[st]);   // V slot may be refilled
                    umma_commit(&pvdone[g & 1u]);             // O up to date
                }
                __syncwarp();
            };
            // S of tile c (global index gs); the pending P V of the current tile is slipped in as soon as it is ready
            auto issue_s = [&](const Cur &c, uint32_t gs) {
                ATT_TRACE(1, 200 + c.j);
                // plain blocking waits and one uninterrupted burst of MMAs: probing the P barrier between the k-chunks
                // (mbarrier.test_wait costs ~150 cycles a probe) stretched the issue of one S tile from ~100 to ~1500 cycles
                if (c.j == 0) mbar_wait(&qtfull[c.a & 1u], (c.a >> 1) & 1u);
                int st;
                if (p.kv_resident) {
                    st = c.j;
                    if (c.first) { mbar_wait(&kfull[st], (kpar >> st) & 1u); kpar ^= 1u << st; }
                } else {
                    st = (int)(gs % (uint32_t)p.k_stages);
                    mbar_wait(&kfull[st], (gs / (uint32_t)p.k_stages) & 1u);
                }
                ATT_TRACE(1, 220 + c.j);
                tc_fence_after();
                const uint32_t k_base = smem_u32(sK + (uint32_t)st * kstage_bytes);
                const uint32_t q_tmem = tmem_base + ((c.a & 1u) ? ATT_Q_COL1 : ATT_Q_COL0);
                const uint32_t d_s = tmem_base + ((gs & 1u) ? ATT_S_COL1 : ATT_S_COL0);
                const uint32_t e_base = p.e_resident ? smem_u32(sE + (uint32_t)(c.j * ne) * ATT_CHUNK) : k_base + ATT_CHUNK;
                if (elect_one()) {
                    for (int cc = 0; cc < p.nchunk; ++cc) {
                        const uint32_t b_base = (cc == 0) ? k_base : e_base + (uint32_t)(cc - 1) * ATT_CHUNK;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t bd = make_smem_desc_sw128(b_base + ks * 32, 1024);
                            umma_bf16_ts(d_s, q_tmem + (uint32_t)(cc * 32 + ks * 8), bd, idesc_s, (cc | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit(&sfull[gs & 1u]);
                    if (c.last) umma_commit(&kempty[st]);                             // K(+E) slot may be refilled
                }
                __syncwarp();
            };
            if (cur.valid) issue_s(cur, 0);
            while (cur.valid) {
                const Cur nxt = next_of(cur);
                if (nxt.valid) issue_s(nxt, g + 1);
                issue_pv();
                cur = nxt;
                ++g;
            }
        }
    } else if (warp >= ATT_SM_WARPS) {
        // ================= auxiliary warps: Q staging -> TMEM, output =================
        const int q = warp & 3;                                  // TMEM lane quarter (ATT_SM_WARPS is a multiple of 4)
        const int r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t g = 0, na = 0, qc = 0;                          // key tiles passed, active items seen, Q chunks copied
        // row r of the Qe tile of active item #a, chunk by chunk from the staging slots -> TMEM buffer a&1
        auto copy_q = [&](uint32_t a) {
            const uint32_t q_tmem = tmem_base + ((a & 1u) ? ATT_Q_COL1 : ATT_Q_COL0) + lane_base;
            for (int c = 0; c < p.nchunk; ++c, ++qc) {
                const uint32_t sl = qc % ATT_QSLOTS;
                mbar_wait(&qcfull[sl], (qc / ATT_QSLOTS) & 1u);
                const uint8_t *src = sQ + sl * ATT_CHUNK;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    uint32_t w[16];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                                    // 16-byte pieces of the row
                        const uint4 t = *reinterpret_cast<const uint4 *>(src + sw128_offset(r, u * 4 + k));
                        w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
                    }
                    tmem_st_32x32b_x16n(q_tmem + (uint32_t)(c * 32 + u * 16), w);
                }
                __syncwarp();                                                        // the stores consumed the shared-memory reads
                if (lane == 0) mbar_arrive(&qcempty[sl]);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&qtfull[a & 1u]);
        };
        if (seq_valid(next_active(pos0))) copy_q(0);
        for (Pos s = pos0; seq_valid(s); s = seq_next(s)) {
            const int b = s.b, hd = s.hd, qt = s.qt;
            const int ntiles = key_tiles(s);
            if (ntiles == 0) {
                // no keys: emit zeros like the SIMT path
                const int i = qt * ATT_BQ + r;
                if (i < p.T)
                    for (int c = 0; c < p.dh; ++c) p.ctx[((int64_t)b * p.T + i) * p.d + hd * p.dh + c] = __float2bfloat16_rn(0.0f);
                continue;
            }
            // one item ahead: the TMEM buffer (na+1)&1 was last read by item na-1, whose output (previous iteration) waited
            // for all of its MMAs
            if (seq_valid(next_active(seq_next(s)))) copy_q(na + 1);
            g += (uint32_t)ntiles;
            // ---- output: O / l -> ctx (head-major columns hd*dh .. +dh), fragment layout, two 16-row halves ----
            mbar_wait(lfull, na & 1u);
            const int rin = lane >> 2, qd = lane & 3;
            float inv[2][2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int rowA = q * 32 + hh * 16 + rin, rowB = rowA + 8;
                const float lA = xl[rowA] + xl[128 + rowA], lB = xl[rowB] + xl[128 + rowB];
                inv[hh][0] = (lA > 0.0f) ? 1.0f / lA : 0.0f;
                inv[hh][1] = (lB > 0.0f) ? 1.0f / lB : 0.0f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(lempty);
            mbar_wait(&pvdone[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
            tc_fence_after();
            const bool odd = qd & 1;
            const bool vec_ok = (p.dh % 4 == 0) && (p.d % 4 == 0);                     // 8-byte aligned row segments
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int rowA = q * 32 + hh * 16 + rin, rowB = rowA + 8;             // rows of the tile
                const float invA = inv[hh][0], invB = inv[hh][1];
                uint32_t o0[16], o1[4];
                const uint32_t o_addr = tmem_base + ATT_O_COL + ((uint32_t)(q * 32 + hh * 16) << 16);
                tmem_ld_16x256b_x4(o_addr, o0);
                tmem_ld_16x256b_x1(o_addr + 32, o1);
                tmem_ld_wait_dep(o0); tmem_ld_wait_dep(o1);
                // packed pairs: pa[gq] = row A columns 8*gq + 2*qd, +1
                uint32_t pa[5], pb[5];
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    pa[gq] = attn_pack_bf16x2(__uint_as_float(o0[4 * gq]) * invA, __uint_as_float(o0[4 * gq + 1]) * invA);
                    pb[gq] = attn_pack_bf16x2(__uint_as_float(o0[4 * gq + 2]) * invB, __uint_as_float(o0[4 * gq + 3]) * invB);
                }
                pa[4] = attn_pack_bf16x2(__uint_as_float(o1[0]) * invA, __uint_as_float(o1[1]) * invA);
                pb[4] = attn_pack_bf16x2(__uint_as_float(o1[2]) * invB, __uint_as_float(o1[3]) * invB);
                const int iA = qt * ATT_BQ + rowA, iB = qt * ATT_BQ + rowB;           // frame inside the line
                const bool okA = iA < p.T, okB = iB < p.T;
                __nv_bfloat16 *dA = p.ctx + ((int64_t)b * p.T + iA) * p.d + hd * p.dh;
                __nv_bfloat16 *dB = p.ctx + ((int64_t)b * p.T + iB) * p.d + hd * p.dh;
#pragma unroll
                for (int gq = 0; gq < 4; gq += 2) {
                    // neighbouring lanes swap one pair: every lane then owns 4 consecutive columns (8 bytes)
                    const uint32_t sA = odd ? pa[gq] : pa[gq + 1], sB = odd ? pb[gq] : pb[gq + 1];
                    const uint32_t rA = __shfl_xor_sync(0xffffffffu, sA, 1), rB = __shfl_xor_sync(0xffffffffu, sB, 1);
                    const int col = odd ? 8 * (gq + 1) + 2 * (qd - 1) : 8 * gq + 2 * qd;
                    const uint2 oA = odd ? make_uint2(rA, pa[gq + 1]) : make_uint2(pa[gq], rA);
                    const uint2 oB = odd ? make_uint2(rB, pb[gq + 1]) : make_uint2(pb[gq], rB);
                    if (vec_ok && col + 4 <= p.dh) {
                        if (okA) *reinterpret_cast<uint2 *>(dA + col) = oA;
                        if (okB) *reinterpret_cast<uint2 *>(dB + col) = oB;
                    } else {
                        const __nv_bfloat16 *hA = reinterpret_cast<const __nv_bfloat16 *>(&oA), *hB = reinterpret_cast<const __nv_bfloat16 *>(&oB);
                        for (int e = 0; e < 4; ++e) {
                            if (col + e < p.dh) {
                                if (okA) dA[col + e] = hA[e];
                                if (okB) dB[col + e] = hB[e];
                            }
                        }
                    }
                }
                {
                    const int col = 32 + 2 * qd;
                    const __nv_bfloat16 *hA = reinterpret_cast<const __nv_bfloat16 *>(&pa[4]), *hB = reinterpret_cast<const __nv_bfloat16 *>(&pb[4]);
                    if (col + 2 <= p.dh && (p.dh % 2 == 0) && (p.d % 2 == 0)) {
                        if (okA) *reinterpret_cast<uint32_t *>(dA + col) = pa[4];
                        if (okB) *reinterpret_cast<uint32_t *>(dB + col) = pb[4];
                    } else {
                        for (int e = 0; e < 2; ++e) {
                            if (col + e < p.dh) {
                                if (okA) dA[col + e] = hA[e];
                                if (okB) dB[col + e] = hB[e];
                            }
                        }
                    }
                }
            }
            // O has been read (wait::ld): the first P V of the next item may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(oempty);
            ++na;
        }
    } else {
        // ================= softmax =================
        const int q = warp & 3, hf = warp >> 2;                  // TMEM lane quarter, half of the keys of a tile
        const int r = q * 32 + lane;                             // row in the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t g = 0, xr = 0, na = 0;                          // key tiles, exchange rounds, active items seen
        int tr_n = (threadIdx.x == 0) ? 0 : ATT_TRACE_EVENTS;
        for (Pos s = pos0; seq_valid(s); s = seq_next(s)) {
            const int Tk = valid_keys(s.b);
            const int ntiles = (Tk + ATT_BK - 1) / ATT_BK;
            if (ntiles == 0) continue;                           // the aux warps write the zero rows
            float m_run = -INFINITY, l_run = 0.0f;
            for (int j = 0; j < ntiles; ++j, ++g) {
                const uint32_t sb = g & 1u;
                ATT_TRACE(0, 300 + j);
                mbar_wait(&sfull[sb], (g >> 1) & 1u);
                ATT_TRACE(0, 310 + j);
                tc_fence_after();
                const uint32_t s_base = tmem_base + (sb ? ATT_S_COL1 : ATT_S_COL0) + lane_base;
                uint32_t v0[32], v1[32];
                tmem_ld_32x32b_x32(s_base + (uint32_t)(hf * 64), v0);
                tmem_ld_32x32b_x32(s_base + (uint32_t)(hf * 64 + 32), v1);
                tmem_ld_wait_dep(v0); tmem_ld_wait_dep(v1);
                const int nvalid = Tk - j * ATT_BK - hf * 64;        // keys >= nvalid of this thread's 64 are masked
                if (nvalid < 64) {
                    const uint32_t ninf = __float_as_uint(-INFINITY);
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (e >= nvalid) v0[e] = ninf;
                        if (32 + e >= nvalid) v1[e] = ninf;
                    }
                }
                float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    t0 = attn_max3(t0, __uint_as_float(v0[e]), __uint_as_float(v0[e + 1]));
                    t1 = attn_max3(t1, __uint_as_float(v1[e]), __uint_as_float(v1[e + 1]));
                }
                // row maximum over both halves (the partner warp holds the other 64 keys), exchanged as bf16: BOTH warps use
                // the rounded values of both halves, so they agree bit for bit on the reference point (any reference near
                // the maximum is valid).  The barrier also orders the partner's S loads before this warp's P stores into
                // the same TMEM columns.
                __nv_bfloat16 *xb = xmax + (xr & 1u) * 256;
                xb[hf * 128 + r] = __float2bfloat16_rn(fmaxf(t0, t1));
                named_bar_sync(1 + q, 64);
                const float m_new = fmaxf(m_run, fmaxf(__bfloat162float(xb[r]), __bfloat162float(xb[128 + r])));
                ++xr;
                if (j == 0) {
                    m_run = m_new;
                } else {
                    const bool need = (m_new - m_run) * p.scale_log2 > ATT_LAZY_LOG2;
                    if (__any_sync(0xffffffffu, need)) {       // same rows, same values in both warps of the quarter
                        const float corr = ex2_fast((m_run - m_new) * p.scale_log2);
                        m_run = m_new;
                        l_run *= corr;
                        mbar_wait(&pvdone[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
                        tc_fence_after();
                        // this warp rescales O columns [hf*32, hf*32+32) (48 columns cover dh <= 48)
                        if (hf == 0) {
                            uint32_t o[32];
                            tmem_ld_32x32b_x32(tmem_base + ATT_O_COL + lane_base, o);
                            tmem_ld_wait_dep(o);
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
                            tmem_st_32x32b_x32(tmem_base + ATT_O_COL + lane_base, o);
                        } else {
                            uint32_t o[16];
                            tmem_ld_32x32b_x16(tmem_base + ATT_O_COL + lane_base + 32, o);
                            tmem_ld_wait_dep(o);
#pragma unroll
                            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
                            tmem_st_32x32b_x16n(tmem_base + ATT_O_COL + lane_base + 32, o);
                        }
                    }
                }
                ATT_TRACE(0, 320 + j);
                // probabilities -> packed bf16 in TMEM: keys [hf*64, hf*64+64) of P live in columns [hf*32, hf*32+32)
                const float mb = m_run * p.scale_log2;
                const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nmb2 = make_float2(-mb, -mb);
                float2 ps = make_float2(0.f, 0.f);
                uint32_t pk[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float2 s2 = e < 16 ? make_float2(__uint_as_float(v0[2 * e]), __uint_as_float(v0[2 * e + 1]))
                                             : make_float2(__uint_as_float(v1[2 * e - 32]), __uint_as_float(v1[2 * e - 31]));
                    const float2 x = attn_fma2(s2, sc2, nmb2);
                    float2 pr;
                    if (e % ATT_POLY_EVERY == ATT_POLY_EVERY - 1) pr = attn_ex2_poly2(x);      // this pair on the FMA pipe
                    else pr = make_float2(ex2_fast(x.x), ex2_fast(x.y));
                    ps = attn_add2(ps, pr);
                    pk[e] = attn_pack_bf16x2(pr.x, pr.y);
                }
                const float ps0 = ps.x, ps1 = ps.y;
                tmem_st_32x32b_x32(s_base + (uint32_t)(hf * 32), pk);
                l_run += ps0 + ps1;
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pfull[sb]);
                ATT_TRACE(0, 340 + j);
            }
            // hand the row sums to the aux warps (single buffer: wait until they have read the previous item's)
            mbar_wait(lempty, (na & 1u) ^ 1u);
            xl[hf * 128 + r] = l_run;
            __syncwarp();
            if (lane == 0) mbar_arrive(lfull);
            ++na;
            ATT_TRACE(0, 360);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ATT_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

}  // namespace cocr
