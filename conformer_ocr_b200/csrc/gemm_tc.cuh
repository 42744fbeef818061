// Persistent warp-specialised bf16 GEMM on tcgen05 tensor cores:  C[M,N] = epi(A[M,K] * W[N,K]^T + bias)
//
//   * A (activations, row-major bf16) streams through a ring of 128x64 SWIZZLE_128B tiles loaded by TMA;
//   * W (weights, row-major [N,K] bf16, i.e. K-major "B" operand) is either RESIDENT in shared memory for
//     the whole kernel (one N-slice per CTA; all K=144 layers) or streamed through the same ring (K=576/768);
//   * one elected thread issues tcgen05.mma (UMMA 128 x block_n x 16) into a double-buffered TMEM accumulator;
//   * 8 epilogue warps (two per TMEM lane quarter) drain the accumulator.
//
// Epilogues
//   bf16 outputs (bias | bias+SiLU | bias+position handling): thread = row (tcgen05.ld 32x32b), the next 32-column
//     piece is loaded from TMEM while the current one is converted; each piece goes through a per-warp ring of
//     SWIZZLE_64B staging tiles and leaves with one asynchronous TMA store (cp.async.bulk.tensor), so no thread
//     ever waits on a global store and rows / columns outside the matrix are clipped by the tensor map;
//   residual (EPI_RESID):  x <- x + alpha*(acc + bias) [-> LayerNorm_final(x)] ; xn = bf16(LayerNorm_next(x))
//     keeps the fp32 residual stream and both LayerNorms of a conformer block inside the GEMM
//     (modules.py:31-32, encoder.py:99, feed_forward.py:46, attention.py:148, convolution.py:136).  It reads TMEM in
//     the 16x256b fragment layout (a quad of lanes owns 8 consecutive columns of a row), so x is read and written
//     straight from registers in full 32-byte sectors, LayerNorm statistics need two quad shuffles, and no shared
//     memory or barrier is involved; the two warps of a lane quarter take alternate tiles;
//   fp32 logits and the fused arg-max (CTC decode; logits never stored) keep a staged scalar path (cold).
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
#pragma once

#include <cuda_fp16.h>

#include "tc_utils.cuh"

namespace cocr {

enum GemmEpi : int {
    EPI_BF16 = 0,        // out bf16 = acc + bias
    EPI_BF16_SILU = 1,   // out bf16 = silu(2*(acc + bias)): the caller passes weights and bias scaled by 1/2
    EPI_F32 = 2,         // out fp32 = acc + bias
    EPI_ARGMAX = 3,      // best[row] = max over columns of (acc + bias), lowest column wins ties (packed u64)
    EPI_BF16_POS = 4,    // out bf16 = acc + bias, with per-16-column-group position handling (attention operands):
                         //   kind 0 plain; kind 1 rotate pairs by the row's angle (relative-position term);
                         //   kind 2 fill with (cos, sin) of the row's angle
    EPI_RESID = 5,       // fp32 residual stream update + LayerNorm(s); N == block_n == 8 * GEMM_RESID_NG (one N tile)
    EPI_BF16_GLU = 6,    // GLU of interleaved (value, gate) column pairs: out[:, c] = v * sigmoid(g), N/2 bf16 columns.  The caller
                         // passes value AND gate weights / biases scaled by 1/2:  v*sigmoid(g) = hv + hv*tanh(hg)
};

struct GemmArgs {
    int M, N, K;             // K % 16 == 0
    int block_n;             // UMMA N: multiple of 16, 16..256 (multiple of 32 when there is more than one N tile)
    int n_tiles, m_tiles;
    int nkc;                 // ceil(K / 64) k-chunks
    int last_ksteps;         // k-steps (of 16) in the last chunk
    int resident;            // 1: W slice resident in smem (grid is a multiple of n_tiles)
    int stages;              // ring slots
    int tmem_cols;           // power of two >= 2*block_n
    const float *bias;       // (N) or null
    void *out;               // fp32 (M, ldc) for EPI_F32; EPI_RESID: bf16 xn (M, ldc); bf16 epilogues write through tmC
    int ldc;
    unsigned long long *best;   // EPI_ARGMAX: (M) packed keys, zero-initialised by the caller
    // EPI_BF16_POS
    const int16_t *col_meta;    // (N/16): kind << 12 | first pair index
    const int16_t *piece_dst;   // (N/32): destination column / 32 of every 32-column piece (the GEMM columns are a
                                // permutation of the output columns, arranged so that the pieces one warp handles in a
                                // tile share their rotation angles), or null = identity
    const float *omega;         // (n_omega) angular frequencies: the angle of pair m at frame t is t * omega[m]
    int n_omega, T;             // row r of the GEMM is frame t = r % T of its line
    // EPI_RESID
    float *x;                   // (M, N) fp32 residual stream, updated in place.  PRIVATE LAYOUT (only these epilogues and
                                // k_resid_unpermute touch it): inside a row, each pair of 8-column groups (16 columns) is
                                // stored as 4 x float4, one per lane of the owning quad: float4 q of pair gp holds columns
                                // (16gp + 2q, +1, 16gp + 8 + 2q, +1).  Every lane then moves 16 bytes per access and a quad
                                // covers 64 contiguous bytes: half the memory requests of the plain fragment layout.
    float alpha;                // module factor (0.5 for the half-step FFNs)
    int init;                   // 1: x = acc + bias (start of the stream)
    const float *gf, *bf;       // block-final LayerNorm applied to x itself (or null)
    const float *g1, *b1;       // LayerNorm of the next module -> xn (null: xn = bf16(x))
    int relu;                   // EPI_BF16: out = max(acc + bias, 0)
    int dbg;                    // experiments (COCR_GEMM_DBG): 1 skip the TMA stores, 2 skip the position math
    long long *trace;           // debug: (tag, clock) pairs of CTA 0 (COCR_GEMM_TRACE=<tag>:<file>), else null
};

constexpr int GEMM_TRACE_EVENTS = 1024;    // per role: 0 epilogue warp 0, 1 MMA warp, 2 producer
#define GEMM_TRACE(role, tag)                                                                     \
    do {                                                                                          \
        if (p.trace && blockIdx.x == 0 && tr_n < GEMM_TRACE_EVENTS) {                             \
            p.trace[((role) * GEMM_TRACE_EVENTS + tr_n) * 2] = (tag);                             \
            p.trace[((role) * GEMM_TRACE_EVENTS + tr_n) * 2 + 1] = clock64();                     \
            ++tr_n;                                                                               \
        }                                                                                         \
    } while (0)

// epilogue warps: the thread-per-row bf16 epilogues are chains of dependent latencies (TMEM load -> math -> staging ->
// proxy fence -> TMA store), so they run 16 warps (4 per scheduler, <= 112 registers) to overlap them.  The residual
// epilogue also runs 16 warps since it handles ONE row per lane quad (96 registers, no spills; the two-row form needed
// 168 and spilled at 16 warps): it is bound by the latency of its own x loads / stores, so rows in flight per SM are what
// counts (attention-out / pw2: 37 -> 34 us).
__host__ __device__ constexpr bool gemm_row_pieces(int epi) { return epi == 0 || epi == 1 || epi == 4 || epi == 6; }
__host__ __device__ constexpr int gemm_ne(int epi) { return (gemm_row_pieces(epi) || epi == 5) ? 16 : 8; }
__host__ __device__ constexpr int gemm_threads(int epi) { return 64 + 32 * gemm_ne(epi); }
// staging tiles per warp.  The positional projection (N tiles of 256 columns, 96 KB of resident weights) keeps one: the
// 32 KB it frees are two more A slots (5 instead of 3, i.e. the next tile's A is in flight while this one computes; 75.7 -> 74.1 us)
__host__ __device__ constexpr int gemm_nstg(int epi) { return epi == 4 ? 1 : (gemm_row_pieces(epi) ? 2 : 4); }
constexpr int GEMM_BM = 128;
constexpr int GEMM_KC = 64;                       // k-chunk (elements) = one 128-byte swizzle span of bf16
constexpr int GEMM_RESID_NG = 18;                 // EPI_RESID: 8-column groups per row (encoder_dim 144)
constexpr uint32_t GEMM_A_CHUNK = GEMM_BM * 128;  // bytes of one A k-chunk tile
constexpr uint32_t GEMM_STAGE_BYTES = 2048;       // one staging tile: 32 rows x 64 B (32 bf16 / 16 fp32 columns)
constexpr uint32_t GEMM_MISC_SMEM = 2 * 256 * 4 + 4 * 256 * 4 + 512 + 1024;   // bias, LayerNorm / position tables, barriers, slack

__host__ __device__ inline uint32_t gemm_b_chunk_bytes(int block_n) { return (uint32_t)block_n * 128u; }
__host__ __device__ inline uint32_t gemm_fixed_smem(int epi) {
    return GEMM_MISC_SMEM + (epi == EPI_RESID ? 0u : (uint32_t)(gemm_ne(epi) * gemm_nstg(epi)) * GEMM_STAGE_BYTES);
}

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

// packed fp32 pairs (sm_100 FADD2 / FFMA2: one issue slot for two lanes of arithmetic)
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
        "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// EPI_BF16_SILU receives h = x/2 (the FFN-up weights and bias are pre-scaled by 1/2, exact in binary floating point):
// silu(x) = x*sigmoid(x) = h + h*tanh(h).  One MUFU and half an FFMA2 per element.
__device__ __forceinline__ float2 silu_from_half(float2 h) {
    return ffma2(h, make_float2(tanh_fast(h.x), tanh_fast(h.y)), h);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h2);
}

// byte offset of (row, 16-byte chunk 0..3) in a [32 rows][64 B] tile in the SWIZZLE_64B pattern (address bits 4-5 xor
// bits 7-8): conflict-free for row-wise 16-byte writes (thread = row), and what a SWIZZLE_64B tensor map expects
__device__ __forceinline__ uint32_t stg64_offset(uint32_t row, uint32_t chunk) {
    return row * 64u + (((chunk ^ (row >> 1)) & 3u) << 4);
}

template <int EPI>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmC, const GemmArgs p) {
    using namespace tc;
    // SWIZZLE_128B tiles need 1024-byte alignment.  The array is used directly (no integer round-up of the pointer):
    // that keeps every access in the shared state space (LDS/STS); through a laundered pointer the compiler emits
    // generic LD/ST and a full MEMBAR.ALL.CTA in front of every fence.proxy.async.
    extern __shared__ __align__(1024) uint8_t smem[];

    constexpr int GEMM_NE = gemm_ne(EPI), GEMM_THREADS = gemm_threads(EPI), GEMM_NSTG = gemm_nstg(EPI);
    constexpr bool kTmaStore = (EPI == EPI_BF16 || EPI == EPI_BF16_SILU || EPI == EPI_BF16_POS);
    constexpr bool kRowPieces = kTmaStore || EPI == EPI_BF16_GLU;      // thread = row, 32-column pieces, TMEM loads one ahead
    constexpr uint32_t kStagingBytes = (EPI == EPI_RESID) ? 0u : (uint32_t)GEMM_NE * GEMM_NSTG * GEMM_STAGE_BYTES;
    const uint32_t b_chunk = gemm_b_chunk_bytes(p.block_n);
    const uint32_t bres_bytes = p.resident ? (uint32_t)p.nkc * b_chunk : 0u;
    const uint32_t slot_bytes = GEMM_A_CHUNK + (p.resident ? 0u : b_chunk);
    uint8_t *bres = smem;
    uint8_t *ring = smem + bres_bytes;
    uint8_t *staging = ring + (uint32_t)p.stages * slot_bytes;                        // NE warps x NSTG x 2 KB
    float *sbias = reinterpret_cast<float *>(staging + kStagingBytes);                  // [2][256] bias (per acc stage)
    float *sln = sbias + 512;                                                           // [4][256] gf, bf, g1, b1  | position tables
    uint64_t *bars = reinterpret_cast<uint64_t *>(sln + 1024);
    uint64_t *full = bars;                       // [stages]
    uint64_t *empty = bars + p.stages;           // [stages]
    uint64_t *bfull = bars + 2 * p.stages;       // [1]
    uint64_t *tfull = bfull + 1;                 // [2]
    uint64_t *tempty = tfull + 2;                // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    // EPI_BF16_POS tables (alias of sln)
    float *somega = sln;                                               // [<= 256]
    int16_t *smeta = reinterpret_cast<int16_t *>(sln + 256);           // [<= 1024]
    int16_t *spdst = reinterpret_cast<int16_t *>(sln + 768);           // [<= 512]

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
    // one N offset for the whole kernel: the bias slice is staged once, before the roles start
    const bool bias_once = p.resident || p.n_tiles == 1;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0u) __trap();
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if constexpr (kTmaStore) prefetch_tmap(&tmC);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(bfull, 1);
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        constexpr uint32_t drainers = GEMM_NE;                                       // warps draining one accumulator stage
        mbar_init(&tempty[0], drainers); mbar_init(&tempty[1], drainers);
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
    if constexpr (EPI == EPI_BF16_POS) {
        for (int c = threadIdx.x; c < p.n_omega; c += GEMM_THREADS) somega[c] = p.omega[c];
        for (int c = threadIdx.x; c < p.N / 16; c += GEMM_THREADS) smeta[c] = p.col_meta[c];
        for (int c = threadIdx.x; c < (p.N + 31) / 32; c += GEMM_THREADS) spdst[c] = p.piece_dst ? p.piece_dst[c] : (int16_t)c;
    }
    if (bias_once) {
        const int n0 = n_fixed * p.block_n;
        for (int c = threadIdx.x; c < 256; c += GEMM_THREADS) {
            const float b = (p.bias && c < p.block_n && n0 + c < p.N) ? p.bias[n0 + c] : 0.0f;
            sbias[c] = b;
            sbias[256 + c] = b;
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
            int tr_n = 0;
            for (int j = 0; j < my_count; ++j) {
                int mt, nt;
                tile_mn(j, mt, nt);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&empty[slot], phase ^ 1u);
                    GEMM_TRACE(2, 100 + kc);
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
        // all lanes walk the loop (uniform values), one elected lane issues: a divergent `lane == 0` branch would make
        // the compiler wrap every UTCHMMA in an elect/branch loop
        {
            const uint32_t idesc = make_idesc_bf16(GEMM_BM, p.block_n);
            if (p.resident && my_count > 0) mbar_wait(bfull, 0);
            int slot = 0;
            uint32_t phase = 0;
            int tr_n = (lane == 0) ? 0 : GEMM_TRACE_EVENTS;
            for (int j = 0; j < my_count; ++j) {
                const int acc = j & 1;
                const uint32_t acc_phase = (uint32_t)(j >> 1) & 1u;
                GEMM_TRACE(1, 200);
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                GEMM_TRACE(1, 201);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(&full[slot], phase);
                    GEMM_TRACE(1, 210 + kc);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(ring + (uint32_t)slot * slot_bytes);
                    const uint32_t b_base = p.resident ? smem_u32(bres + (uint32_t)kc * b_chunk) : a_base + GEMM_A_CHUNK;
                    const int nks = (kc == p.nkc - 1) ? p.last_ksteps : 4;
                    if (elect_one()) {
                        for (int ks = 0; ks < nks; ++ks) {
                            const uint64_t ad = make_smem_desc_sw128(a_base + ks * 32, 1024);
                            const uint64_t bd = make_smem_desc_sw128(b_base + ks * 32, 1024);
                            umma_bf16_ss(d_tmem, ad, bd, idesc, (kc | ks) ? 1u : 0u);
                        }
                        umma_commit(&empty[slot]);
                        if (kc == p.nkc - 1) umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++slot == p.stages) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ================= epilogue =================
        const int ew = warp - 2;                         // 0 .. NE-1
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = ew >> 2;                        // which of the NSH warps of the quarter
        constexpr int NSH = GEMM_NE / 4;                 // warps sharing one lane quarter
        const int et = ew * 32 + lane;                   // 0 .. 32*NE-1
        uint8_t *stage_base = staging + (uint32_t)ew * GEMM_NSTG * GEMM_STAGE_BYTES;
        const int npieces = (p.block_n + 31) / 32;       // 32-column pieces
        uint32_t bufctr = 0;                             // staging ring position (TMA-store epilogues)
        int tr_n = (threadIdx.x == 64) ? 0 : GEMM_TRACE_EVENTS;
        for (int j = 0; j < my_count; ++j) {
            int mt, nt;
            tile_mn(j, mt, nt);
            const int acc = j & 1;
            const uint32_t acc_phase = (uint32_t)(j >> 1) & 1u;
            const int n0 = nt * p.block_n;
            float *sb = sbias + acc * 256;
            if (!bias_once) {
                if (et < p.block_n) sb[et] = (p.bias && n0 + et < p.N) ? p.bias[n0 + et] : 0.0f;
                named_bar_sync(1, 32 * GEMM_NE);
            }
            const uint32_t t_acc = tmem_base + (uint32_t)(acc * p.block_n);
            const uint32_t t_row = t_acc + ((uint32_t)(q * 32) << 16);
            const int grow0 = mt * GEMM_BM + q * 32;                 // first global row of this warp
            auto release_acc = [&]() {                               // this warp is done reading the accumulator stage
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            };

            if constexpr (EPI == EPI_RESID) {
                // ---- fp32 residual update + LayerNorm(s), fragment layout: 16 warps, 8 rows of the tile each, ONE row per lane
                // quad (the 16x256b load returns rows r and r+8 together; a warp keeps the one it owns).  The epilogue is
                // bound by the latency of its own x loads and stores, so what counts is how many rows are in flight per SM.
                GEMM_TRACE(0, 400);
                constexpr int NG = GEMM_RESID_NG;
                const int r_in = lane >> 2, qd = lane & 3;
                const int hh = half >> 1, pass = half & 1;            // 16-row half of the lane quarter, row r / r+8 of the load
                const float inv_n = 1.0f / (float)(8 * NG);
                const int64_t row = (int64_t)grow0 + hh * 16 + pass * 8 + r_in;
                const bool ok = row < p.M;
                float2 v[NG];
                float *xr = p.x + row * (8 * NG) + 4 * qd;
                if (!p.init) {
#pragma unroll
                    for (int gp = 0; gp < NG / 2; ++gp) {
                        const float4 ta = ok ? *reinterpret_cast<const float4 *>(xr + 16 * gp) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[2 * gp] = make_float2(ta.x, ta.y); v[2 * gp + 1] = make_float2(ta.z, ta.w);
                    }
                    if (j + 1 < my_count) {
                        // this warp's rows of the NEXT tile into L2 (8 rows x 576 B = 36 lines of 128 B)
                        int mt2, nt2;
                        tile_mn(j + 1, mt2, nt2);
                        const int64_t r2 = (int64_t)mt2 * GEMM_BM + q * 32 + hh * 16 + pass * 8;
                        const char *nb = reinterpret_cast<const char *>(p.x + r2 * (8 * NG));
                        const int64_t lim = ((int64_t)p.M - r2) * (8 * NG) * 4;
                        for (int o = lane * 128; o < 8 * 8 * NG * 4 && o < lim; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + o));
                    }
                }
                GEMM_TRACE(0, 410);
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                GEMM_TRACE(0, 420);
                const uint32_t taddr = t_acc + ((uint32_t)(q * 32 + hh * 16) << 16);
#pragma unroll
                for (int blk = 0; blk < NG / 2; ++blk) {              // 16 columns (two groups) per load keeps the live set small
                    uint32_t t1[8];
                    tmem_ld_16x256b_x2(taddr + 16u * blk, t1);
                    tmem_ld_wait_dep(t1);
                    if (blk == NG / 2 - 1) release_acc();
#pragma unroll
                    for (int gg = 0; gg < 2; ++gg) {
                        const int g = 2 * blk + gg;
                        const float2 b2 = *reinterpret_cast<const float2 *>(sb + 8 * g + 2 * qd);
                        const float a0 = __uint_as_float(pass ? t1[4 * gg + 2] : t1[4 * gg]) + b2.x;
                        const float a1 = __uint_as_float(pass ? t1[4 * gg + 3] : t1[4 * gg + 1]) + b2.y;
                        if (p.init) v[g] = make_float2(a0, a1);
                        else { v[g].x = fmaf(p.alpha, a0, v[g].x); v[g].y = fmaf(p.alpha, a1, v[g].y); }
                    }
                }
                GEMM_TRACE(0, 430);
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
                        const float2 bb = *reinterpret_cast<const float2 *>(sln + 256 + 8 * g + 2 * qd);
                        v[g].x = (v[g].x - m) * r * gg.x + bb.x; v[g].y = (v[g].y - m) * r * gg.y + bb.y;
                    }
                }
                // x_new: 8 rows x 64 contiguous bytes per store instruction (private layout, see GemmArgs::x)
#pragma unroll
                for (int gp = 0; gp < NG / 2; ++gp)
                    if (ok) *reinterpret_cast<float4 *>(xr + 16 * gp) = make_float4(v[2 * gp].x, v[2 * gp].y, v[2 * gp + 1].x, v[2 * gp + 1].y);
                // xn = bf16(LayerNorm_next(x_new)) (or bf16(x_new))
                float m = 0.f, r = 1.f;
                if (p.g1) stats(m, r);
                uint32_t pa[NG];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float a0 = v[g].x, a1 = v[g].y;
                    if (p.g1) {
                        const float2 gg = *reinterpret_cast<const float2 *>(sln + 512 + 8 * g + 2 * qd);
                        const float2 bb = *reinterpret_cast<const float2 *>(sln + 768 + 8 * g + 2 * qd);
                        a0 = (a0 - m) * r * gg.x + bb.x; a1 = (a1 - m) * r * gg.y + bb.y;
                    }
                    pa[g] = pack_bf16x2(a0, a1);
                }
                // neighbouring lanes swap one packed pair so that every lane stores 8 bytes and a quad writes the
                // 32 contiguous bytes of two column groups
                __nv_bfloat16 *na = reinterpret_cast<__nv_bfloat16 *>(p.out) + row * p.ldc;
                const bool odd = qd & 1;
#pragma unroll
                for (int g = 0; g < NG; g += 2) {
                    const uint32_t send = odd ? pa[g] : pa[g + 1];
                    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
                    const int col = odd ? 8 * (g + 1) + 2 * (qd - 1) : 8 * g + 2 * qd;
                    const uint2 o = odd ? make_uint2(recv, pa[g + 1]) : make_uint2(pa[g], recv);
                    if (ok) *reinterpret_cast<uint2 *>(na + col) = o;
                }
                GEMM_TRACE(0, 440);
            } else if constexpr (kRowPieces) {
                // ---- bf16 output through per-warp staging tiles and TMA stores; TMEM loads run one piece ahead ----
                const int my_np = (npieces - half + NSH - 1) / NSH;        // pieces half, half + 2, ...
                GEMM_TRACE(0, 300);
                mbar_wait(&tfull[acc], acc_phase);
                GEMM_TRACE(0, 301);
                tc_fence_after();
                float tf = 0.0f;
                if constexpr (EPI == EPI_BF16_POS) tf = (float)((grow0 + lane) % p.T);
                uint32_t ra[32];
                auto issue = [&](int k, uint32_t (&r)[32]) {
                    const int c0 = (half + NSH * k) * 32;
                    if (p.block_n - c0 >= 32) tmem_ld_32x32b_x32(t_row + (uint32_t)c0, r);
                    else tmem_ld_32x32b_x16_lo(t_row + (uint32_t)c0, r);
                };
                auto process = [&](int k, uint32_t (&r)[32]) {
                    const int c0 = (half + NSH * k) * 32;
                    const int gcol = n0 + c0;
                    if (gcol >= p.N) return;                               // warp-uniform
                    if (p.block_n - c0 < 32) {
#pragma unroll
                        for (int i = 16; i < 32; ++i) r[i] = 0u;
                    }
                    float2 f2[16];                                         // (column 2e, column 2e+1) of this thread's row
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(sb + c0 + e4 * 4);
                        f2[2 * e4] = fadd2(make_float2(__uint_as_float(r[4 * e4]), __uint_as_float(r[4 * e4 + 1])), make_float2(b4.x, b4.y));
                        f2[2 * e4 + 1] = fadd2(make_float2(__uint_as_float(r[4 * e4 + 2]), __uint_as_float(r[4 * e4 + 3])), make_float2(b4.z, b4.w));
                    }
                    if constexpr (EPI == EPI_BF16_POS) {
                        // rotation / fill of the positional columns by the angle of this thread's frame
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const int meta = smeta[(gcol >> 4) + g];
                            const int kind = meta >> 12;                   // warp-uniform
                            if (kind != 0 && !(p.dbg & 2)) {
                                const int pbase = meta & 0xFFF;            // multiple of 4 (16-byte aligned omega reads)
#pragma unroll
                                for (int u4 = 0; u4 < 2; ++u4) {
                                    const float4 om = *reinterpret_cast<const float4 *>(somega + pbase + 4 * u4);
                                    const float omv[4] = {om.x, om.y, om.z, om.w};
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        float sn, cs;
                                        __sincosf(tf * omv[u], &sn, &cs);
                                        float2 &v = f2[g * 8 + u4 * 4 + u];
                                        const float as = v.x, ac = v.y;
                                        if (kind == 1) v = make_float2(fmaf(as, sn, ac * cs), fmaf(ac, sn, -as * cs));
                                        else v = make_float2(cs, sn);
                                    }
                                }
                            }
                        }
                    }
                    if constexpr (EPI == EPI_BF16_GLU) {
                        // f2[e] = (value/2, gate/2) of output column (gcol + 2e)/2: 16 outputs = 32 contiguous bytes per row,
                        // written straight from registers (two full-sector halves per thread)
                        uint32_t po[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float o0 = fmaf(f2[2 * e].x, tanh_fast(f2[2 * e].y), f2[2 * e].x);
                            const float o1 = fmaf(f2[2 * e + 1].x, tanh_fast(f2[2 * e + 1].y), f2[2 * e + 1].x);
                            po[e] = pack_bf16x2(o0, o1);
                        }
                        const int64_t grow = (int64_t)grow0 + lane;
                        if (grow < p.M) {
                            __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(p.out) + grow * p.ldc + (gcol >> 1);
                            *reinterpret_cast<uint4 *>(dst) = make_uint4(po[0], po[1], po[2], po[3]);
                            // a 16-column last piece (block_n = 144) holds 8 outputs: the other 8 belong to the NEXT N tile's first piece
                            // (written by that tile's CTA) - storing the zero-filled half here raced with it
                            if (p.block_n - c0 >= 32 && gcol + 16 < p.N) *reinterpret_cast<uint4 *>(dst + 8) = make_uint4(po[4], po[5], po[6], po[7]);
                        }
                        return;
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        if constexpr (EPI == EPI_BF16_SILU) f2[e] = silu_from_half(f2[e]);
                        if constexpr (EPI == EPI_BF16) {
                            if (p.relu) f2[e] = make_float2(fmaxf(f2[e].x, 0.0f), fmaxf(f2[e].y, 0.0f));
                        }
                        pk[e] = pack_bf16x2(f2[e].x, f2[e].y);
                    }
                    uint8_t *sbuf = stage_base + (bufctr % GEMM_NSTG) * GEMM_STAGE_BYTES;
                    ++bufctr;
                    if (lane == 0 && !(p.dbg & 16)) tma_store_wait_read<GEMM_NSTG - 1>();    // the store that last used this tile has read it
                    if (!(p.dbg & 32)) __syncwarp();
                    if (!(p.dbg & 8)) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            *reinterpret_cast<uint4 *>(sbuf + stg64_offset(lane, jj)) = make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
                    } else if (pk[0] == 0x12345678u && pk[7] == 0x9abcdef0u && pk[15] == 1u) {
                        *reinterpret_cast<uint4 *>(sbuf) = make_uint4(pk[1], pk[5], pk[9], pk[13]);
                    }
                    if (!(p.dbg & 4)) fence_proxy_async_smem();
                    if (!(p.dbg & 32)) __syncwarp();
                    if (lane == 0 && !(p.dbg & 1)) {
                        int dcol = gcol;
                        if constexpr (EPI == EPI_BF16_POS) dcol = (int)spdst[gcol >> 5] * 32;
                        tma_store_2d(&tmC, sbuf, dcol, grow0);
                        tma_store_commit();
                    }
                };
                if (p.dbg & 256) {              // experiment: main loop only, no epilogue work
                    release_acc();
                    continue;
                }
                if (my_np == 0) release_acc();
                for (int k = 0; k < my_np; ++k) {
                    issue(k, ra);
                    tmem_ld_wait_dep(ra);
                    GEMM_TRACE(0, 310 + k);
                    if (k + 1 == my_np) release_acc();          // last TMEM read of this warp for this tile
                    process(k, ra);
                    GEMM_TRACE(0, 320 + k);
                }
            } else {
                // ---- fp32 logits / fused arg-max (cold paths) ----
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                float best_v = -INFINITY;
                int best_c = 0x7fffffff;
                int buf = 0;
                for (int pi = half; pi < npieces; pi += NSH) {
                    const int c0 = pi * 32;
                    const int ncol = min(32, p.block_n - c0);            // 32 or 16
                    uint32_t r[32];
                    if (ncol == 32) tmem_ld_32x32b_x32(t_row + (uint32_t)c0, r);
                    else tmem_ld_32x32b_x16_lo(t_row + (uint32_t)c0, r);
                    tmem_ld_wait_dep(r);
                    if (ncol < 32) {
#pragma unroll
                        for (int i = 16; i < 32; ++i) r[i] = 0u;
                    }
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
                    } else {
                        // EPI_F32: two 16-column halves, each a [32 rows][64 B] staging tile, coalesced 16-byte row segments
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
                                    const int64_t grow = (int64_t)grow0 + row;
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
                                __syncwarp();
                                buf ^= 1;
                            }
                        }
                    }
                }
                if constexpr (EPI == EPI_ARGMAX) {
                    const int64_t grow = (int64_t)grow0 + lane;
                    if (grow < p.M && best_c != 0x7fffffff) atomicMax(p.best + grow, pack_best(best_v, best_c));
                }
                release_acc();
            }
        }
        if constexpr (kTmaStore) {
            if (lane == 0) tma_store_wait_all<0>();      // every store of this warp has landed before the CTA retires
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
