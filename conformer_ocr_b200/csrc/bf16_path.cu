// bf16 tensor-core path (cocr_precision COCR_BF16).
//
// Every dense contraction (subsampling out-projection, FFN, QKV / output projections, pointwise convs,
// CTC projection) runs on tcgen05 tensor cores through k_gemm_tc (gemm_tc.cuh): bf16 operands staged by
// TMA, fp32 accumulation in TMEM.  The residual stream, LayerNorm statistics and softmax stay fp32.
// Reference symbols: same as fp32_path.cu.
#include <math.h>

#include <algorithm>

#include "common.h"
#include "gemm_tc.cuh"
#include "attn_tc.cuh"
#include "ffn_tc.cuh"
#include "ffn2_tc.cuh"

namespace cocr {

typedef __nv_bfloat16 bf16;

struct LayerGraph {      // captured encoder-layer sequence for one (batch, width, mask, buffers) combination
    int B, T, T2;        // T2 (frames after the first conv) fixes the workspace offsets: widths that share T may differ in T2
    const void *lens, *ws, *etab, *pos;
    cudaGraphExec_t exec;
    int64_t launches;
    int state;           // 1 seen once (ran eagerly), 2 captured, -1 capture failed: always eager
    uint64_t last_use;   // LRU stamp
};

struct Bf16State {
    bf16 *arena = nullptr;        // bf16 mirror of the fp32 weight arena (same offsets)
    int smem_optin = 0;
    // tensor-core attention (attn_tc.cuh): fused Qe|Kx|Vx projection weights and position tables
    bool tc_attn = false;
    int qe_cols = 0, nx = 0, nchunk = 0;
    bf16 *w_qkvx = nullptr;       // (L, nx, d)
    float *b_qkvx = nullptr;      // (L, nx)
    int16_t *meta = nullptr;      // (nx / 16), in GEMM column order
    int16_t *piece_dst = nullptr; // (nx / 32): output column / 32 of every 32-column piece of the GEMM
    float *omega = nullptr;       // (d/2) angular frequencies omega_m (embedding.py:42-44)
    bf16 *etab = nullptr;         // (T, d - 16) (cos, sin) interleaved
    // conv module: first pointwise conv with (value, gate) rows interleaved and scaled by 1/2 for the GLU epilogue
    bf16 *w_pw1g = nullptr;       // (L, 2d, d)
    float *b_pw1g = nullptr;      // (L, 2d)
    bf16 *w_fe_pw2 = nullptr;     // (cc, 2 cc): [Wpw | Wpw] for the hi/lo split 1x1 conv GEMM of the front-end
    bf16 *w_ffn2p = nullptr;      // (L, 2, ff/96 * d, 96): FFN-down weights cut into 96-column k-chunks stacked along the rows
                                  // (row c*d + n, column k = W2[n][c*96 + k]) for the fused FFN kernel (ffn_tc.cuh)
    bf16 *w_ffn1p = nullptr;      // (L, 2, ff, 160): FFN-up weights (scaled by 1/2) with the bias split into bf16 hi / lo parts in
                                  // columns 144 / 145 (the xn rows carry ones there), zeros up to column 159
    float *dw_taps = nullptr;     // (L, K*d + d): depthwise taps [k][pair](2) * bn_scale / 2, then bn_shift / 2
    int tab_T = 0;
    std::vector<LayerGraph> graphs;
    uint64_t graph_clock = 0;
    cudaStream_t gstream = nullptr;
    cudaEvent_t g_in = nullptr, g_out = nullptr;
};

// ------------------------------------------------------------------------------------------
// driver entry point for tensor maps (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D row-major bf16 matrix (rows x cols, leading dimension ld elements) -> box (box_cols x box_rows).
// Operand tiles use 64-column SWIZZLE_128B boxes; epilogue stores use 32-column SWIZZLE_64B boxes.
static int make_tmap_bf16(cocr_handle *h, CUtensorMap *m, const void *base, int64_t rows, int64_t cols, int64_t ld,
                          int box_rows, int box_cols = 64, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return set_error(h, COCR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(h, COCR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
                         (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------------
static int pick_block_n(int N) {
    if (N <= 256) return (N + 15) / 16 * 16;
    // several N tiles: the epilogue stores 32-column pieces, so tiles are multiples of 32; prefer an exact divisor
    for (int bn = 256; bn >= 64; bn -= 32)
        if (N % bn == 0) return bn;
    return 256;
}

struct PosArgs { const int16_t *meta; const float *omega; int n_omega, T; const int16_t *piece_dst; };
struct ResidArgs { float *x; float alpha; int init; const float *gf, *bf, *g1, *b1; };

int launch_gemm_tc(cocr_handle *h, int epi, const bf16 *A, int lda, const bf16 *W, int ldw, const float *bias, void *out,
                   int ldc, unsigned long long *best, int M, int N, int K, cudaStream_t st, int block_n_override = 0,
                   const PosArgs *pos = nullptr, const char *tag = "gemm_tc", const ResidArgs *res = nullptr, int relu = 0) {
    if (M <= 0 || N <= 0) return COCR_OK;
    if (K % 16 != 0 || lda % 8 != 0 || ldw % 8 != 0)
        return set_error(h, COCR_ERR_UNSUPPORTED, "tensor-core GEMM needs K %% 16 == 0 and 16-byte aligned rows (K=%d lda=%d ldw=%d)", K,
                         lda, ldw);
    GemmArgs p;
    p.M = M; p.N = N; p.K = K;
    p.block_n = block_n_override ? block_n_override : pick_block_n(N);
    p.n_tiles = (N + p.block_n - 1) / p.block_n;
    p.m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    p.nkc = (K + GEMM_KC - 1) / GEMM_KC;
    p.last_ksteps = (K - (p.nkc - 1) * GEMM_KC) / 16;
    p.bias = bias; p.out = out; p.ldc = ldc; p.best = best; p.relu = relu;
    p.col_meta = pos ? pos->meta : nullptr; p.omega = pos ? pos->omega : nullptr;
    p.n_omega = pos ? pos->n_omega : 0; p.T = pos ? pos->T : 1; p.piece_dst = pos ? pos->piece_dst : nullptr;
    if (epi == EPI_BF16_POS && (!pos || p.n_omega > 256 || N / 16 > 1024)) return set_error(h, COCR_ERR_BAD_ARG, "POS epilogue needs position tables");
    if (epi == EPI_BF16_GLU && (N % 32 != 0 || ldc % 8 != 0)) return set_error(h, COCR_ERR_BAD_ARG, "GLU epilogue needs N %% 32 == 0 and 16-byte aligned output rows");
    // the TMA-store epilogues map 32-column pieces to output columns; the GLU epilogue stores 16 outputs per piece straight from
    // registers and only needs 16-column granularity
    if (p.n_tiles > 1 && p.block_n % (epi == EPI_BF16_GLU ? 16 : 32) != 0)
        return set_error(h, COCR_ERR_BAD_ARG, "block_n %d must be a multiple of 32 with several N tiles", p.block_n);
    p.x = res ? res->x : nullptr; p.alpha = res ? res->alpha : 1.f; p.init = res ? res->init : 0;
    p.gf = res ? res->gf : nullptr; p.bf = res ? res->bf : nullptr; p.g1 = res ? res->g1 : nullptr; p.b1 = res ? res->b1 : nullptr;
    if (epi == EPI_RESID && (!res || p.n_tiles != 1 || p.block_n != N || N != 8 * GEMM_RESID_NG || ldc % 4 != 0))
        return set_error(h, COCR_ERR_BAD_ARG, "residual epilogue is built for N = %d (got N=%d)", 8 * GEMM_RESID_NG, N);
    { static const int dbg = getenv("COCR_GEMM_DBG") ? atoi(getenv("COCR_GEMM_DBG")) : 0; p.dbg = dbg; }
    int tm = 32;
    while (tm < 2 * p.block_n) tm <<= 1;
    p.tmem_cols = tm;
    const uint32_t b_chunk = gemm_b_chunk_bytes(p.block_n);
    const uint32_t fixed = gemm_fixed_smem(epi);   // staging + bias / LayerNorm vectors + barriers + alignment slack
    const uint32_t budget = (uint32_t)h->bf->smem_optin;
    // resident weights when the whole K extent of one N-slice fits next to >= 4 A slots
    const uint32_t bres = (uint32_t)p.nkc * b_chunk;
    p.resident = (bres + 3 * GEMM_A_CHUNK + fixed <= budget) && (p.n_tiles <= h->num_sms) ? 1 : 0;
    const uint32_t slot = GEMM_A_CHUNK + (p.resident ? 0u : b_chunk);
    const uint32_t avail = budget - fixed - (p.resident ? bres : 0u);
    int stages = (int)(avail / slot);
    if (stages > 8) stages = 8;
    if (stages < 2) return set_error(h, COCR_ERR_UNSUPPORTED, "GEMM tile does not fit shared memory (N=%d K=%d)", N, K);
    p.stages = stages;
    const uint32_t smem = fixed + (p.resident ? bres : 0u) + (uint32_t)stages * slot;

    CUtensorMap tmA, tmB, tmC;
    COCR_TRY(make_tmap_bf16(h, &tmA, A, M, K, lda, GEMM_BM));
    COCR_TRY(make_tmap_bf16(h, &tmB, W, N, K, ldw, p.block_n));
    if (epi == EPI_BF16 || epi == EPI_BF16_SILU || epi == EPI_BF16_POS) {
        if (ldc % 8 != 0) return set_error(h, COCR_ERR_UNSUPPORTED, "bf16 GEMM output rows must be 16-byte aligned (ldc=%d)", ldc);
        COCR_TRY(make_tmap_bf16(h, &tmC, out, M, N, ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
    } else {
        tmC = tmA;   // unused
    }

    int grid;
    if (p.resident) {
        int groups = h->num_sms / p.n_tiles;
        if (groups > p.m_tiles) groups = p.m_tiles;
        if (groups < 1) groups = 1;
        grid = groups * p.n_tiles;
    } else {
        grid = p.m_tiles * p.n_tiles;
        if (grid > h->num_sms) grid = h->num_sms;
    }
    const double flops = 2.0 * M * (double)N * K;
    const double bytes = 2.0 * ((double)M * K + (double)N * K) + (epi == EPI_ARGMAX ? 8.0 * M : (epi == EPI_F32 ? 4.0 : 2.0) * (double)M * N);
    ProfScope ps(h, tag, st, flops, bytes);
    // COCR_GEMM_TRACE=<tag>:<file>: the 5th launch with that tag writes a (role, tag, clock) trace of CTA 0
    p.trace = nullptr;
    long long *trace_dev = nullptr;
    const char *trace_file = nullptr;
    {
        static int seen = 0;
        const char *spec = getenv("COCR_GEMM_TRACE");
        if (spec) {
            const char *colon = strchr(spec, ':');
            if (colon && strncmp(spec, tag, (size_t)(colon - spec)) == 0 && strlen(tag) == (size_t)(colon - spec) && ++seen == 5) {
                COCR_CUDA(h, cudaMalloc((void **)&trace_dev, (size_t)3 * GEMM_TRACE_EVENTS * 16));
                COCR_CUDA(h, cudaMemset(trace_dev, 0, (size_t)3 * GEMM_TRACE_EVENTS * 16));
                p.trace = trace_dev;
                trace_file = colon + 1;
            }
        }
    }
#define COCR_GEMM_CASE(E)                                                                                              \
    case E: {                                                                                                          \
        static PerDevice<int> attr_smem;                                                                               \
        if ((int)smem > attr_smem.get()) {                                                                             \
            COCR_CUDA(h, cudaFuncSetAttribute(k_gemm_tc<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget)); \
            attr_smem.get() = (int)budget;                                                                             \
        }                                                                                                              \
        k_gemm_tc<E><<<grid, gemm_threads(E), smem, st>>>(tmA, tmB, tmC, p);                                                  \
        break;                                                                                                         \
    }
    switch (epi) {
        COCR_GEMM_CASE(EPI_BF16)
        COCR_GEMM_CASE(EPI_BF16_SILU)
        COCR_GEMM_CASE(EPI_F32)
        COCR_GEMM_CASE(EPI_ARGMAX)
        COCR_GEMM_CASE(EPI_BF16_POS)
        COCR_GEMM_CASE(EPI_RESID)
        COCR_GEMM_CASE(EPI_BF16_GLU)
        default: return set_error(h, COCR_ERR_BAD_ARG, "unknown GEMM epilogue %d", epi);
    }
#undef COCR_GEMM_CASE
    COCR_LAUNCHED(h);
    if (trace_dev) {
        COCR_CUDA(h, cudaStreamSynchronize(st));
        std::vector<long long> hb((size_t)3 * GEMM_TRACE_EVENTS * 2);
        COCR_CUDA(h, cudaMemcpy(hb.data(), trace_dev, hb.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(trace_dev);
        if (FILE *fp = fopen(trace_file, "w")) {
            for (int role = 0; role < 3; ++role)
                for (int e = 0; e < GEMM_TRACE_EVENTS; ++e)
                    if (hb[((size_t)role * GEMM_TRACE_EVENTS + e) * 2]) fprintf(fp, "%d %lld %lld\n", role, hb[((size_t)role * GEMM_TRACE_EVENTS + e) * 2], hb[((size_t)role * GEMM_TRACE_EVENTS + e) * 2 + 1]);
            fclose(fp);
        }
    }
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// memory-bound row kernels of the bf16 path
// ------------------------------------------------------------------------------------------
__global__ void k_scale_bf16(bf16 *p, size_t n, float s) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = __float2bfloat16_rn(__bfloat162float(p[i]) * s);
}
__global__ void k_scale_f32(float *p, size_t n, float s) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] *= s;
}

__global__ void k_f32_to_bf16(const float *__restrict__ src, bf16 *__restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per frame.  x (fp32 residual stream) <- [x +] alpha*y ; optional block-final LayerNorm applied to x
// itself (encoder.py:99); then xn = bf16(LayerNorm(x; g1,b1)) for the next module (or bf16(x) when g1 == null).
//   init != 0 : x = y   (start of the stream, after the subsampling projection)
constexpr int ROW_MAXPER = 16;   // supports d <= 512
__global__ void __launch_bounds__(256)
k_resid_ln(float *__restrict__ x, const bf16 *__restrict__ y, float alpha, int init, const float *__restrict__ gf,
           const float *__restrict__ bfin, const float *__restrict__ g1, const float *__restrict__ b1,
           bf16 *__restrict__ xn, int M, int d) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float v[ROW_MAXPER];
    float *xr = x + (int64_t)row * d;
    const bf16 *yr = y + (int64_t)row * d;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_MAXPER; ++i) {
        int c = lane + 32 * i;
        float t = 0.f;
        if (c < d) {
            float yy = __bfloat162float(yr[c]);
            t = init ? yy : fmaf(alpha, yy, xr[c]);
        }
        v[i] = t;
        s += t;
    }
    const float inv_d = 1.0f / (float)d;
    if (gf) {   // block-final LayerNorm replaces the residual stream
        float mean = warp_sum_f(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < ROW_MAXPER; ++i) {
            int c = lane + 32 * i;
            float t = (c < d) ? v[i] - mean : 0.f;
            q = fmaf(t, t, q);
        }
        float rstd = rsqrtf(warp_sum_f(q) * inv_d + 1e-5f);
        s = 0.f;
#pragma unroll
        for (int i = 0; i < ROW_MAXPER; ++i) {
            int c = lane + 32 * i;
            if (c < d) {
                v[i] = (v[i] - mean) * rstd * gf[c] + bfin[c];
                s += v[i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < ROW_MAXPER; ++i) {
        int c = lane + 32 * i;
        if (c < d) xr[c] = v[i];
    }
    bf16 *nr = xn + (int64_t)row * d;
    if (g1) {
        float mean = warp_sum_f(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < ROW_MAXPER; ++i) {
            int c = lane + 32 * i;
            float t = (c < d) ? v[i] - mean : 0.f;
            q = fmaf(t, t, q);
        }
        float rstd = rsqrtf(warp_sum_f(q) * inv_d + 1e-5f);
#pragma unroll
        for (int i = 0; i < ROW_MAXPER; ++i) {
            int c = lane + 32 * i;
            if (c < d) nr[c] = __float2bfloat16_rn((v[i] - mean) * rstd * g1[c] + b1[c]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < ROW_MAXPER; ++i) {
            int c = lane + 32 * i;
            if (c < d) nr[c] = __float2bfloat16_rn(v[i]);
        }
    }
}

static int launch_resid_ln(cocr_handle *h, float *x, const bf16 *y, float alpha, int init, const float *gf, const float *bfin,
                           const float *g1, const float *b1, bf16 *xn, int M, int d, cudaStream_t st) {
    ProfScope ps(h, "resid_layernorm", st, 0.0, (double)M * d * (init ? 8.0 : 12.0));
    k_resid_ln<<<(M + 7) / 8, 256, 0, st>>>(x, y, alpha, init, gf, bfin, g1, b1, xn, M, d);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

// private quad-interleaved layout of the residual stream (GemmArgs::x) -> plain row-major rows
__global__ void k_resid_unpermute(const float *__restrict__ x, float *__restrict__ out, int64_t M, int d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one float4 of the private layout
    const int per_row = d / 4;
    if (i >= M * per_row) return;
    const int64_t row = i / per_row;
    const int s4 = (int)(i - row * per_row), gp = s4 >> 2, q = s4 & 3;
    const float4 t = *reinterpret_cast<const float4 *>(x + row * d + (int64_t)s4 * 4);
    float *o = out + row * d + 16 * gp + 2 * q;
    o[0] = t.x; o[1] = t.y; o[8] = t.z; o[9] = t.w;
}

__global__ void k_unpack_best(const unsigned long long *__restrict__ packed, int64_t M, float *__restrict__ val,
                              int32_t *__restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float v;
    int c;
    unpack_best(packed[i], &v, &c);
    val[i] = v;
    idx[i] = c;
}

// ------------------------------------------------------------------------------------------
// plan / forward
// ------------------------------------------------------------------------------------------
struct Bf16Plan {
    int B, W, T2, T, M;
    size_t off_a1, off_a2, off_feat, off_x, off_xn, off_y, off_h, off_g, off_best, off_best_val, off_best_idx, total;
};

// Row pitch (elements) of xn, the bf16 A operand of every K = d GEMM.  Dense by default.  COCR_XN_PAD=1 pads the pitch to
// a multiple of 64 elements so that every 128-byte TMA box row stays inside one cache line (with d = 144 three of four
// rows straddle two lines); measured on B200 it changes nothing (pw1 27.0 us either way), so the A stream is not bound
// by L2 request count.
static bool ffn_fused_ok(const cocr_handle *h);
static int xn_pitch(const cocr_handle *h) {
    static const bool pad = getenv("COCR_XN_PAD") != nullptr;
    if (ffn_fused_ok(h)) return FFN_K1;          // the fused FFN reads two columns of ones behind the features (ffn_tc.cuh)
    return (pad && h->dm.d == 8 * GEMM_RESID_NG) ? (h->dm.d + 63) / 64 * 64 : h->dm.d;
}

static bool resid_fused(const cocr_handle *h) { return h->dm.d == 8 * GEMM_RESID_NG && !getenv("COCR_SPLIT_RESIDUAL"); }

static bool use_tc_attn(const cocr_handle *h) { return h->bf && h->bf->tc_attn; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// the one-kernel tensor-core front-end (frontend_tc.cu) serves the default geometry and needs no intermediate in HBM
static bool frontend_tc_selected(const cocr_handle *h) {
    static const bool forced_off = getenv("COCR_FRONTEND_V2") != nullptr || getenv("COCR_FRONTEND_V1") != nullptr;
    const Dims &dm = h->dm;
    return !forced_off && dm.cc == 32 && dm.H == 96 && dm.F2 == 48 && dm.F == 24;
}

static void bf16_plan(const cocr_handle *h, int B, int W, Bf16Plan *p) {
    const Dims &dm = h->dm;
    p->B = B; p->W = W;
    p->T2 = out_len1(W);
    p->T = out_len1(p->T2);
    p->M = B * p->T;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    const size_t M = (size_t)p->M;
    // conv1 / depthwise outputs of the fallback front-ends (1.5 GB at B = 256, W = 1536); the tensor-core front-end keeps them on chip
    const bool fe_tc = frontend_tc_selected(h);
    p->off_a1 = take(fe_tc ? 0 : (size_t)B * p->T2 * dm.F2 * dm.cc * 4);
    p->off_a2 = take(fe_tc ? 0 : M * dm.F * dm.cc * 4);
    p->off_feat = take(M * dm.feat * 2);
    p->off_x = take(M * dm.d * 4);
    p->off_xn = take(M * (size_t)xn_pitch(h) * 2);
    p->off_y = take(M * dm.d * 2);
    size_t wide = (size_t)dm.ff > (size_t)3 * dm.d ? dm.ff : 3 * dm.d;
    if (use_tc_attn(h) && (size_t)h->bf->nx > wide) wide = h->bf->nx;
    if ((size_t)dm.feat > wide) wide = dm.feat;              // the front-end parks a2 (M*F, cc) here
    p->off_h = take(M * wide * 2);
    p->off_g = take(M * dm.d * 2);
    p->off_best = take(M * 8);
    p->off_best_val = take(M * 4);
    p->off_best_idx = take(M * 4);
    p->total = off;
}

size_t bf16_workspace_bytes(const cocr_handle *h, int B, int W) {
    Bf16Plan p;
    bf16_plan(h, B, W, &p);
    return p.total;
}


static uint16_t f2bf(float f) {   // round-to-nearest-even, host side
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// Fused projection for the tensor-core attention: per layer one (nx, d) weight producing
//   [Qe(h=0..nh-1) | Kx(h) | Vx(h)]  (layouts in attn_tc.cuh).  a~ = R(i)^T Wp_h^T (Wq_h x + bq_h + v_h) is folded into a
//   (d, d) merged weight Wa_h = Wp_h^T Wq_h (fp32 product, rounded once to bf16) and a bias Wp_h^T (bq_h + v_h).
static int pack_attention(cocr_handle *h) {
    Bf16State *bs = h->bf;
    const Dims &dm = h->dm;
    const int d = dm.d, nh = dm.nh, dh = dm.dh, L = dm.L;
    bs->tc_attn = (dh <= 40) && (d >= 80) && ((d - 16) % 64 == 0) && ((48 + d) / 64 <= ATT_MAXCHUNK) && L > 0;
    if (getenv("COCR_SIMT_ATTENTION")) bs->tc_attn = false;
    if (!bs->tc_attn) return COCR_OK;
    bs->qe_cols = 48 + d;
    bs->nchunk = bs->qe_cols / 64;
    bs->nx = nh * (bs->qe_cols + 64 + 64);
    const int nx = bs->nx, qc = bs->qe_cols;
    std::vector<uint16_t> W((size_t)L * nx * d, 0);
    std::vector<float> Bv((size_t)L * nx, 0.0f);
    std::vector<int16_t> meta(nx / 16, 0);
    const int tail_pair = (d - 16) / 2;
    for (int hd = 0; hd < nh; ++hd) {
        for (int g = 0; g < qc / 16; ++g) {
            int16_t m = 0;
            if (g == 3) m = (int16_t)((1 << 12) | tail_pair);
            else if (g >= 4) m = (int16_t)((1 << 12) | ((g - 4) * 8));
            meta[(hd * qc) / 16 + g] = m;
        }
        meta[(nh * qc + hd * 64) / 16 + 3] = (int16_t)((2 << 12) | tail_pair);
    }
    std::vector<float> wa((size_t)d * d);
    for (int l = 0; l < L; ++l) {
        const std::string q = "encoder.layers." + std::to_string(l) + ".sequential.1.module.attention.";
        const std::vector<float> &Wq = h->host_w.at(q + "query_proj.linear.weight").data, &bq = h->host_w.at(q + "query_proj.linear.bias").data;
        const std::vector<float> &Wk = h->host_w.at(q + "key_proj.linear.weight").data, &bk = h->host_w.at(q + "key_proj.linear.bias").data;
        const std::vector<float> &Wv = h->host_w.at(q + "value_proj.linear.weight").data, &bv = h->host_w.at(q + "value_proj.linear.bias").data;
        const std::vector<float> &Wp = h->host_w.at(q + "pos_proj.linear.weight").data;
        const std::vector<float> &ub = h->host_w.at(q + "u_bias").data, &vb = h->host_w.at(q + "v_bias").data;
        uint16_t *Wl = W.data() + (size_t)l * nx * d;
        float *Bl = Bv.data() + (size_t)l * nx;
        for (int hd = 0; hd < nh; ++hd) {
            // merged positional weight of this head: wa[m][:] = sum_c Wp[hd*dh+c][m] * Wq[hd*dh+c][:]
            std::vector<double> acc((size_t)d * d, 0.0), accb(d, 0.0);
            for (int c = 0; c < dh; ++c) {
                const float *wp = &Wp[(size_t)(hd * dh + c) * d];
                const float *wq = &Wq[(size_t)(hd * dh + c) * d];
                const double bb = (double)bq[hd * dh + c] + (double)vb[hd * dh + c];
                for (int m = 0; m < d; ++m) {
                    const double wpm = wp[m];
                    accb[m] += wpm * bb;
                    double *row = &acc[(size_t)m * d];
                    for (int k = 0; k < d; ++k) row[k] += wpm * (double)wq[k];
                }
            }
            const int qbase = hd * qc;
            for (int r = 0; r < dh; ++r) {
                for (int k = 0; k < d; ++k) Wl[(size_t)(qbase + r) * d + k] = f2bf(Wq[(size_t)(hd * dh + r) * d + k]);
                Bl[qbase + r] = bq[hd * dh + r] + ub[hd * dh + r];
            }
            for (int col = 48; col < qc; ++col) {
                const int m = (col < 64) ? (d - 16) + (col - 48) : (col - 64);
                for (int k = 0; k < d; ++k) Wl[(size_t)(qbase + col) * d + k] = f2bf((float)acc[(size_t)m * d + k]);
                Bl[qbase + col] = (float)accb[m];
            }
            const int kbase = nh * qc + hd * 64, vbase = nh * qc + nh * 64 + hd * 64;
            for (int r = 0; r < dh; ++r) {
                for (int k = 0; k < d; ++k) {
                    Wl[(size_t)(kbase + r) * d + k] = f2bf(Wk[(size_t)(hd * dh + r) * d + k]);
                    Wl[(size_t)(vbase + r) * d + k] = f2bf(Wv[(size_t)(hd * dh + r) * d + k]);
                }
                Bl[kbase + r] = bk[hd * dh + r];
                Bl[vbase + r] = bv[hd * dh + r];
            }
        }
    }
    // GEMM column order: a permutation of the 32-column pieces of the layout above.  A 256-column tile is drained by
    // two warps per lane quarter taking alternate pieces (even / odd).  For the default spec the pieces are arranged so
    // that (a) the rotated pieces a warp sees in a tile use the SAME 16 frequency pairs (angles evaluated once per
    // tile and warp instead of once per piece) and (b) every tile carries the same mix of rotated and plain pieces, so
    // the CTAs (one N tile each) finish together:
    //   tiles 0..3 (pair block j): even = [main j head 0, main j head 1, q+u head j, K head j (first half)]
    //                              odd  = [main j head 2, main j head 3, V piece 2j, V piece 2j+1]
    //   tile 4:                    even = K second halves (E-tail fill) of heads 0..3, odd = [q+u rest | a~ tail] of heads 0..3
    const int npieces = nx / 32;
    std::vector<int16_t> pdst(npieces);
    for (int i = 0; i < npieces; ++i) pdst[i] = (int16_t)i;
    if (qc == 192 && nh == 4 && !getenv("COCR_QKVX_NO_PERMUTE")) {
        for (int j = 0; j < 4; ++j) {
            int16_t *t = &pdst[8 * j];
            t[0] = (int16_t)(6 * 0 + 2 + j); t[2] = (int16_t)(6 * 1 + 2 + j); t[4] = (int16_t)(6 * j); t[6] = (int16_t)(24 + 2 * j);
            t[1] = (int16_t)(6 * 2 + 2 + j); t[3] = (int16_t)(6 * 3 + 2 + j); t[5] = (int16_t)(32 + 2 * j); t[7] = (int16_t)(33 + 2 * j);
        }
        for (int hd = 0; hd < 4; ++hd) {
            pdst[32 + 2 * hd] = (int16_t)(24 + 2 * hd + 1);
            pdst[32 + 2 * hd + 1] = (int16_t)(6 * hd + 1);
        }
        std::vector<int> seen(npieces, 0);
        for (int i = 0; i < npieces; ++i) seen[pdst[i]]++;
        for (int i = 0; i < npieces; ++i)
            if (seen[i] != 1) return set_error(h, COCR_ERR_STATE, "internal: QKVX column permutation is not a bijection");
    }
    {
        std::vector<uint16_t> W2(W.size());
        std::vector<float> B2(Bv.size());
        std::vector<int16_t> meta2(meta.size());
        for (int l = 0; l < L; ++l)
            for (int gp = 0; gp < npieces; ++gp) {
                const int dp = pdst[gp];
                memcpy(&W2[((size_t)l * nx + gp * 32) * d], &W[((size_t)l * nx + dp * 32) * d], (size_t)32 * d * 2);
                memcpy(&B2[(size_t)l * nx + gp * 32], &Bv[(size_t)l * nx + dp * 32], 32 * 4);
            }
        for (int gp = 0; gp < npieces; ++gp) { meta2[2 * gp] = meta[2 * pdst[gp]]; meta2[2 * gp + 1] = meta[2 * pdst[gp] + 1]; }
        W.swap(W2); Bv.swap(B2); meta.swap(meta2);
    }
    if (bs->w_qkvx) { cudaFree(bs->w_qkvx); bs->w_qkvx = nullptr; }
    if (bs->b_qkvx) { cudaFree(bs->b_qkvx); bs->b_qkvx = nullptr; }
    if (bs->meta) { cudaFree(bs->meta); bs->meta = nullptr; }
    if (bs->piece_dst) { cudaFree(bs->piece_dst); bs->piece_dst = nullptr; }
    void *p = nullptr;
    COCR_CUDA(h, cudaMalloc(&p, pdst.size() * 2)); bs->piece_dst = (int16_t *)p;
    COCR_CUDA(h, cudaMemcpy(bs->piece_dst, pdst.data(), pdst.size() * 2, cudaMemcpyHostToDevice));
    COCR_CUDA(h, cudaMalloc(&p, W.size() * 2)); bs->w_qkvx = (bf16 *)p;
    COCR_CUDA(h, cudaMalloc(&p, Bv.size() * 4)); bs->b_qkvx = (float *)p;
    COCR_CUDA(h, cudaMalloc(&p, meta.size() * 2)); bs->meta = (int16_t *)p;
    COCR_CUDA(h, cudaMemcpy(bs->w_qkvx, W.data(), W.size() * 2, cudaMemcpyHostToDevice));
    COCR_CUDA(h, cudaMemcpy(bs->b_qkvx, Bv.data(), Bv.size() * 4, cudaMemcpyHostToDevice));
    COCR_CUDA(h, cudaMemcpy(bs->meta, meta.data(), meta.size() * 2, cudaMemcpyHostToDevice));
    bs->tab_T = 0;
    return COCR_OK;
}

// omega_m and the shared key-side table E(t) for a batch whose lines have T frames (embedding.py:42-47 angles)
static int ensure_tables(cocr_handle *h, int T, cudaStream_t st) {
    Bf16State *bs = h->bf;
    // E(t) rows do not depend on the batch: the table only grows (a shorter batch reads a prefix through its tensor map),
    // so batches of mixed widths neither stall on a device synchronise nor invalidate captured graphs
    if (bs->tab_T >= T) return COCR_OK;
    if (bs->tab_T > 0 && T < 2 * bs->tab_T) T = std::min(2 * bs->tab_T, 8192);       // amortise regrowth
    const int d = h->dm.d, np = d / 2, ecols = d - 16;
    if (bs->etab) { COCR_CUDA(h, cudaDeviceSynchronize()); cudaFree(bs->omega); cudaFree(bs->etab); bs->omega = nullptr; bs->etab = nullptr; }
    std::vector<float> om(np);
    std::vector<uint16_t> et((size_t)T * ecols);
    const float neg = (float)(-(log(10000.0) / (double)d));
    for (int m = 0; m < np; ++m) om[m] = expf((float)(2 * m) * neg);
    for (int t = 0; t < T; ++t)
        for (int m = 0; m < np; ++m) {
            const float arg = (float)t * om[m];
            if (2 * m + 1 < ecols) {
                et[(size_t)t * ecols + 2 * m] = f2bf(cosf(arg));
                et[(size_t)t * ecols + 2 * m + 1] = f2bf(sinf(arg));
            }
        }
    void *p = nullptr;
    COCR_CUDA(h, cudaMalloc(&p, om.size() * 4)); bs->omega = (float *)p;
    COCR_CUDA(h, cudaMalloc(&p, et.size() * 2)); bs->etab = (bf16 *)p;
    COCR_CUDA(h, cudaMemcpyAsync(bs->omega, om.data(), om.size() * 4, cudaMemcpyHostToDevice, st));
    COCR_CUDA(h, cudaMemcpyAsync(bs->etab, et.data(), et.size() * 2, cudaMemcpyHostToDevice, st));
    COCR_CUDA(h, cudaStreamSynchronize(st));
    bs->tab_T = T;
    return COCR_OK;
}

static int launch_attn_tc(cocr_handle *h, const bf16 *qkvx, const int32_t *lens, bf16 *ctx, int B, int T, cudaStream_t st) {
    Bf16State *bs = h->bf;
    const Dims &dm = h->dm;
    const int64_t M = (int64_t)B * T;
    CUtensorMap tmQ, tmK, tmV, tmE;
    COCR_TRY(make_tmap_bf16(h, &tmQ, qkvx, M, dm.nh * bs->qe_cols, bs->nx, ATT_BQ));
    COCR_TRY(make_tmap_bf16(h, &tmK, qkvx + dm.nh * bs->qe_cols, M, dm.nh * 64, bs->nx, ATT_BK));
    COCR_TRY(make_tmap_bf16(h, &tmV, qkvx + dm.nh * bs->qe_cols + dm.nh * 64, M, dm.nh * 64, bs->nx, ATT_BK));
    COCR_TRY(make_tmap_bf16(h, &tmE, bs->etab, T, dm.d - 16, dm.d - 16, ATT_BK));
    AttnArgs a;
    a.B = B; a.T = T; a.nh = dm.nh; a.dh = dm.dh; a.d = dm.d; a.nchunk = bs->nchunk; a.lens = lens; a.ctx = ctx;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)dm.dh);
    { static const bool dv64 = getenv("COCR_ATT_DV64") != nullptr; a.dv_mma = (dv64 || dm.dh > 40) ? ATT_DV : ATT_DV_MMA; }
    a.nqt = (T + ATT_BQ - 1) / ATT_BQ;
    a.n_items = B * dm.nh * a.nqt;
    // E table resident in shared memory when it leaves room for the Q tile and two-deep K and V rings
    const uint32_t cap = (uint32_t)bs->smem_optin;
    a.e_resident = (a.nqt <= ATT_ERES_TILES && attn_smem_bytes(bs->nchunk, 1, 2, 2, T) <= cap && !getenv("COCR_ATTN_STREAM_E")) ? 1 : 0;
    // K/V-resident groups: every K and V tile of a (line, head) is read once instead of once per query tile
    a.kv_resident = (a.e_resident && a.nqt >= 2 && a.nqt <= ATT_MAXSTAGES && attn_smem_bytes(bs->nchunk, 1, a.nqt, a.nqt, T) <= cap &&
                     !getenv("COCR_ATTN_STREAM_KV")) ? 1 : 0;
    if (a.kv_resident) {
        a.k_stages = a.v_stages = a.nqt;
    } else {
        a.k_stages = 2; a.v_stages = 2;
        while (a.k_stages < ATT_MAXSTAGES && attn_smem_bytes(bs->nchunk, a.e_resident, a.k_stages + 1, a.v_stages, T) <= cap) ++a.k_stages;
        while (a.v_stages < ATT_MAXSTAGES && attn_smem_bytes(bs->nchunk, a.e_resident, a.k_stages, a.v_stages + 1, T) <= cap) ++a.v_stages;
    }
    const uint32_t smem = attn_smem_bytes(bs->nchunk, a.e_resident, a.k_stages, a.v_stages, T);
    if (smem > cap) return set_error(h, COCR_ERR_UNSUPPORTED, "attention tiles do not fit shared memory");
    static PerDevice<bool> attr;
    if (!attr.get()) {
        COCR_CUDA(h, cudaFuncSetAttribute(k_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs->smem_optin));
        attr.get() = true;
    }
    const int n_units = a.kv_resident ? B * dm.nh : a.n_items;              // what the CTAs deal out among themselves
    const int grid = n_units < h->num_sms ? n_units : h->num_sms;
    a.trace = nullptr;
    static int traced = 0;
    const char *trace_path = getenv("COCR_ATTN_TRACE");
    if (trace_path && traced < 40 && ++traced == 40) {      // the 40th launch (warm) writes a (role, tag, clock) trace of CTA 0
        const size_t n = (size_t)3 * ATT_TRACE_EVENTS * 2;
        long long *dbuf = nullptr;
        COCR_CUDA(h, cudaMalloc((void **)&dbuf, n * 8));
        COCR_CUDA(h, cudaMemset(dbuf, 0, n * 8));
        a.trace = dbuf;
        k_attn_tc<<<grid, ATT_THREADS, smem, st>>>(tmQ, tmK, tmV, tmE, a);
        COCR_CUDA(h, cudaStreamSynchronize(st));
        std::vector<long long> hb(n);
        COCR_CUDA(h, cudaMemcpy(hb.data(), dbuf, n * 8, cudaMemcpyDeviceToHost));
        cudaFree(dbuf);
        if (FILE *fp = fopen(trace_path, "w")) {
            for (int role = 0; role < 3; ++role)
                for (int e = 0; e < ATT_TRACE_EVENTS; ++e) {
                    const long long tag = hb[((size_t)role * ATT_TRACE_EVENTS + e) * 2], clk = hb[((size_t)role * ATT_TRACE_EVENTS + e) * 2 + 1];
                    if (tag) fprintf(fp, "%d %lld %lld\n", role, tag, clk);
                }
            fclose(fp);
        }
        a.trace = nullptr;
    }
    ProfScope ps(h, "attention_tc", st, 6.0 * B * T * (double)T * dm.d, 2.0 * M * (bs->nx + dm.d));
    k_attn_tc<<<grid, ATT_THREADS, smem, st>>>(tmQ, tmK, tmV, tmE, a);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// fused feed-forward module (ffn_tc.cuh)
// ------------------------------------------------------------------------------------------
static bool ffn_fused_ok(const cocr_handle *h) {
    const bool v1_ok = h->dm.ff % FFN_HC == 0 && h->dm.ff / FFN_HC <= FFN_MAX_NC;
    const bool v2_ok = h->dm.ff % FFN2_HC == 0 && h->dm.ff / FFN2_HC <= FFN2_MAX_NC && !getenv("COCR_FFN_V1");
    return h->dm.d == FFN_D && (v1_ok || v2_ok) && !getenv("COCR_FFN_SPLIT") && !getenv("COCR_SPLIT_RESIDUAL");
}

// second generation (ffn2_tc.cuh): xn as a TMEM operand, hidden chunks of 64; COCR_FFN_V1 selects the first one
static bool ffn_v2(int ff) {
    static const bool v1 = getenv("COCR_FFN_V1") != nullptr;
    return !v1 && ff % FFN2_HC == 0 && ff / FFN2_HC <= FFN2_MAX_NC;
}
static int ffn_hc(int ff) { return ffn_v2(ff) ? FFN2_HC : FFN_HC; }

// W1 (ff, d) bf16 + bias (ff) -> (ff, kp): [W1 row | bias hi | bias lo | 0 ...]
__global__ void k_pack_ffn1(const bf16 *w1, const float *b1, bf16 *out, int d, int ff, int kp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ff * kp) return;
    const int row = i / kp, k = i - row * kp;
    bf16 v = __float2bfloat16_rn(0.0f);
    if (k < d) v = w1[(size_t)row * d + k];
    else if (k == d) v = __float2bfloat16_rn(b1[row]);
    else if (k == d + 1) v = __float2bfloat16_rn(b1[row] - __bfloat162float(__float2bfloat16_rn(b1[row])));
    out[i] = v;
}
// pad columns of the xn buffer: (1, 1, 0, ...) behind the d features of every row
__global__ void k_xn_pad_init(bf16 *xn, int64_t M, int d, int ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int np = ld - d;
    if (i >= M * np) return;
    const int64_t row = i / np;
    const int k = (int)(i - row * np);
    xn[row * ld + d + k] = __float2bfloat16_rn(k < 2 ? 1.0f : 0.0f);
}

// W2 (d, ff) -> (ff/hc * d, hc): row c*d + n, column k  <-  W2[n][c*hc + k]
__global__ void k_pack_ffn2(const bf16 *w2, bf16 *out, int d, int ff, int hc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d * ff) return;
    const int row = i / hc, k = i - row * hc;
    const int c = row / d, n = row - c * d;
    out[i] = w2[(size_t)n * ff + c * hc + k];
}

// x <- x + alpha * FFN(xn) [-> LN_final]; xn_out = bf16(LN_next(x)).  xn: rows of 144 features followed by (1, 1, 0 x 14)
// (k_xn_pad_init), w1p: packed by k_pack_ffn1, w2p: packed by k_pack_ffn2.  xn_out may alias xn (a tile is written after it
// has been read, and only by its own CTA; the pad columns are never written).
static int launch_ffn_fused(cocr_handle *h, const bf16 *xn, int ldx, const bf16 *w1p, const bf16 *w2p, const float *b2,
                            float *x, bf16 *xn_out, int ldo, float alpha, const float *gf, const float *bfin, const float *g1, const float *b1n,
                            int M, int ff, cudaStream_t st, const char *tag) {
    if (M <= 0) return COCR_OK;
    if (ffn_v2(ff)) {
        if (ldx % 8 != 0 || ldx < FFN_K1 || ldo % 4 != 0) return set_error(h, COCR_ERR_UNSUPPORTED, "fused FFN needs aligned rows (ldx=%d ldo=%d)", ldx, ldo);
        if (FFN2_SMEM_BYTES > (uint32_t)h->bf->smem_optin) return set_error(h, COCR_ERR_UNSUPPORTED, "fused FFN does not fit shared memory");
        FfnArgs a;
        a.M = M; a.NC = ff / FFN2_HC; a.m_tiles = (M + 127) / 128;
        a.b2 = b2; a.x = x; a.xn_out = xn_out; a.ldo = ldo; a.alpha = alpha;
        a.gf = gf; a.bf = bfin; a.g1 = g1; a.b1n = b1n; a.trace = nullptr;
        { static const int dbg = getenv("COCR_FFN_DBG") ? atoi(getenv("COCR_FFN_DBG")) : 0; a.dbg = dbg; }
        CUtensorMap tmX, tmW1, tmW2;
        COCR_TRY(make_tmap_bf16(h, &tmX, xn, M, FFN_K1, ldx, 128));
        COCR_TRY(make_tmap_bf16(h, &tmW1, w1p, ff, FFN_K1, FFN_K1, FFN2_HC));
        COCR_TRY(make_tmap_bf16(h, &tmW2, w2p, (int64_t)a.NC * FFN_D, FFN2_HC, FFN2_HC, FFN_D));
        static PerDevice<bool> attr2;
        if (!attr2.get()) {
            COCR_CUDA(h, cudaFuncSetAttribute(k_ffn_fused2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FFN2_SMEM_BYTES));
            attr2.get() = true;
        }
        const int grid = a.m_tiles < h->num_sms ? a.m_tiles : h->num_sms;
        ProfScope ps(h, tag, st, 4.0 * M * (double)ff * FFN_D, (double)M * FFN_D * (2.0 + 2.0 + 4.0 + 4.0) + 4.0 * ff * FFN_D);
        k_ffn_fused2<<<grid, FFN_THREADS, FFN2_SMEM_BYTES, st>>>(tmX, tmW1, tmW2, a);
        COCR_LAUNCHED(h);
        return COCR_OK;
    }
    if (ff % FFN_HC != 0 || ff / FFN_HC > FFN_MAX_NC || ldx % 8 != 0 || ldx < FFN_K1 || ldo % 4 != 0)
        return set_error(h, COCR_ERR_UNSUPPORTED, "fused FFN needs ff = 96 * n (n <= %d) and aligned rows (ff=%d ldx=%d ldo=%d)", FFN_MAX_NC, ff, ldx, ldo);
    if (FFN_SMEM_BYTES > (uint32_t)h->bf->smem_optin) return set_error(h, COCR_ERR_UNSUPPORTED, "fused FFN does not fit shared memory");
    FfnArgs a;
    a.M = M; a.NC = ff / FFN_HC; a.m_tiles = (M + 127) / 128;
    a.b2 = b2; a.x = x; a.xn_out = xn_out; a.ldo = ldo; a.alpha = alpha;
    a.gf = gf; a.bf = bfin; a.g1 = g1; a.b1n = b1n;
    { static const int dbg = getenv("COCR_FFN_DBG") ? atoi(getenv("COCR_FFN_DBG")) : 0; a.dbg = dbg; }
    CUtensorMap tmX, tmW1, tmW2;
    COCR_TRY(make_tmap_bf16(h, &tmX, xn, M, FFN_K1, ldx, 128));
    COCR_TRY(make_tmap_bf16(h, &tmW1, w1p, ff, FFN_K1, FFN_K1, FFN_HC));
    COCR_TRY(make_tmap_bf16(h, &tmW2, w2p, (int64_t)a.NC * FFN_D, FFN_HC, FFN_HC, FFN_D));
    static PerDevice<bool> attr;
    if (!attr.get()) {
        COCR_CUDA(h, cudaFuncSetAttribute(k_ffn_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FFN_SMEM_BYTES));
        attr.get() = true;
    }
    const int grid = a.m_tiles < h->num_sms ? a.m_tiles : h->num_sms;
    ProfScope ps(h, tag, st, 4.0 * M * (double)ff * FFN_D, (double)M * FFN_D * (2.0 + 2.0 + 4.0 + 4.0) + 4.0 * ff * FFN_D);
    // COCR_FFN_TRACE=<file>: the 5th launch writes a (role, tag, clock) trace of CTA 0
    a.trace = nullptr;
    long long *trace_dev = nullptr;
    const char *trace_file = getenv("COCR_FFN_TRACE");
    {
        static int seen = 0;
        static const int at = getenv("COCR_FFN_TRACE_AT") ? atoi(getenv("COCR_FFN_TRACE_AT")) : 5;
        if (trace_file && ++seen == at) {
            COCR_CUDA(h, cudaMalloc((void **)&trace_dev, (size_t)3 * GEMM_TRACE_EVENTS * 16));
            COCR_CUDA(h, cudaMemset(trace_dev, 0, (size_t)3 * GEMM_TRACE_EVENTS * 16));
            a.trace = trace_dev;
        }
    }
    k_ffn_fused<<<grid, FFN_THREADS, FFN_SMEM_BYTES, st>>>(tmX, tmW1, tmW2, a);
    COCR_LAUNCHED(h);
    if (trace_dev) {
        COCR_CUDA(h, cudaStreamSynchronize(st));
        std::vector<long long> hb((size_t)3 * GEMM_TRACE_EVENTS * 2);
        COCR_CUDA(h, cudaMemcpy(hb.data(), trace_dev, hb.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(trace_dev);
        if (FILE *fp = fopen(trace_file, "w")) {
            for (int role = 0; role < 3; ++role)
                for (int e = 0; e < GEMM_TRACE_EVENTS; ++e)
                    if (hb[((size_t)role * GEMM_TRACE_EVENTS + e) * 2]) fprintf(fp, "%d %lld %lld\n", role, hb[((size_t)role * GEMM_TRACE_EVENTS + e) * 2], hb[((size_t)role * GEMM_TRACE_EVENTS + e) * 2 + 1]);
            fclose(fp);
        }
    }
    return COCR_OK;
}

// Which kernels these hyper-parameters select on the bf16 path, one "component: choice" per line; `fallbacks` counts the
// components that run on a slower general kernel because the fast one is built for another geometry.
int bf16_kernel_plan(const cocr_handle *h, char *buf, size_t cap, int *fallbacks) {
    const Dims &dm = h->dm;
    int nf = 0;
    std::string s;
    auto line = [&](const char *comp, bool fast, const char *a, const char *b) {
        s += comp; s += ": "; s += fast ? a : b; s += "\n";
        if (!fast) ++nf;
    };
    line("front-end", frontend_tc_selected(h), "k_frontend_tc (conv1 + depthwise + 1x1 conv on tcgen05, one kernel)",
         "k_frontend_convs2 / CUDA-core convs + tcgen05 1x1-conv GEMM (fast kernel needs height 96, 32 conv channels)");
    line("residual GEMMs", resid_fused(h), "tcgen05 GEMM with the residual update and LayerNorm(s) in the epilogue",
         "tcgen05 GEMM + separate residual/LayerNorm kernel (fused epilogue needs encoder_dim 144)");
    line("feed-forward module", ffn_fused_ok(h), ffn_v2(dm.ff) ? "k_ffn_fused2 (one kernel, hidden activation in TMEM)" : "k_ffn_fused (one kernel, hidden activation in TMEM)",
         "two tcgen05 GEMMs with the hidden activation in HBM (fused kernel needs encoder_dim 144 and ff = 64 n or 96 n)");
    line("attention", use_tc_attn(h), "k_attn_tc (tcgen05, rotation form of the relative-position term)",
         "CUDA-core attention with the projected positional table (tcgen05 kernel needs head dim <= 40 and encoder_dim = 16 + 64 n)");
    line("depthwise conv module", dm.d == 144 && dm.K == 31, "k_dw_v2 (compile-time geometry)", "general depthwise kernel (fast kernel needs encoder_dim 144, kernel size 31)");
    if (buf && cap) {
        const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    if (fallbacks) *fallbacks = nf;
    return COCR_OK;
}

int bf16_init(cocr_handle *h) {
    const Dims &dm = h->dm;
    if (dm.d % 16 != 0 || dm.ff % 16 != 0 || dm.feat % 16 != 0 || dm.d > 32 * ROW_MAXPER)
        return set_error(h, COCR_ERR_UNSUPPORTED,
                         "bf16 path needs encoder_dim, ff dim and subsampled feature dim to be multiples of 16 and encoder_dim <= %d "
                         "(got %d, %d, %d)", 32 * ROW_MAXPER, dm.d, dm.ff, dm.feat);
    if (!h->bf) h->bf = new Bf16State();
    Bf16State *bs = h->bf;
    for (LayerGraph &e : bs->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);   // captured weight pointers die with the arena
    bs->graphs.clear();
    if (bs->arena) { cudaFree(bs->arena); bs->arena = nullptr; }
    void *p = nullptr;
    COCR_CUDA(h, cudaMalloc(&p, h->arena_floats * sizeof(bf16)));
    bs->arena = (bf16 *)p;
    k_f32_to_bf16<<<1024, 256>>>(h->arena, bs->arena, h->arena_floats);
    COCR_LAUNCHED(h);
    COCR_CUDA(h, cudaDeviceSynchronize());
    // EPI_BF16_SILU evaluates silu(x) = h + h*tanh(h) from h = x/2: scale the FFN-up weights (bf16 mirror) and biases
    // (this handle's fp32 arena is not used by any other kernel in bf16 mode) by 1/2 - exact
    for (int l = 0; l < dm.L; ++l)
        for (int f = 0; f < 2; ++f) {
            const size_t nw = (size_t)dm.ff * dm.d;
            k_scale_bf16<<<(unsigned)((nw + 255) / 256), 256>>>(bs->arena + (h->layers[l].ffn_w1[f] - h->arena), nw, 0.5f);
            k_scale_f32<<<(dm.ff + 255) / 256, 256>>>(h->arena + (h->layers[l].ffn_b1[f] - h->arena), (size_t)dm.ff, 0.5f);
        }
    COCR_LAUNCHED(h);
    if (bs->w_ffn2p) { cudaFree(bs->w_ffn2p); bs->w_ffn2p = nullptr; }
    if (bs->w_ffn1p) { cudaFree(bs->w_ffn1p); bs->w_ffn1p = nullptr; }
    if (ffn_fused_ok(h) && dm.L > 0) {
        const size_t nw = (size_t)dm.ff * dm.d, nw1 = (size_t)dm.ff * FFN_K1;
        void *pw = nullptr;
        COCR_CUDA(h, cudaMalloc(&pw, (size_t)dm.L * 2 * nw * sizeof(bf16)));
        bs->w_ffn2p = (bf16 *)pw;
        COCR_CUDA(h, cudaMalloc(&pw, (size_t)dm.L * 2 * nw1 * sizeof(bf16)));
        bs->w_ffn1p = (bf16 *)pw;
        for (int l = 0; l < dm.L; ++l)
            for (int f = 0; f < 2; ++f) {
                k_pack_ffn2<<<(unsigned)((nw + 255) / 256), 256>>>(bs->arena + (h->layers[l].ffn_w2[f] - h->arena), bs->w_ffn2p + ((size_t)l * 2 + f) * nw, dm.d, dm.ff, ffn_hc(dm.ff));
                k_pack_ffn1<<<(unsigned)((nw1 + 255) / 256), 256>>>(bs->arena + (h->layers[l].ffn_w1[f] - h->arena), h->layers[l].ffn_b1[f],
                                                                     bs->w_ffn1p + ((size_t)l * 2 + f) * nw1, dm.d, dm.ff, FFN_K1);
            }
        COCR_LAUNCHED(h);
    }
    {
        // pw1 for the GLU epilogue (convolution.py:138-139): GEMM column 2c = value channel c, 2c+1 = gate channel d + c
        const int d = dm.d;
        std::vector<uint16_t> W((size_t)dm.L * 2 * d * d);
        std::vector<float> Bv((size_t)dm.L * 2 * d);
        for (int l = 0; l < dm.L; ++l) {
            const std::string q = "encoder.layers." + std::to_string(l) + ".sequential.2.module.sequential.2.conv.";
            const std::vector<float> &w = h->host_w.at(q + "weight").data, &b = h->host_w.at(q + "bias").data;
            for (int c = 0; c < d; ++c)
                for (int half = 0; half < 2; ++half) {
                    const int src = half * d + c, dst = 2 * c + half;
                    for (int k = 0; k < d; ++k) W[((size_t)l * 2 * d + dst) * d + k] = f2bf(0.5f * w[(size_t)src * d + k]);
                    Bv[(size_t)l * 2 * d + dst] = 0.5f * b[src];
                }
        }
        // depthwise taps transposed to [k][channel] with half the BatchNorm scale folded in; half shift behind them
        const int Kt = dm.K;
        std::vector<float> taps((size_t)dm.L * (Kt + 1) * d);
        for (int l = 0; l < dm.L; ++l) {
            const std::string q = "encoder.layers." + std::to_string(l) + ".sequential.2.module.sequential.";
            const std::vector<float> &w = h->host_w.at(q + "4.conv.weight").data;
            const std::vector<float> &gm = h->host_w.at(q + "5.weight").data, &bt = h->host_w.at(q + "5.bias").data;
            const std::vector<float> &mu = h->host_w.at(q + "5.running_mean").data, &var = h->host_w.at(q + "5.running_var").data;
            float *dst = taps.data() + (size_t)l * (Kt + 1) * d;
            for (int c = 0; c < d; ++c) {
                const float sc = gm[c] * (1.0f / sqrtf(var[c] + 1e-5f)), sh = bt[c] - mu[c] * sc;
                for (int k = 0; k < Kt; ++k) dst[(size_t)k * d + c] = 0.5f * sc * w[(size_t)c * Kt + k];
                dst[(size_t)Kt * d + c] = 0.5f * sh;
            }
        }
        if (bs->dw_taps) { cudaFree(bs->dw_taps); bs->dw_taps = nullptr; }
        void *pt = nullptr;
        COCR_CUDA(h, cudaMalloc(&pt, taps.size() * 4)); bs->dw_taps = (float *)pt;
        COCR_CUDA(h, cudaMemcpy(bs->dw_taps, taps.data(), taps.size() * 4, cudaMemcpyHostToDevice));
        {
            const int cc = dm.cc;
            const std::vector<float> &wp = h->host_w.at("encoder.conv_subsample.conv.3.weight").data;
            std::vector<uint16_t> w2((size_t)cc * 2 * cc);
            for (int n = 0; n < cc; ++n)
                for (int k = 0; k < cc; ++k) w2[(size_t)n * 2 * cc + k] = w2[(size_t)n * 2 * cc + cc + k] = f2bf(wp[(size_t)n * cc + k]);
            if (bs->w_fe_pw2) { cudaFree(bs->w_fe_pw2); bs->w_fe_pw2 = nullptr; }
            void *pw = nullptr;
            COCR_CUDA(h, cudaMalloc(&pw, w2.size() * 2)); bs->w_fe_pw2 = (bf16 *)pw;
            COCR_CUDA(h, cudaMemcpy(bs->w_fe_pw2, w2.data(), w2.size() * 2, cudaMemcpyHostToDevice));
        }
        if (bs->w_pw1g) { cudaFree(bs->w_pw1g); bs->w_pw1g = nullptr; }
        if (bs->b_pw1g) { cudaFree(bs->b_pw1g); bs->b_pw1g = nullptr; }
        void *pp = nullptr;
        COCR_CUDA(h, cudaMalloc(&pp, W.size() * 2)); bs->w_pw1g = (bf16 *)pp;
        COCR_CUDA(h, cudaMalloc(&pp, Bv.size() * 4)); bs->b_pw1g = (float *)pp;
        COCR_CUDA(h, cudaMemcpy(bs->w_pw1g, W.data(), W.size() * 2, cudaMemcpyHostToDevice));
        COCR_CUDA(h, cudaMemcpy(bs->b_pw1g, Bv.data(), Bv.size() * 4, cudaMemcpyHostToDevice));
    }
    COCR_CUDA(h, cudaDeviceSynchronize());
    COCR_CUDA(h, cudaDeviceGetAttribute(&bs->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    if (!get_encode()) return set_error(h, COCR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    return pack_attention(h);
}

void bf16_destroy(cocr_handle *h) {
    if (!h->bf) return;
    Bf16State *bs = h->bf;
    if (bs->arena) cudaFree(bs->arena);
    if (bs->w_qkvx) cudaFree(bs->w_qkvx);
    if (bs->b_qkvx) cudaFree(bs->b_qkvx);
    if (bs->meta) cudaFree(bs->meta);
    if (bs->piece_dst) cudaFree(bs->piece_dst);
    if (bs->omega) cudaFree(bs->omega);
    if (bs->etab) cudaFree(bs->etab);
    if (bs->w_pw1g) cudaFree(bs->w_pw1g);
    if (bs->b_pw1g) cudaFree(bs->b_pw1g);
    if (bs->dw_taps) cudaFree(bs->dw_taps);
    if (bs->w_fe_pw2) cudaFree(bs->w_fe_pw2);
    if (bs->w_ffn2p) cudaFree(bs->w_ffn2p);
    if (bs->w_ffn1p) cudaFree(bs->w_ffn1p);
    for (LayerGraph &e : bs->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
    if (bs->g_in) cudaEventDestroy(bs->g_in);
    if (bs->g_out) cudaEventDestroy(bs->g_out);
    if (bs->gstream) cudaStreamDestroy(bs->gstream);
    delete h->bf;
    h->bf = nullptr;
}

static inline const bf16 *WB(const cocr_handle *h, const float *w) { return h->bf->arena + (w - h->arena); }

// L2 residency of the residual stream (COCR_L2_PERSIST=<MB to keep>, experiment): x (fp32) and xn (bf16) are read and written by
// every kernel of a layer while the wide intermediates (QKVX operands, 252 MB) stream through once; an access-policy window
// marks [x | xn] as persisting so that the streams do not evict them.
static int set_l2_window(cocr_handle *h, cudaStream_t st, void *base, size_t bytes) {
    static const int want_mb = getenv("COCR_L2_PERSIST") ? atoi(getenv("COCR_L2_PERSIST")) : 0;
    if (want_mb <= 0) return COCR_OK;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
    if (max_persist <= 0 || max_window <= 0) return COCR_OK;
    size_t persist = std::min((size_t)want_mb << 20, (size_t)max_persist);
    static PerDevice<size_t> limit;
    if (limit.get() != persist) {
        COCR_CUDA(h, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
        limit.get() = persist;
        fprintf(stderr, "cocr: persisting L2 %zu MB (device max %d MB, window max %d MB)\n", persist >> 20, max_persist >> 20, max_window >> 20);
    }
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = base;
    v.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)max_window);
    v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)v.accessPolicyWindow.num_bytes);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    COCR_CUDA(h, cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v));
    return COCR_OK;
}

// encoder forward; leaves x (fp32) and xn = bf16(x) in the workspace
static int bf16_forward(cocr_handle *h, const void *img_any, bool u8, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int B,
                        int W, cudaStream_t st, int32_t *out_lens, Bf16Plan &p) {
    const Dims &dm = h->dm;
    bf16_plan(h, B, W, &p);
    COCR_TRY(ensure_workspace(h, p.total));
    const bool tc_attn = use_tc_attn(h);
    if (tc_attn) COCR_TRY(ensure_tables(h, p.T, st));
    else COCR_TRY(ensure_pos_proj(h, p.T, st));
    char *ws = h->ws.base;
    const float *img = static_cast<const float *>(img_any);        // only dereferenced as fp32 when !u8
    float *a1 = (float *)(ws + p.off_a1), *a2 = (float *)(ws + p.off_a2);
    bf16 *feat = (bf16 *)(ws + p.off_feat);
    float *x = (float *)(ws + p.off_x);
    bf16 *xn = (bf16 *)(ws + p.off_xn), *y = (bf16 *)(ws + p.off_y), *hb = (bf16 *)(ws + p.off_h), *g = (bf16 *)(ws + p.off_g);
    const int M = p.M, d = dm.d;
    const int ldx = xn_pitch(h);
    if (ldx > d && ffn_fused_ok(h)) {           // ones columns behind the features (bias of the fused FFN's first GEMM)
        const int64_t n = (int64_t)M * (ldx - d);
        k_xn_pad_init<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(xn, M, d, ldx);
        COCR_LAUNCHED(h);
    }

    if (out_lens) COCR_TRY(launch_out_lens(h, widths, B, W, out_lens, st));
    const int32_t *lens_dev = nullptr;
    if (h->hp.mask_mode == COCR_MASK_RAGGED && widths) {
        int32_t *tmp = (int32_t *)(ws + p.off_best_idx);
        COCR_TRY(launch_out_lens(h, widths, B, W, tmp, st));
        lens_dev = tmp;
    }
    // front-end (convolution.py:227-237): direct convs on CUDA cores, 1x1 conv as an fp32 GEMM writing bf16 features,
    // 768->144 projection on tensor cores
    // conv1 + depthwise conv on CUDA cores -> a2 (M*F, [hi | lo] 2x32) bf16, then the 1x1 conv + ReLU as a K = 64
    // tensor-core GEMM whose (M*F, 32) output IS the (M, F*32) feature matrix of the 768 -> 144 projection
    bf16 *a2hl = (bf16 *)a2;                                       // (M*F, 2 cc) bf16 = the fp32 a2 region of the plan
    // third generation (frontend_tc.cu): conv1 and the 1x1 conv on tensor cores inside ONE kernel, image in -> features out
    int rc_fe = !frontend_tc_selected(h) ? COCR_ERR_UNSUPPORTED : launch_frontend_tc(h, img_any, u8, sb, sw, sh, B, W, h->bf->w_fe_pw2, feat, st);
    if (frontend_tc_selected(h) && rc_fe == COCR_ERR_UNSUPPORTED)
        return set_error(h, COCR_ERR_STATE, "internal: the tensor-core front-end rejected a geometry the plan selected it for");
    const bool fe_done = rc_fe == COCR_OK;
    if (rc_fe == COCR_ERR_UNSUPPORTED)
        rc_fe = getenv("COCR_FRONTEND_V1") ? COCR_ERR_UNSUPPORTED : launch_frontend_convs2(h, img_any, u8, sb, sw, sh, B, W, a2hl, st);
    if (rc_fe == COCR_ERR_UNSUPPORTED && u8)
        return set_error(h, COCR_ERR_UNSUPPORTED, "uint8 input is implemented by the fused front-ends only (subsampling_conv_channels = 32)");
    if (fe_done) {
        // features are complete
    } else if (rc_fe == COCR_OK) {
        rc_fe = launch_gemm_tc(h, EPI_BF16, a2hl, 2 * dm.cc, h->bf->w_fe_pw2, 2 * dm.cc, h->sub_bpw, feat, dm.cc, nullptr, M * dm.F, dm.cc,
                               2 * dm.cc, st, 0, nullptr, "gemm_frontend_pw", nullptr, /*relu=*/1);
    } else if (rc_fe == COCR_ERR_UNSUPPORTED) {
        rc_fe = launch_frontend_fused(h, img, sb, sw, sh, B, W, feat, st);
    }
    if (rc_fe == COCR_ERR_UNSUPPORTED) {
        COCR_TRY(launch_frontend_convs(h, img, sb, sw, sh, B, W, a1, a2, st));
        COCR_TRY(launch_sgemm(h, a2, dm.cc, h->sub_wpw, dm.cc, h->sub_bpw, (float *)feat, dm.cc, M * dm.F, dm.cc, dm.cc, 2, nullptr,
                              0, 0.f, st, /*out_bf16=*/1));
    } else if (rc_fe != COCR_OK) {
        return rc_fe;
    }
    // Residual-stream GEMMs: with encoder_dim 144 the update x += alpha*(..) and the LayerNorm feeding the next module
    // run in the GEMM epilogue (EPI_RESID); otherwise GEMM -> bf16 y -> k_resid_ln.
    const bool fused = resid_fused(h);
    // rows [r0, r0 + rows) of the residual stream; A points at the first of those rows
    auto resid_rows = [&](const bf16 *A, int K, const float *Wf, const float *bias, float alpha, int init, const float *gf,
                          const float *bfin, const float *g1, const float *b1, const char *tag, int r0, int rows) -> int {
        float *xr = x + (size_t)r0 * d;
        bf16 *xnr = xn + (size_t)r0 * ldx;
        if (fused) {
            ResidArgs ra = {xr, alpha, init, gf, bfin, g1, b1};
            return launch_gemm_tc(h, EPI_RESID, A, K, WB(h, Wf), K, bias, xnr, ldx, nullptr, rows, d, K, st, 0, nullptr, tag, &ra);
        }
        bf16 *yr = y + (size_t)r0 * d;
        COCR_TRY(launch_gemm_tc(h, EPI_BF16, A, K, WB(h, Wf), K, bias, yr, d, nullptr, rows, d, K, st, 0, nullptr, tag));
        return launch_resid_ln(h, xr, yr, alpha, init, gf, bfin, g1, b1, xnr, rows, d, st);
    };
    auto resid_gemm = [&](const bf16 *A, int K, const float *Wf, const float *bias, float alpha, int init, const float *gf,
                          const float *bfin, const float *g1, const float *b1, const char *tag) -> int {
        return resid_rows(A, K, Wf, bias, alpha, init, gf, bfin, g1, b1, tag, 0, M);
    };
    // Optional L2-residency experiment (COCR_FFN_CHUNKS / COCR_ATT_CHUNKS = number of chunks): producer and consumer of
    // the wide intermediates (FFN hidden M x 4d, attention operands M x nx) run on chunks of whole lines that fit L2.
    // Default is one chunk: measured slower (12.1 vs 10.7 ms/step at 2/4 chunks), the kernels are not HBM-bound.
    auto chunk_lines = [&](size_t bytes_per_row, const char *env) {
        int n = 1;
        if (const char *e = getenv(env)) n = atoi(e);
        (void)bytes_per_row;   // measured on B200: every extra launch costs 5-10 us of ramp-up/tail, more than the L2 hits return
        if (n < 1) n = 1;
        return (B + n - 1) / n;                      // lines per chunk
    };
    const int ffn_lines = chunk_lines((size_t)dm.ff * 2, "COCR_FFN_CHUNKS");
    auto ffn_module = [&](int l, int f, const float *gf, const float *bfin, const float *g1, const float *b1, const char *tag_down) -> int {
        const LayerF32 &lw = h->layers[l];
        if (fused && h->bf->w_ffn2p)      // one kernel, the hidden activation stays on the SM
            return launch_ffn_fused(h, xn, ldx, h->bf->w_ffn1p + ((size_t)l * 2 + f) * dm.ff * FFN_K1, h->bf->w_ffn2p + ((size_t)l * 2 + f) * dm.ff * d, lw.ffn_b2[f], x, xn,
                                    ldx, dm.ffn_factor, gf, bfin, g1, b1, M, dm.ff, st, f ? "ffn_fused_final" : "ffn_fused");
        for (int b0 = 0; b0 < B; b0 += ffn_lines) {
            const int nb = std::min(ffn_lines, B - b0), r0 = b0 * p.T, rows = nb * p.T;
            COCR_TRY(launch_gemm_tc(h, EPI_BF16_SILU, xn + (size_t)r0 * ldx, ldx, WB(h, lw.ffn_w1[f]), d, lw.ffn_b1[f], hb, dm.ff, nullptr, rows,
                                    dm.ff, d, st, 0, nullptr, "gemm_ffn_up"));
            COCR_TRY(resid_rows(hb, dm.ff, lw.ffn_w2[f], lw.ffn_b2[f], dm.ffn_factor, 0, gf, bfin, g1, b1, tag_down, r0, rows));
        }
        return COCR_OK;
    };
    const float *g0 = dm.L > 0 ? h->layers[0].ffn_ln_g[0] : nullptr, *b0 = dm.L > 0 ? h->layers[0].ffn_ln_b[0] : nullptr;
    COCR_TRY(resid_gemm(feat, dm.feat, h->sub_wout, h->sub_bout, 1.f, 1, nullptr, nullptr, g0, b0, "gemm_sub_out"));
    // the encoder layers: ~11 launches per layer that touch nothing but the workspace and the packed weights, so the
    // sequence is replayed as a CUDA graph (see below); every lambda above launches on `st` by reference
    auto run_layers = [&]() -> int {
    for (int l = 0; l < dm.L; ++l) {
        const LayerF32 &lw = h->layers[l];
        // FFN 1 (feed_forward.py:45-55; half-step residual encoder.py:68-75) -> LayerNorm of the attention module
        COCR_TRY(ffn_module(l, 0, nullptr, nullptr, lw.att_ln_g, lw.att_ln_b, "gemm_ffn_down"));
        // MHSA (attention.py:72-151)
        if (tc_attn) {
            Bf16State *bs = h->bf;
            PosArgs pa = {bs->meta, bs->omega, d / 2, p.T, bs->piece_dst};
            const int att_lines = chunk_lines((size_t)bs->nx * 2, "COCR_ATT_CHUNKS");
            for (int b0 = 0; b0 < B; b0 += att_lines) {
                const int nb = std::min(att_lines, B - b0), r0 = b0 * p.T, rows = nb * p.T;
                COCR_TRY(launch_gemm_tc(h, EPI_BF16_POS, xn + (size_t)r0 * ldx, ldx, bs->w_qkvx + (size_t)l * bs->nx * d, d,
                                        bs->b_qkvx + (size_t)l * bs->nx, hb, bs->nx, nullptr, rows, bs->nx, d, st, 0, &pa, "gemm_qkvx_pos"));
                COCR_TRY(launch_attn_tc(h, hb, lens_dev ? lens_dev + b0 : nullptr, g + (size_t)r0 * d, nb, p.T, st));
            }
        } else {
            COCR_TRY(launch_gemm_tc(h, EPI_BF16, xn, ldx, WB(h, lw.w_qkv), d, lw.b_qkv, hb, 3 * d, nullptr, M, 3 * d, d, st, 0, nullptr, "gemm_qkv"));
            COCR_TRY(launch_attn_simt_bf16(h, hb, h->pos_proj + (size_t)l * (2 * p.T - 1) * d, lw.u_bias, lw.v_bias, lens_dev, g,
                                           B, p.T, st));
        }
        COCR_TRY(resid_gemm(g, d, lw.w_out, lw.b_out, 1.f, 0, nullptr, nullptr, lw.conv_ln_g, lw.conv_ln_b, "gemm_attn_out"));
        // convolution module (convolution.py:135-148)
        int rc_dw = COCR_ERR_UNSUPPORTED;
        if (d % 16 == 0) {
            // pointwise conv 1 with the GLU in its epilogue (value, gate interleaved) -> depthwise conv + BatchNorm + SiLU
            Bf16State *bs = h->bf;
            // two N tiles of 144 (value / gate pairs of 72 channels each) instead of three of 96: A is read twice, not three
            // times, and a CTA walks 10.4 instead of 15.7 M tiles (COCR_PW1_BN overrides)
            // (chosen per handle: a function-local static would freeze the choice of the first model the process ran)
            static const int pw1_bn_env = getenv("COCR_PW1_BN") ? atoi(getenv("COCR_PW1_BN")) : -1;
            const int pw1_bn = pw1_bn_env >= 0 ? pw1_bn_env : (d == 144 ? 144 : 0);
            COCR_TRY(launch_gemm_tc(h, EPI_BF16_GLU, xn, ldx, bs->w_pw1g + (size_t)l * 2 * d * d, d, bs->b_pw1g + (size_t)l * 2 * d, hb, d,
                                    nullptr, M, 2 * d, d, st, pw1_bn, nullptr, "gemm_pw1_glu"));
            const float *taps = bs->dw_taps + (size_t)l * (dm.K + 1) * d;
            rc_dw = launch_dw_fast(h, hb, taps, taps + (size_t)dm.K * d, g, B, p.T, st);
        }
        if (rc_dw == COCR_ERR_UNSUPPORTED) {
            COCR_TRY(launch_gemm_tc(h, EPI_BF16, xn, ldx, WB(h, lw.w_pw1), d, lw.b_pw1, hb, 2 * d, nullptr, M, 2 * d, d, st, 0, nullptr, "gemm_pw1"));
            rc_dw = launch_glu_dw_fast(h, hb, lw, g, B, p.T, st);
            if (rc_dw == COCR_ERR_UNSUPPORTED) rc_dw = launch_glu_dw_bf16(h, hb, lw, g, B, p.T, st);
        }
        if (rc_dw != COCR_OK) return rc_dw;
        COCR_TRY(resid_gemm(g, d, lw.w_pw2, lw.b_pw2, 1.f, 0, nullptr, nullptr, lw.ffn_ln_g[1], lw.ffn_ln_b[1], "gemm_pw2"));
        // FFN 2 + block-final LayerNorm (encoder.py:91-99) + next block's first LayerNorm
        const float *ng = (l + 1 < dm.L) ? h->layers[l + 1].ffn_ln_g[0] : nullptr;
        const float *nb = (l + 1 < dm.L) ? h->layers[l + 1].ffn_ln_b[0] : nullptr;
        COCR_TRY(ffn_module(l, 1, lw.final_ln_g, lw.final_ln_b, ng, nb, "gemm_ffn_down_final"));
    }
    return COCR_OK;
    };

    // CUDA graph of the layer sequence, keyed by everything its kernel arguments depend on.  First call with a key runs
    // eagerly (it also settles the one-time function attributes), the second captures, later ones replay.  The graph
    // runs on the handle's own stream, fenced against the caller's stream with two events, because the caller's stream
    // may be the legacy default stream, which cannot be captured.  Measured: 7.49 -> 7.19 ms per step (launch gaps).
    Bf16State *bs = h->bf;
    if (!bs->gstream) {
        COCR_CUDA(h, cudaStreamCreateWithFlags(&bs->gstream, cudaStreamNonBlocking));
        COCR_CUDA(h, cudaEventCreateWithFlags(&bs->g_in, cudaEventDisableTiming));
        COCR_CUDA(h, cudaEventCreateWithFlags(&bs->g_out, cudaEventDisableTiming));
    }
    {
        const size_t span = (size_t)((char *)xn - (char *)x) + (size_t)M * ldx * 2;      // x and xn are neighbours in the plan
        COCR_TRY(set_l2_window(h, st, x, span));
        COCR_TRY(set_l2_window(h, bs->gstream, x, span));
    }
    static const bool no_graph = getenv("COCR_NO_GRAPH") != nullptr || getenv("COCR_FFN_TRACE") || getenv("COCR_ATTN_TRACE") || getenv("COCR_GEMM_TRACE");
    bool use_graph = !no_graph && !h->profiling && dm.L > 0;
    if (use_graph) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); use_graph = false; }
    }
    if (!use_graph) return run_layers();
    LayerGraph *lg = nullptr;
    for (LayerGraph &e : bs->graphs)
        if (e.B == B && e.T == p.T && e.T2 == p.T2 && e.lens == (const void *)lens_dev && e.ws == ws && e.etab == (const void *)bs->etab &&
            e.pos == (const void *)h->pos_proj) lg = &e;
    if (lg) lg->last_use = ++bs->graph_clock;
    if (!lg) {
        if (bs->graphs.size() >= 16) {                        // shapes keep changing: evict the least recently used graph
            size_t victim = 0;
            for (size_t i = 1; i < bs->graphs.size(); ++i)
                if (bs->graphs[i].last_use < bs->graphs[victim].last_use) victim = i;
            if (bs->graphs[victim].exec) cudaGraphExecDestroy(bs->graphs[victim].exec);
            bs->graphs.erase(bs->graphs.begin() + (long)victim);
        }
        bs->graphs.push_back({B, p.T, p.T2, (const void *)lens_dev, ws, (const void *)bs->etab, (const void *)h->pos_proj, nullptr, 0, 1, ++bs->graph_clock});
        return run_layers();
    }
    if (lg->state < 0) return run_layers();
    if (!bs->gstream) {
        COCR_CUDA(h, cudaStreamCreateWithFlags(&bs->gstream, cudaStreamNonBlocking));
        COCR_CUDA(h, cudaEventCreateWithFlags(&bs->g_in, cudaEventDisableTiming));
        COCR_CUDA(h, cudaEventCreateWithFlags(&bs->g_out, cudaEventDisableTiming));
    }
    if (lg->state == 1) {
        const int64_t launches0 = h->launches;
        cudaStream_t caller = st;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(bs->gstream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        int rc = COCR_OK;
        if (ok) {
            st = bs->gstream;
            rc = run_layers();
            st = caller;
            ok = cudaStreamEndCapture(bs->gstream, &graph) == cudaSuccess && rc == COCR_OK && graph;
        }
        if (ok) ok = cudaGraphInstantiate(&lg->exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        lg->launches = h->launches - launches0;
        h->launches = launches0;                             // nothing has run yet
        if (!ok) {
            cudaGetLastError();
            h->err.clear();
            lg->state = -1;
            lg->exec = nullptr;
            return run_layers();
        }
        lg->state = 2;
    }
    COCR_CUDA(h, cudaEventRecord(bs->g_in, st));
    COCR_CUDA(h, cudaStreamWaitEvent(bs->gstream, bs->g_in, 0));
    COCR_CUDA(h, cudaGraphLaunch(lg->exec, bs->gstream));
    h->launches += lg->launches;
    COCR_CUDA(h, cudaEventRecord(bs->g_out, bs->gstream));
    COCR_CUDA(h, cudaStreamWaitEvent(st, bs->g_out, 0));
    return COCR_OK;
}

int bf16_encode(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int B, int W,
                cudaStream_t st, float *enc_out, int32_t *out_lens) {
    Bf16Plan p;
    COCR_TRY(bf16_forward(h, img, false, sb, sw, sh, widths, B, W, st, out_lens, p));
    if (resid_fused(h)) {       // the fused residual epilogues keep x in their private layout
        const int64_t n4 = (int64_t)p.M * (h->dm.d / 4);
        k_resid_unpermute<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>((const float *)(h->ws.base + p.off_x), enc_out, p.M, h->dm.d);
        COCR_LAUNCHED(h);
        return COCR_OK;
    }
    COCR_CUDA(h, cudaMemcpyAsync(enc_out, h->ws.base + p.off_x, (size_t)p.M * h->dm.d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return COCR_OK;
}

int bf16_logits(cocr_handle *h, const float *img, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int B, int W,
                cudaStream_t st, float *logits, int32_t *out_lens) {
    Bf16Plan p;
    COCR_TRY(bf16_forward(h, img, false, sb, sw, sh, widths, B, W, st, out_lens, p));
    const Dims &dm = h->dm;
    return launch_gemm_tc(h, EPI_F32, (const bf16 *)(h->ws.base + p.off_xn), xn_pitch(h), WB(h, h->dec_w), dm.d, h->dec_b, logits, dm.C,
                          nullptr, p.M, dm.C, dm.d, st, 0, nullptr, "gemm_logits");
}

int bf16_forward_decode(cocr_handle *h, const void *img, bool u8, int64_t sb, int64_t sw, int64_t sh, const int32_t *widths, int B,
                        int W, cudaStream_t st, int32_t *labels, int32_t *starts, int32_t *ends, float *conf, int32_t *counts,
                        int32_t *out_lens) {
    Bf16Plan p;
    COCR_TRY(bf16_forward(h, img, u8, sb, sw, sh, widths, B, W, st, out_lens, p));
    const Dims &dm = h->dm;
    char *ws = h->ws.base;
    unsigned long long *best = (unsigned long long *)(ws + p.off_best);
    float *bv = (float *)(ws + p.off_best_val);
    int32_t *bi = (int32_t *)(ws + p.off_best_idx);
    COCR_CUDA(h, cudaMemsetAsync(best, 0, (size_t)p.M * 8, st));
    // CTC projection fused with the per-frame arg-max: logits never reach HBM (pred.py:121 + :138-143)
    COCR_TRY(launch_gemm_tc(h, EPI_ARGMAX, (const bf16 *)(ws + p.off_xn), xn_pitch(h), WB(h, h->dec_w), dm.d, h->dec_b, nullptr, 0, best,
                            p.M, dm.C, dm.d, st, 0, nullptr, "gemm_ctc_argmax"));
    k_unpack_best<<<(p.M + 255) / 256, 256, 0, st>>>(best, p.M, bv, bi);
    COCR_LAUNCHED(h);
    return launch_collapse(h, bv, bi, out_lens, B, p.T, st, labels, starts, ends, conf, counts);
}

// ------------------------------------------------------------------------------------------
// self test of the tensor-core GEMM against a scalar device reference
// ------------------------------------------------------------------------------------------
__global__ void k_ref_gemm(const bf16 *A, const bf16 *W, const float *bias, float *C, int M, int N, int K, int silu) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    int m = (int)(idx / N), n = (int)(idx % N);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(__bfloat162float(A[(int64_t)m * K + k]), __bfloat162float(W[(int64_t)n * K + k]), acc);
    acc += bias[n];
    if (silu) acc = 2.0f * acc / (1.0f + expf(-2.0f * acc));   // the SiLU epilogue receives x/2
    C[idx] = acc;
}

__global__ void k_fill_bf16(bf16 *p, size_t n, uint32_t seed, float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = (uint32_t)i * 2654435761u + seed;
    s ^= s >> 16; s *= 2246822519u; s ^= s >> 13; s *= 3266489917u; s ^= s >> 16;
    p[i] = __float2bfloat16_rn(((float)(s & 0xFFFF) / 32768.0f - 1.0f) * scale);
}
__global__ void k_fill_f32(float *p, size_t n, uint32_t seed, float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = (uint32_t)i * 2654435761u + seed;
    s ^= s >> 16; s *= 2246822519u; s ^= s >> 13; s *= 3266489917u; s ^= s >> 16;
    p[i] = ((float)(s & 0xFFFF) / 32768.0f - 1.0f) * scale;
}

// which: 10 + epi ; p0 = M, p1 = N, p2 = K  (block_n chosen as in production; p2 >= 100000 encodes an override:
// K = p2 % 100000, block_n = p2 / 100000)
int selftest_gemm(int device, int epi, int M, int N, int K, int block_n, double *max_err) {
    cocr_hparams hp = {96, 144, 0, 4, 4, 2, 31, 1, 32, 4, 0, COCR_FP32, 0};
    cocr_handle *h = nullptr;
    int rc = cocr_create(&hp, device, &h);
    if (rc != COCR_OK) return rc;
    h->bf = new Bf16State();
    cudaDeviceGetAttribute(&h->bf->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    bf16 *A = nullptr, *Wt = nullptr;
    float *bias = nullptr, *ref = nullptr;
    void *out = nullptr;
    unsigned long long *best = nullptr;
    auto fail = [&](int code, const char *msg) {
        set_global_error(h->err.empty() ? msg : h->err.c_str());
        cudaFree(A); cudaFree(Wt); cudaFree(bias); cudaFree(ref); cudaFree(out); cudaFree(best);
        cocr_destroy(h);
        return code;
    };
    cudaMalloc((void **)&A, (size_t)M * K * 2);
    cudaMalloc((void **)&Wt, (size_t)N * K * 2);
    cudaMalloc((void **)&bias, (size_t)N * 4);
    cudaMalloc((void **)&ref, (size_t)M * N * 4);
    cudaMalloc(&out, (size_t)M * N * 4);
    cudaMalloc((void **)&best, (size_t)M * 8);
    cudaMemset(best, 0, (size_t)M * 8);
    cudaMemset(out, 0xFF, (size_t)M * N * 4);
    k_fill_bf16<<<(unsigned)(((size_t)M * K + 255) / 256), 256>>>(A, (size_t)M * K, 1u, 1.0f);
    k_fill_bf16<<<(unsigned)(((size_t)N * K + 255) / 256), 256>>>(Wt, (size_t)N * K, 2u, 0.25f);
    k_fill_f32<<<(N + 255) / 256, 256>>>(bias, (size_t)N, 3u, 0.5f);
    k_ref_gemm<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(A, Wt, bias, ref, M, N, K, epi == EPI_BF16_SILU);
    const int ldc = epi == EPI_BF16_GLU ? N / 2 : N;      // GLU: (value, gate) column pairs -> N/2 outputs per row
    rc = launch_gemm_tc(h, epi, A, K, Wt, K, bias, out, ldc, best, M, N, K, 0, block_n);
    if (rc != COCR_OK) return fail(rc, "launch failed");
    if (getenv("COCR_SELFTEST_TIME")) {     // timing aid for experiments: mean of 5 launches after the first
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, 0);
        for (int i = 0; i < 5; ++i) launch_gemm_tc(h, epi, A, K, Wt, K, bias, out, ldc, best, M, N, K, 0, block_n);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "selftest_gemm epi=%d M=%d N=%d K=%d block_n=%d: %.1f us/launch\n", epi, M, N, K, block_n, 1e3f * ms / 5);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error(h, COCR_ERR_CUDA, "selftest_gemm: %s", cudaGetErrorString(e));
        return fail(COCR_ERR_CUDA, "sync failed");
    }
    std::vector<float> href((size_t)M * N);
    cudaMemcpy(href.data(), ref, href.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0.0;
    if (epi == EPI_ARGMAX) {
        std::vector<unsigned long long> hb(M);
        cudaMemcpy(hb.data(), best, (size_t)M * 8, cudaMemcpyDeviceToHost);
        for (int m = 0; m < M; ++m) {
            uint32_t b = (uint32_t)(hb[m] >> 32);
            b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
            float v;
            memcpy(&v, &b, 4);
            int col = (int)(0xFFFFFFFFu - (uint32_t)(hb[m] & 0xFFFFFFFFull));
            float rv = -INFINITY;
            for (int n = 0; n < N; ++n) rv = fmaxf(rv, href[(size_t)m * N + n]);
            double err = fabs((double)v - rv);
            if (col < 0 || col >= N) err = 1e9;
            else err = fmax(err, fabs((double)href[(size_t)m * N + col] - rv));   // the chosen column must be a (near-)maximum
            worst = fmax(worst, err);
        }
    } else if (epi == EPI_F32) {
        std::vector<float> ho((size_t)M * N);
        cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < ho.size(); ++i) {
            double err = fabs((double)ho[i] - href[i]);
            if (!(err == err)) err = 1e9;
            worst = fmax(worst, err);
        }
    } else if (epi == EPI_BF16_GLU) {
        // the epilogue receives (value / 2, gate / 2) in columns (2c, 2c + 1): out[:, c] = v * sigmoid(g) = hv + hv * tanh(hg).
        // Every output of every N tile is checked (the buffer starts as NaN patterns): a tile must write its own columns only.
        const int No = N / 2;
        std::vector<uint16_t> ho((size_t)M * No);
        cudaMemcpy(ho.data(), out, ho.size() * 2, cudaMemcpyDeviceToHost);
        for (int m = 0; m < M; ++m)
            for (int c = 0; c < No; ++c) {
                const double hv = href[(size_t)m * N + 2 * c], hg = href[(size_t)m * N + 2 * c + 1];
                const double want = hv + hv * tanh(hg);
                uint32_t b = (uint32_t)ho[(size_t)m * No + c] << 16;
                float v;
                memcpy(&v, &b, 4);
                double err = fabs((double)v - want) / (1.0 + fabs(want));
                if (!(err == err)) err = 1e9;
                worst = fmax(worst, err);
            }
    } else {
        std::vector<uint16_t> ho((size_t)M * N);
        cudaMemcpy(ho.data(), out, ho.size() * 2, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < ho.size(); ++i) {
            uint32_t b = (uint32_t)ho[i] << 16;
            float v;
            memcpy(&v, &b, 4);
            double err = fabs((double)v - href[i]) / (1.0 + fabs((double)href[i]));
            if (!(err == err)) err = 1e9;
            worst = fmax(worst, err);
        }
    }
    if (max_err) *max_err = worst;
    cudaFree(A); cudaFree(Wt); cudaFree(bias); cudaFree(ref); cudaFree(out); cudaFree(best);
    cocr_destroy(h);
    return COCR_OK;
}

// residual epilogue vs host reference.  flags: bit0 init, bit1 block-final LayerNorm, bit2 next LayerNorm
int selftest_resid(int device, int M, int N, int K, int flags, double *max_err) {
    cocr_hparams hp = {96, 144, 0, 4, 4, 2, 31, 1, 32, 4, 0, COCR_FP32, 0};
    cocr_handle *h = nullptr;
    int rc = cocr_create(&hp, device, &h);
    if (rc != COCR_OK) return rc;
    h->bf = new Bf16State();
    cudaDeviceGetAttribute(&h->bf->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    bf16 *A, *Wt, *xn;
    float *bias, *ref, *x, *ln;
    cudaMalloc((void **)&A, (size_t)M * K * 2); cudaMalloc((void **)&Wt, (size_t)N * K * 2); cudaMalloc((void **)&xn, (size_t)M * N * 2);
    cudaMalloc((void **)&bias, (size_t)N * 4); cudaMalloc((void **)&ref, (size_t)M * N * 4); cudaMalloc((void **)&x, (size_t)M * N * 4);
    cudaMalloc((void **)&ln, (size_t)4 * N * 4);
    k_fill_bf16<<<(unsigned)(((size_t)M * K + 255) / 256), 256>>>(A, (size_t)M * K, 1u, 1.0f);
    k_fill_bf16<<<(unsigned)(((size_t)N * K + 255) / 256), 256>>>(Wt, (size_t)N * K, 2u, 0.25f);
    k_fill_f32<<<(N + 255) / 256, 256>>>(bias, (size_t)N, 3u, 0.5f);
    k_fill_f32<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(x, (size_t)M * N, 4u, 2.0f);
    k_fill_f32<<<(4 * N + 255) / 256, 256>>>(ln, (size_t)4 * N, 5u, 1.0f);
    k_ref_gemm<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(A, Wt, bias, ref, M, N, K, 0);
    std::vector<float> hx((size_t)M * N), href((size_t)M * N), hln(4 * N);
    cudaMemcpy(hx.data(), x, hx.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hln.data(), ln, hln.size() * 4, cudaMemcpyDeviceToHost);
    const bool init = flags & 1, fin = flags & 2, nxt = flags & 4;
    const float alpha = 0.5f;
    ResidArgs ra = {x, alpha, init ? 1 : 0, fin ? ln : nullptr, fin ? ln + N : nullptr, nxt ? ln + 2 * N : nullptr, nxt ? ln + 3 * N : nullptr};
    rc = launch_gemm_tc(h, EPI_RESID, A, K, Wt, K, bias, xn, N, nullptr, M, N, K, 0, 0, nullptr, "selftest", &ra);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc == COCR_OK && e != cudaSuccess) rc = set_error(h, COCR_ERR_CUDA, "selftest_resid: %s", cudaGetErrorString(e));
    double worst = 0.0;
    if (rc == COCR_OK) {
        cudaMemcpy(href.data(), ref, href.size() * 4, cudaMemcpyDeviceToHost);
        std::vector<float> gx((size_t)M * N);
        std::vector<uint16_t> gn((size_t)M * N);
        cudaMemcpy(gx.data(), x, gx.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(gn.data(), xn, gn.size() * 2, cudaMemcpyDeviceToHost);
        std::vector<double> v(N);
        // logical column -> position in the private layout of the residual stream (GemmArgs::x)
        auto xperm = [](int n) { const int gp = n >> 4, hh = (n >> 3) & 1, q = (n >> 1) & 3, e = n & 1; return 16 * gp + 4 * q + 2 * hh + e; };
        auto lnorm = [&](const float *g, const float *b) {
            double mu = 0, var = 0;
            for (int n = 0; n < N; ++n) mu += v[n];
            mu /= N;
            for (int n = 0; n < N; ++n) var += (v[n] - mu) * (v[n] - mu);
            var /= N;
            const double r = 1.0 / sqrt(var + 1e-5);
            for (int n = 0; n < N; ++n) v[n] = (v[n] - mu) * r * g[n] + b[n];
        };
        for (int m = 0; m < M; ++m) {
            for (int n = 0; n < N; ++n)
                v[n] = init ? (double)href[(size_t)m * N + n] : (double)hx[(size_t)m * N + xperm(n)] + alpha * (double)href[(size_t)m * N + n];
            if (fin) lnorm(hln.data(), hln.data() + N);
            for (int n = 0; n < N; ++n) {
                double err = fabs((double)gx[(size_t)m * N + xperm(n)] - v[n]) / (1.0 + fabs(v[n]));
                if (!(err == err)) err = 1e9;
                worst = fmax(worst, err * 8.0);      // fp32 stream: weighted 8x so one threshold covers both outputs
            }
            if (nxt) lnorm(hln.data() + 2 * N, hln.data() + 3 * N);
            for (int n = 0; n < N; ++n) {
                uint32_t b = (uint32_t)gn[(size_t)m * N + n] << 16;
                float f;
                memcpy(&f, &b, 4);
                double err = fabs((double)f - v[n]) / (1.0 + fabs(v[n]));
                if (!(err == err)) err = 1e9;
                worst = fmax(worst, err);
            }
        }
    } else {
        set_global_error(h->err.c_str());
    }
    if (max_err) *max_err = worst;
    cudaFree(A); cudaFree(Wt); cudaFree(xn); cudaFree(bias); cudaFree(ref); cudaFree(x); cudaFree(ln);
    cocr_destroy(h);
    return rc;
}


// fused FFN kernel against a host reference: M rows, ff = 96 * NC; flags: 2 block-final LayerNorm, 4 next LayerNorm;
// 8: time 20 launches and print us per launch to stderr
int selftest_ffn(int device, int M, int NC, int flags, double *max_err) {
    cocr_hparams hp = {96, 144, 0, 4, 4, 2, 31, 1, 32, 4, 0, COCR_FP32, 0};
    cocr_handle *h = nullptr;
    int rc = cocr_create(&hp, device, &h);
    if (rc != COCR_OK) return rc;
    h->bf = new Bf16State();
    cudaDeviceGetAttribute(&h->bf->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    const int d = FFN_D, ff = NC * FFN_HC;                         // NC counts 96-unit chunks for either kernel generation
    bf16 *xn, *w1, *w2, *w2p, *xo, *xnp, *w1p;
    float *b1, *b2, *x, *ln, *href;
    cudaMalloc((void **)&xn, (size_t)M * d * 2); cudaMalloc((void **)&xo, (size_t)M * d * 2);
    cudaMalloc((void **)&xnp, (size_t)M * FFN_K1 * 2); cudaMalloc((void **)&w1p, (size_t)ff * FFN_K1 * 2);
    cudaMalloc((void **)&w1, (size_t)ff * d * 2); cudaMalloc((void **)&w2, (size_t)ff * d * 2); cudaMalloc((void **)&w2p, (size_t)ff * d * 2);
    cudaMalloc((void **)&b1, (size_t)ff * 4); cudaMalloc((void **)&b2, (size_t)d * 4); cudaMalloc((void **)&x, (size_t)M * d * 4);
    cudaMalloc((void **)&ln, (size_t)4 * d * 4); cudaMalloc((void **)&href, (size_t)M * ff * 4);
    k_fill_bf16<<<(unsigned)(((size_t)M * d + 255) / 256), 256>>>(xn, (size_t)M * d, 11u, 1.0f);
    k_fill_bf16<<<(unsigned)(((size_t)ff * d + 255) / 256), 256>>>(w1, (size_t)ff * d, 12u, 0.2f);
    k_fill_bf16<<<(unsigned)(((size_t)ff * d + 255) / 256), 256>>>(w2, (size_t)ff * d, 13u, 0.1f);
    k_fill_f32<<<(ff + 255) / 256, 256>>>(b1, (size_t)ff, 14u, 0.5f);
    k_fill_f32<<<(d + 255) / 256, 256>>>(b2, (size_t)d, 15u, 0.5f);
    k_fill_f32<<<(unsigned)(((size_t)M * d + 255) / 256), 256>>>(x, (size_t)M * d, 16u, 2.0f);
    k_fill_f32<<<(4 * d + 255) / 256, 256>>>(ln, (size_t)4 * d, 17u, 1.0f);
    k_pack_ffn2<<<(unsigned)(((size_t)ff * d + 255) / 256), 256>>>(w2, w2p, d, ff, ffn_hc(ff));
    k_pack_ffn1<<<(unsigned)(((size_t)ff * FFN_K1 + 255) / 256), 256>>>(w1, b1, w1p, d, ff, FFN_K1);
    cudaMemcpy2D(xnp, (size_t)FFN_K1 * 2, xn, (size_t)d * 2, (size_t)d * 2, (size_t)M, cudaMemcpyDeviceToDevice);
    k_xn_pad_init<<<(unsigned)(((size_t)M * (FFN_K1 - d) + 255) / 256), 256>>>(xnp, M, d, FFN_K1);
    k_ref_gemm<<<(unsigned)(((size_t)M * ff + 255) / 256), 256>>>(xn, w1, b1, href, M, ff, d, 0);     // h = xn W1^T + b1 (the "half" pre-activation)
    std::vector<float> hx((size_t)M * d), hh((size_t)M * ff), hln(4 * d), hb2(d);
    std::vector<uint16_t> hw2((size_t)ff * d);
    cudaMemcpy(hx.data(), x, hx.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hh.data(), href, hh.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hln.data(), ln, hln.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb2.data(), b2, hb2.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hw2.data(), w2, hw2.size() * 2, cudaMemcpyDeviceToHost);
    const bool fin = flags & 2, nxt = flags & 4;
    const float alpha = 0.5f;
    rc = launch_ffn_fused(h, xnp, FFN_K1, w1p, w2p, b2, x, xo, d, alpha, fin ? ln : nullptr, fin ? ln + d : nullptr, nxt ? ln + 2 * d : nullptr,
                          nxt ? ln + 3 * d : nullptr, M, ff, 0, "selftest_ffn");
    cudaError_t e = cudaDeviceSynchronize();
    if (rc == COCR_OK && e != cudaSuccess) rc = set_error(h, COCR_ERR_CUDA, "selftest_ffn: %s", cudaGetErrorString(e));
    double worst = 0.0;
    if (rc == COCR_OK) {
        std::vector<float> gx((size_t)M * d);
        std::vector<uint16_t> gn((size_t)M * d);
        cudaMemcpy(gx.data(), x, gx.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(gn.data(), xo, gn.size() * 2, cudaMemcpyDeviceToHost);
        auto bf2f = [](uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; };
        auto f2bf_rn = [](float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); };
        auto xperm = [](int n) { const int gp = n >> 4, hh2 = (n >> 3) & 1, q = (n >> 1) & 3, e2 = n & 1; return 16 * gp + 4 * q + 2 * hh2 + e2; };
        std::vector<double> v(d), act(ff);
        auto lnorm = [&](const float *g, const float *b) {
            double mu = 0, var = 0;
            for (int n = 0; n < d; ++n) mu += v[n];
            mu /= d;
            for (int n = 0; n < d; ++n) var += (v[n] - mu) * (v[n] - mu);
            var /= d;
            const double r = 1.0 / sqrt(var + 1e-5);
            for (int n = 0; n < d; ++n) v[n] = (v[n] - mu) * r * g[n] + b[n];
        };
        for (int m = 0; m < M; ++m) {
            for (int j = 0; j < ff; ++j) {
                const double hv = hh[(size_t)m * ff + j];
                act[j] = bf2f(f2bf_rn((float)(hv + hv * tanh(hv))));          // silu(2h), rounded to bf16 like the kernel's operand
            }
            for (int n = 0; n < d; ++n) {
                double acc = hb2[n];
                for (int j = 0; j < ff; ++j) acc += act[j] * (double)bf2f(hw2[(size_t)n * ff + j]);
                v[n] = (double)hx[(size_t)m * d + xperm(n)] + alpha * acc;
            }
            if (fin) lnorm(hln.data(), hln.data() + d);
            for (int n = 0; n < d; ++n) {
                double err = fabs((double)gx[(size_t)m * d + xperm(n)] - v[n]) / (1.0 + fabs(v[n]));
                if (!(err == err)) err = 1e9;
                worst = fmax(worst, err);
            }
            if (nxt) lnorm(hln.data() + 2 * d, hln.data() + 3 * d);
            for (int n = 0; n < d; ++n) {
                double err = fabs((double)bf2f(gn[(size_t)m * d + n]) - v[n]) / (1.0 + fabs(v[n]));
                if (!(err == err)) err = 1e9;
                worst = fmax(worst, err);
            }
        }
        if (flags & 8) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int i = 0; i < 1500; ++i)      // ~100 ms of warm-up: the SM clock ramps up only under sustained load
                launch_ffn_fused(h, xnp, FFN_K1, w1p, w2p, b2, x, xo, d, alpha, nullptr, nullptr, ln + 2 * d, ln + 3 * d, M, ff, 0, "selftest_ffn");
            cudaEventRecord(e0);
            for (int i = 0; i < 20; ++i)
                launch_ffn_fused(h, xnp, FFN_K1, w1p, w2p, b2, x, xo, d, alpha, nullptr, nullptr, ln + 2 * d, ln + 3 * d, M, ff, 0, "selftest_ffn");
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            fprintf(stderr, "selftest_ffn M=%d ff=%d: %.1f us per launch\n", M, ff, ms * 1000.0f / 20.0f);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
    } else {
        set_global_error(h->err.c_str());
    }
    if (max_err) *max_err = worst;
    cudaFree(xnp); cudaFree(w1p); cudaFree(xn); cudaFree(xo); cudaFree(w1); cudaFree(w2); cudaFree(w2p); cudaFree(b1); cudaFree(b2); cudaFree(x); cudaFree(ln); cudaFree(href);
    cocr_destroy(h);
    return rc;
}


}  // namespace cocr
