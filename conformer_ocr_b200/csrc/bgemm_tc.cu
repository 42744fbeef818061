// Batched tcgen05 GEMM for the training step (train.cu): D[z] (M x N) = alpha * A[z] (M x K) * B[z] (N x K)^T (+ bias),
// bf16 operands fed by 4-D TMA tensor maps, fp32 accumulation in TMEM.  Either operand may be K-major (stored row-major
// with the contraction index contiguous) or MN-major (stored [k][m]: the contraction index runs along the rows), so the
// same kernel computes the forward linears (NT), the data gradients (NN: B = W stored (N_w, K_w) read MN-major), the weight
// gradients (TN: both operands MN-major, contraction over the B*T frames split over CTAs into partial products or fp32 atomics) and every
// contraction of the attention module per (line, head).  Out-of-range rows / columns / contraction indices of a batch
// item are zero-filled by the tensor map (its extents are the batch item's own), so no operand needs padding.
//
// One CTA = one 128 x block_n output tile of one batch item and one k-slice: warp 0 TMA producer, warp 1 MMA issuer
// (one elected lane), warps 2-5 epilogue (TMEM -> registers -> a swizzled per-warp staging tile -> global, so that a store
// instruction covers whole 128-byte rows).  Several CTAs are resident per SM (<= 112 KB of shared memory, <= 256 TMEM columns
// each), which is what overlaps the epilogue of one tile with the loads and MMAs of another.
#include <cuda.h>

#include "common.h"
#include "tc_utils.cuh"

namespace cocr {

using namespace tc;

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

constexpr int BG_STAGES = 3;               // at most; short contractions use fewer so that more CTAs share an SM
constexpr uint32_t BG_A_BYTES = 128 * 128;      // 128 (m) x 64 (k) bf16, either major
constexpr int BG_THREADS = 192;

struct BgArgs {
    int M, N, K;
    int block_n, nbox_b;           // nbox_b: 64-column boxes of an MN-major B tile
    int a_mn, b_mn;
    int nz2, ksplit, kchunks;      // batch z = z1 * nz2 + z2; k-chunks (of 64) per split
    int out_mode;                  // 0 fp32 store, 1 bf16 store, 2 fp32 atomic add, 3 fp32 read-add-store
    int tmem_cols;
    float alpha;
    const float *bias;             // (N) or null; added by k-slice 0 only
    void *out;
    long long ldc, out_z1, out_z2, out_ks; // elements
    int relu, stages;
    int a_u1, a_u2, b_u1, b_u2;    // 1 when the operand has its own block per z1 / z2 (0: shared by the batch)
};

__global__ void __launch_bounds__(BG_THREADS)
k_bgemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const BgArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[BG_STAGES], empty[BG_STAGES], tfull;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b_bytes = p.b_mn ? (uint32_t)p.nbox_b * 8192u : (uint32_t)p.block_n * 128u;
    const uint32_t stage_bytes = (BG_A_BYTES + b_bytes + 1023u) & ~1023u;
    const int n0 = blockIdx.x * p.block_n, m0 = blockIdx.y * 128;
    const int z = blockIdx.z / p.ksplit, ks_idx = blockIdx.z % p.ksplit;
    const int z1 = z / p.nz2, z2 = z % p.nz2;
    const int nkc_total = (p.K + 63) / 64;
    const int kc0 = ks_idx * p.kchunks;
    int nkc = nkc_total - kc0;
    if (nkc > p.kchunks) nkc = p.kchunks;
    if (nkc < 0) nkc = 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&tfull, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tmA);
            prefetch_tmap(&tmB);
            for (int kc = 0; kc < nkc; ++kc) {
                const int s = kc % p.stages;
                const uint32_t ph = (uint32_t)(kc / p.stages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                uint8_t *a_dst = smem + (size_t)s * stage_bytes, *b_dst = a_dst + BG_A_BYTES;
                mbar_expect_tx(&full[s], BG_A_BYTES + b_bytes);
                const int k0 = (kc0 + kc) * 64;
                const int az1 = z1 * p.a_u1, az2 = z2 * p.a_u2, bz1 = z1 * p.b_u1, bz2 = z2 * p.b_u2;
                if (p.a_mn) {            // stored [k][m]: two boxes of 64 (k) x 64 (m)
                    tma_load_4d(a_dst, &tmA, &full[s], m0, k0, az2, az1);
                    tma_load_4d(a_dst + 8192, &tmA, &full[s], m0 + 64, k0, az2, az1);
                } else {                 // stored [m][k]: one box of 128 (m) x 64 (k)
                    tma_load_4d(a_dst, &tmA, &full[s], k0, m0, az2, az1);
                }
                if (p.b_mn) {
                    for (int bx = 0; bx < p.nbox_b; ++bx) tma_load_4d(b_dst + (size_t)bx * 8192, &tmB, &full[s], n0 + bx * 64, k0, bz2, bz1);
                } else {
                    tma_load_4d(b_dst, &tmB, &full[s], k0, n0, bz2, bz1);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_bf16(128, p.block_n, p.a_mn, p.b_mn);
        for (int kc = 0; kc < nkc; ++kc) {
            const int s = kc % p.stages;
            const uint32_t ph = (uint32_t)(kc / p.stages) & 1u;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes), b_base = a_base + BG_A_BYTES;
            int nks = (p.K - (kc0 + kc) * 64 + 15) / 16;
            if (nks > 4) nks = 4;
            if (elect_one()) {
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t ad = p.a_mn ? make_smem_desc_sw128(a_base + ks * 2048, 1024, 8192) : make_smem_desc_sw128(a_base + ks * 32, 1024);
                    const uint64_t bd = p.b_mn ? make_smem_desc_sw128(b_base + ks * 2048, 1024, 8192) : make_smem_desc_sw128(b_base + ks * 32, 1024);
                    umma_bf16_ss(tmem_base, ad, bd, idesc, (kc | ks) ? 1u : 0u);
                }
                umma_commit(&empty[s]);
                if (kc == nkc - 1) umma_commit(&tfull);
            }
            __syncwarp();
        }
    } else if (nkc > 0) {
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        mbar_wait(&tfull, 0);
        tc_fence_after();
        const long long zbase = (long long)z1 * p.out_z1 + (long long)z2 * p.out_z2 + (long long)ks_idx * p.out_ks;
        const long long obase = zbase + (long long)m * p.ldc;
        const bool add_bias = p.bias && ks_idx == 0;
        // fp32 stores go through a per-warp staging tile so that a store instruction covers whole 128-byte rows (4 rows x 128 B)
        // instead of 32 rows x 16 B: the L2 handles a quarter of the write transactions (97 MB score tiles: 87 -> see profiles/)
        const bool staged = p.out_mode == 0 && (p.ldc & 3) == 0 && (((p.out_z1 | p.out_z2 | p.out_ks) & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && ((p.block_n & 31) == 0 || p.block_n >= 32);
        float *stg = reinterpret_cast<float *>(smem + (size_t)p.stages * stage_bytes) + (warp - 2) * 1024;      // [32 rows][32 floats], 16-byte chunks XOR-swizzled by row
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
            const int ncol = min(32, p.block_n - c0);          // 32 or 16 (block_n is a multiple of 16)
            uint32_t r[32];
            tmem_ld_32x32b_x16_lo(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            if (ncol > 16) {
                uint32_t r2[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + 16), r2);
                tmem_ld_wait_dep(r2);
#pragma unroll
                for (int e = 0; e < 16; ++e) r[16 + e] = r2[e];
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) r[16 + e] = 0u;
            }
            tmem_ld_wait_dep(r);
            float v[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int n = n0 + c0 + e;
                float t = __uint_as_float(r[e]) * p.alpha;
                if (add_bias && n < p.N) t += p.bias[n];
                if (p.relu) t = fmaxf(t, 0.0f);
                v[e] = t;
            }
            const int nvalid = min(ncol, p.N - (n0 + c0));
            if (staged) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4 *>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                __syncwarp();
                const int ch = lane & 7;                        // 16-byte chunk of the row this lane stores
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3);
                    const int mm = m0 + q * 32 + rr;
                    const float4 w4 = *reinterpret_cast<const float4 *>(stg + rr * 32 + ((ch ^ (rr & 7)) << 2));
                    if (mm < p.M && nvalid > 0) {
                        float *o = (float *)p.out + zbase + (long long)mm * p.ldc + n0 + c0 + ch * 4;
                        const int left = nvalid - ch * 4;
                        if (left >= 4) *reinterpret_cast<float4 *>(o) = w4;
                        else {
                            if (left > 0) o[0] = w4.x;
                            if (left > 1) o[1] = w4.y;
                            if (left > 2) o[2] = w4.z;
                        }
                    }
                }
                __syncwarp();
                continue;
            }
            if (m >= p.M || nvalid <= 0) continue;
            if (p.out_mode == 0) {
                float *o = (float *)p.out + obase + n0 + c0;
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (e < nvalid) o[e] = v[e];
            } else if (p.out_mode == 1) {
                __nv_bfloat16 *o = (__nv_bfloat16 *)p.out + obase + n0 + c0;
                if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                    uint32_t w[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                        w[e] = *reinterpret_cast<const uint32_t *>(&h2);
                    }
#pragma unroll
                    for (int e = 0; e < 16; e += 4) *reinterpret_cast<uint4 *>(o + 2 * e) = make_uint4(w[e], w[e + 1], w[e + 2], w[e + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (e < nvalid) o[e] = __float2bfloat16_rn(v[e]);
                }
            } else if (p.out_mode == 2) {
                float *o = (float *)p.out + obase + n0 + c0;
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (e < nvalid) atomicAdd(o + e, v[e]);
            } else {
                float *o = (float *)p.out + obase + n0 + c0;
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (e < nvalid) o[e] += v[e];
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

// 4-D bf16 tensor map of one operand: dims (inner, rows, z2, z1), strides in elements, box (box_inner x box_rows x 1 x 1)
int make_tmap4(cocr_handle *h, CUtensorMap *m, const void *base, const BgOperand &o, int inner, int rows, int nz2, int nz1, int box_inner, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return set_error(h, COCR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (o.ld % 8 != 0 || (nz2 > 1 && o.z2_stride % 8 != 0) || (nz1 > 1 && o.z1_stride % 8 != 0) || (reinterpret_cast<uintptr_t>(base) & 15))
        return set_error(h, COCR_ERR_UNSUPPORTED, "batched GEMM operand needs 16-byte aligned rows and batch strides (ld=%lld z1=%lld z2=%lld)",
                         o.ld, o.z1_stride, o.z2_stride);
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)(nz2 < 1 ? 1 : nz2), (cuuint64_t)(nz1 < 1 ? 1 : nz1)};
    // a stride of 0 (operand shared by the batch) is not accepted by the encoder: give unit-extent dimensions any aligned stride
    auto st = [&](long long s, int n) { return (cuuint64_t)((n > 1 && s > 0 ? s : o.ld) * 2); };
    if ((nz2 > 1 && o.z2_stride == 0)) dims[2] = 1;
    if ((nz1 > 1 && o.z1_stride == 0)) dims[3] = 1;
    cuuint64_t strides[3] = {(cuuint64_t)o.ld * 2, st(o.z2_stride, (int)dims[2]), st(o.z1_stride, (int)dims[3])};
    cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(h, COCR_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed (%d): inner=%d rows=%d ld=%lld z=(%d,%d) strides=(%lld,%lld) box=%dx%d",
                         (int)r, inner, rows, o.ld, nz1, nz2, o.z1_stride, o.z2_stride, box_inner, box_rows);
    return COCR_OK;
}

}  // namespace

int launch_bgemm(cocr_handle *h, const BgProblem &g, cudaStream_t st, const char *tag) {
    if (g.M <= 0 || g.N <= 0 || g.nz1 <= 0 || g.nz2 <= 0) return COCR_OK;
    BgArgs p;
    p.M = g.M; p.N = g.N; p.K = g.K;
    int bn = g.N <= 256 ? (g.N + 15) / 16 * 16 : 0;
    if (!bn) {
        // several N tiles: an exact divisor (multiple of 16; of 64 for an MN-major B) if there is one
        const int step = g.b.mn ? 64 : 16;
        for (int c = 256; c >= 64; c -= step)
            if (g.N % c == 0) { bn = c; break; }
        if (!bn) bn = 256;
    }
    if (g.K <= 128 && g.N > 128) bn = 128;        // short contraction, wide output: the epilogue dominates; 4 CTAs per SM overlap it
    if (g.b.mn && g.N > bn && bn % 64 != 0) bn = (bn + 63) / 64 * 64;
    if (bn > 256) bn = 256;
    p.block_n = bn;
    p.nbox_b = (bn + 63) / 64;
    p.a_mn = g.a.mn; p.b_mn = g.b.mn;
    p.nz2 = g.nz2;
    const int nkc = (g.K + 63) / 64;
    int ks = g.ksplit < 1 ? 1 : g.ksplit;
    if (ks > nkc) ks = nkc < 1 ? 1 : nkc;
    p.kchunks = (nkc + ks - 1) / ks;
    p.ksplit = p.kchunks > 0 ? (nkc + p.kchunks - 1) / p.kchunks : 1;
    if (p.ksplit < 1) p.ksplit = 1;
    if (p.ksplit > 1 && g.out_mode != 2 && !(g.out_mode == 0 && g.out_ks > 0))
        return set_error(h, COCR_ERR_BAD_ARG, "batched GEMM: a split contraction needs the atomic output mode or a stride between the partial results");
    p.out_mode = g.out_mode; p.alpha = g.alpha; p.bias = g.bias; p.out = g.out; p.ldc = g.ldc; p.out_z1 = g.out_z1; p.out_z2 = g.out_z2; p.relu = g.relu; p.out_ks = g.out_ks;
    p.a_u1 = (g.nz1 > 1 && g.a.z1_stride == 0) ? 0 : 1; p.a_u2 = (g.nz2 > 1 && g.a.z2_stride == 0) ? 0 : 1;
    p.b_u1 = (g.nz1 > 1 && g.b.z1_stride == 0) ? 0 : 1; p.b_u2 = (g.nz2 > 1 && g.b.z2_stride == 0) ? 0 : 1;
    int tm = 32;
    while (tm < bn) tm <<= 1;
    p.tmem_cols = tm;
    CUtensorMap tmA, tmB;
    // K-major: inner = k, rows = m; MN-major: inner = m, rows = k
    if (g.a.mn) COCR_TRY(make_tmap4(h, &tmA, g.a.ptr, g.a, g.M, g.K, g.nz2, g.nz1, 64, 64));
    else COCR_TRY(make_tmap4(h, &tmA, g.a.ptr, g.a, g.K, g.M, g.nz2, g.nz1, 64, 128));
    if (g.b.mn) COCR_TRY(make_tmap4(h, &tmB, g.b.ptr, g.b, g.N, g.K, g.nz2, g.nz1, 64, 64));
    else COCR_TRY(make_tmap4(h, &tmB, g.b.ptr, g.b, g.K, g.N, g.nz2, g.nz1, 64, bn));
    const uint32_t b_bytes = g.b.mn ? (uint32_t)p.nbox_b * 8192u : (uint32_t)bn * 128u;
    const uint32_t stage_bytes = (BG_A_BYTES + b_bytes + 1023u) & ~1023u;
    p.stages = p.kchunks < BG_STAGES ? (p.kchunks < 1 ? 1 : p.kchunks) : BG_STAGES;
    if (p.stages == 3 && 3 * stage_bytes > 96 * 1024 && 2 * stage_bytes <= 96 * 1024) p.stages = 2;   // two CTAs per SM beat a third stage
    const size_t smem = (size_t)p.stages * stage_bytes + 4 * 4096;      // + one 32 x 32 fp32 staging tile per epilogue warp
    static PerDevice<size_t> attr;
    if (smem > attr.get()) {
        COCR_CUDA(h, cudaFuncSetAttribute(k_bgemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BG_STAGES * (BG_A_BYTES + 32768) + 4 * 4096)));
        attr.get() = BG_STAGES * (BG_A_BYTES + 32768) + 4 * 4096;
    }
    const long long nz = (long long)g.nz1 * g.nz2 * p.ksplit;
    if (nz > 65535) return set_error(h, COCR_ERR_UNSUPPORTED, "batched GEMM: %lld batch slices exceed the grid", nz);
    dim3 grid((g.N + bn - 1) / bn, (g.M + 127) / 128, (unsigned)nz);
    const double batches = (double)g.nz1 * g.nz2;
    ProfScope ps(h, tag, st, 2.0 * batches * g.M * (double)g.N * g.K, batches * (2.0 * ((double)g.M * g.K + (double)g.N * g.K) + (g.out_mode == 1 ? 2.0 : 4.0) * g.M * g.N));
    k_bgemm_tc<<<grid, BG_THREADS, smem, st>>>(tmA, tmB, p);
    COCR_LAUNCHED(h);
    return COCR_OK;
}

// ------------------------------------------------------------------------------------------
// self test (cocr_selftest which = 30): every operand-major combination, batched, against a scalar device reference
// ------------------------------------------------------------------------------------------
namespace {
__global__ void k_bg_fill(__nv_bfloat16 *p, size_t n, uint32_t seed) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i * 2654435761u + seed * 40503u;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        p[i] = __float2bfloat16_rn(((float)(x & 0xffff) / 65536.0f - 0.5f) * 2.0f);
    }
}
// reference: element (z, m, n)
__global__ void k_bg_ref(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int M, int N, int K, int a_mn, int b_mn, long long lda, long long ldb,
                         long long az1, long long az2, long long bz1, long long bz2, int nz2, long long ldc, long long oz1, long long oz2, float alpha,
                         const float *bias, long long total) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e % N);
        const int m = (int)((e / N) % M);
        const int z = (int)(e / ((long long)N * M));
        const int z1 = z / nz2, z2 = z % nz2;
        const __nv_bfloat16 *a = A + z1 * az1 + z2 * az2, *b = B + z1 * bz1 + z2 * bz2;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            const float av = __bfloat162float(a_mn ? a[(long long)k * lda + m] : a[(long long)m * lda + k]);
            const float bv = __bfloat162float(b_mn ? b[(long long)k * ldb + n] : b[(long long)n * ldb + k]);
            acc = fmaf(av, bv, acc);
        }
        D[z1 * oz1 + z2 * oz2 + (long long)m * ldc + n] = alpha * acc + (bias ? bias[n] : 0.f);
    }
}
}  // namespace

int selftest_bgemm(int device, int M, int N, int K, int flags, double *max_err) {
    // flags: bit0 a_mn, bit1 b_mn, bit2 batched (3 x 2), bit3 split contraction + atomic output
    cocr_hparams hp = {};
    hp.height = 8; hp.encoder_dim = 8; hp.num_layers = 0; hp.num_heads = 1; hp.ff_expansion = 1; hp.conv_expansion = 2; hp.conv_kernel_size = 3;
    hp.half_step_residual = 1; hp.subsampling_conv_channels = 1; hp.subsampling_factor = 4; hp.num_classes = 0;
    cocr_handle *h = nullptr;
    int rc = cocr_create(&hp, device, &h);
    if (rc != COCR_OK) return rc;
    const int a_mn = flags & 1, b_mn = (flags >> 1) & 1, batched = (flags >> 2) & 1, split = (flags >> 3) & 1;
    const int nz1 = batched ? 3 : 1, nz2 = batched ? 2 : 1;
    auto up8 = [](long long v) { return (v + 7) / 8 * 8; };
    // operand storage: per batch item a (rows x cols) block inside a wider matrix: z2 shifts columns, z1 shifts rows
    const long long a_rows = a_mn ? K : M, a_cols = a_mn ? M : K, b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
    const long long lda = up8(a_cols) * nz2 + 8, ldb = up8(b_cols) * nz2 + 16;
    const long long az2 = up8(a_cols), az1 = (a_rows + 3) * lda, bz2 = up8(b_cols), bz1 = (b_rows + 5) * ldb;
    const size_t na = (size_t)az1 * nz1 + lda, nb = (size_t)bz1 * nz1 + ldb;
    const long long ldc = N + 3, oz2 = (long long)M * ldc + 7, oz1 = oz2 * nz2;
    const size_t nd = (size_t)oz1 * nz1;
    __nv_bfloat16 *A, *B;
    float *D, *R, *bias;
    cudaMalloc((void **)&A, na * 2); cudaMalloc((void **)&B, nb * 2);
    cudaMalloc((void **)&D, nd * 4); cudaMalloc((void **)&R, nd * 4); cudaMalloc((void **)&bias, (size_t)N * 4);
    k_bg_fill<<<256, 256>>>(A, na, 1u);
    k_bg_fill<<<256, 256>>>(B, nb, 2u);
    std::vector<float> hb(N);
    for (int n = 0; n < N; ++n) hb[n] = 0.01f * (float)((n * 37) % 101 - 50);
    cudaMemcpy(bias, hb.data(), (size_t)N * 4, cudaMemcpyHostToDevice);
    cudaMemset(D, 0, nd * 4); cudaMemset(R, 0, nd * 4);
    BgProblem g;
    g.M = M; g.N = N; g.K = K; g.nz1 = nz1; g.nz2 = nz2;
    g.a = {A, a_mn, lda, az1, az2};
    g.b = {B, b_mn, ldb, bz1, bz2};
    g.out = D; g.ldc = ldc; g.out_z1 = oz1; g.out_z2 = oz2; g.out_mode = split ? 2 : 0; g.alpha = 0.5f; g.bias = bias; g.ksplit = split ? 5 : 1; g.relu = 0;
    rc = launch_bgemm(h, g, 0, "selftest_bgemm");
    if (rc == COCR_OK) {
        const long long total = (long long)nz1 * nz2 * M * N;
        k_bg_ref<<<1024, 256>>>(A, B, R, M, N, K, a_mn, b_mn, lda, ldb, az1, az2, bz1, bz2, nz2, ldc, oz1, oz2, 0.5f, bias, total);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = set_error(h, COCR_ERR_CUDA, "selftest_bgemm: %s", cudaGetErrorString(e));
    }
    if (rc == COCR_OK) {
        std::vector<float> d(nd), r(nd);
        cudaMemcpy(d.data(), D, nd * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(r.data(), R, nd * 4, cudaMemcpyDeviceToHost);
        double worst = 0.0;
        for (size_t i = 0; i < nd; ++i) {
            const double e = fabs((double)d[i] - (double)r[i]);
            if (!(e <= worst)) worst = e == e ? e : 1e30;
        }
        if (max_err) *max_err = worst;
    } else {
        set_global_error(cocr_last_error(h));
    }
    cudaFree(A); cudaFree(B); cudaFree(D); cudaFree(R); cudaFree(bias);
    cocr_destroy(h);
    return rc;
}

}  // namespace cocr
