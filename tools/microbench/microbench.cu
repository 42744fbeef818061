// Micro-benchmarks of the tensor pipe, TMEM, MUFU and mbarrier hand-shakes.  NOT part of the product library: built into
// tools/microbench/libcocr_microbench.so (make -C tools/microbench) and called through `cocr_microbench` by tools/*_bench.py.
//   which = 40: cycles of a chain of tcgen05.mma instructions for a given N, number of independent accumulators the chain
//   alternates between, and A operand source (shared / tensor memory); 41-45: see the entry point at the end of the file.
#include <stdarg.h>
#include <stdio.h>

#include "../../conformer_ocr_b200/csrc/common.h"
#include "../../conformer_ocr_b200/csrc/tc_utils.cuh"

namespace cocr {

static char g_mb_err[512] = "";
static int mb_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_mb_err, sizeof(g_mb_err), fmt, ap);
    va_end(ap);
    return code;
}
#define set_error(h, code, ...) mb_error(code, __VA_ARGS__)

__global__ void __launch_bounds__(128, 1) k_mma_bench(int N, int chains, int a_tmem, int count, long long *out) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 1) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 16384);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            t0 = clock64();
            if (elect_one()) {
                for (int i = 0; i < count; ++i) {
                    const uint32_t d = tb + (uint32_t)((i % chains) * N);
                    const uint64_t bd = make_smem_desc_sw128(b_base + (i & 3) * 32, 1024);
                    if (a_tmem) umma_bf16_ts(d, tb + 448u + (uint32_t)((i & 3) * 8), bd, idesc, i >= chains ? 1u : 0u);
                    else umma_bf16_ss(d, make_smem_desc_sw128(a_base + (i & 3) * 32, 1024), bd, idesc, i >= chains ? 1u : 0u);
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, rep & 1);
            t1 = clock64();
        }
        if (threadIdx.x == 32) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

// TMEM read bandwidth: 8 warps (two per lane quarter) loop over tcgen05.ld 32x32b.x32 (4 KB per warp-load), `depth`
// loads in flight per warp; optionally warp 8 keeps the tensor pipe busy with N=256 MMAs into other columns.
__global__ void __launch_bounds__(320, 1) k_tmem_ld_bench(int iters, int depth, int with_mma, int shape, long long *out) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
    if (warp == 9) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp < 8) {
        const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t a[32], b[32];
            if (shape == 0) {
                tmem_ld_32x32b_x32(base, a);
                if (depth > 1) tmem_ld_32x32b_x32(base + 32, b);
            } else {
                tmem_ld_16x256b_x8(base, a);
                if (depth > 1) tmem_ld_16x256b_x8(base + ((uint32_t)16 << 16), b);
            }
            tmem_ld_wait_dep(a);
            if (depth > 1) tmem_ld_wait_dep(b);
            acc ^= a[0] ^ a[31];
            if (depth > 1) acc ^= b[0] ^ b[31];
        }
        const long long t1 = clock64();
        if (acc == 0x12345u) out[7] = acc;
        if (threadIdx.x == 0) { out[0] = t1 - t0; stop = 1; }
    } else if (warp == 8 && with_mma) {
        const uint32_t idesc = make_idesc_bf16(128, 256);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 16384);
        int n = 0;
        while (!stop) {
            if (elect_one()) {
                for (int i = 0; i < 8; ++i)
                    umma_bf16_ss(tb + 256u, make_smem_desc_sw128(a_base + (i & 3) * 32, 1024), make_smem_desc_sw128(b_base + (i & 3) * 32, 1024), idesc, 1u);
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, n & 1);
            ++n;
        }
        if (lane == 0) out[1] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

// MMA rate under TMEM load/store traffic from the other warps: warp 8 issues `groups` groups of 12 TS N=128 MMAs (commit +
// wait per group, like one S tile of the attention kernel) while warps 0-7 loop over tcgen05.ld (mode 1), ld + st (mode 2)
// or stay idle (mode 0).
__global__ void __launch_bounds__(320, 1) k_mma_contention_bench(int groups, int mode, int a_tmem, long long *out) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
    if (warp == 9) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp < 8) {
        if ((mode % 10) > 0) {
            const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + 128u + (uint32_t)((warp >> 2) * 64);   // the "other" S buffer
            uint32_t acc = 0;
            while (!stop) {
                uint32_t a[32];
                tmem_ld_32x32b_x32(base, a);
                tmem_ld_wait_dep(a);
                acc ^= a[0] ^ a[31];
                if ((mode % 10) > 1) {
                    tmem_st_32x32b_x32(base + 32u, a);
                    tmem_st_wait();
                }
            }
            if (acc == 0x12345u) out[7] = acc;
        }
    } else if (warp == 8) {
        const uint32_t idesc = make_idesc_bf16(128, 128);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 16384);
        const long long t0 = clock64();
        for (int n = 0; n < groups; ++n) {
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 12; ++i) {
                    const uint64_t bd = make_smem_desc_sw128(b_base + (uint32_t)(i >> 2) * 16384u * (uint32_t)(mode >= 10) + (i & 3) * 32, 1024);
                    if (a_tmem) umma_bf16_ts(tb, tb + 320u + (uint32_t)(i * 8), bd, idesc, i ? 1u : 0u);
                    else umma_bf16_ss(tb, make_smem_desc_sw128(a_base + (i & 3) * 32, 1024), bd, idesc, i ? 1u : 0u);
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, n & 1);
        }
        const long long t1 = clock64();
        if (threadIdx.x == 256) { out[0] = t1 - t0; stop = 1; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int selftest_mma_contention(int device, int mode, int a_tmem, double *cycles_per_group) {
    const int groups = 200;
    long long *d = nullptr, hres[8] = {0};
    cudaMalloc((void **)&d, 64);
    cudaMemset(d, 0, 64);
    cudaFuncSetAttribute(k_mma_contention_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k_mma_contention_bench<<<1, 320, 64 * 1024>>>(groups, mode, a_tmem, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); return set_error(nullptr, COCR_ERR_CUDA, "mma contention bench: %s", cudaGetErrorString(e)); }
    cudaMemcpy(hres, d, 64, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (cycles_per_group) *cycles_per_group = (double)hres[0] / groups;
    return COCR_OK;
}

int selftest_tmem_bench(int device, int depth, int with_mma, int shape, double *bytes_per_cycle) {
    const int iters = 2000;
    long long *d = nullptr, hres[8] = {0};
    cudaMalloc((void **)&d, 64);
    cudaMemset(d, 0, 64);
    cudaFuncSetAttribute(k_tmem_ld_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k_tmem_ld_bench<<<1, 320, 64 * 1024>>>(iters, depth, with_mma, shape, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); return set_error(nullptr, COCR_ERR_CUDA, "tmem bench: %s", cudaGetErrorString(e)); }
    cudaMemcpy(hres, d, 64, cudaMemcpyDeviceToHost);
    cudaFree(d);
    const double bytes = 8.0 * iters * (depth > 1 ? 2 : 1) * 4096.0;
    if (bytes_per_cycle) *bytes_per_cycle = bytes / (double)hres[0];
    return COCR_OK;
}

// The attention kernel's tensor-pipe sequence without any softmax: per key tile one S group (12 TS MMAs, N=128, K-major B,
// alternating accumulators) and one P V group (8 TS MMAs into O).  mode bits: 1 = P V with MN-major B (as in the kernel),
// 2 = P V with N=48 instead of 64, 4 = no P V, 8 = no S, 16 = wait for every group's commit (else every 4th pair),
// 32 = P V as 4 MMAs of K=32?? (not available for bf16: ignored), 64 = S as 6 MMAs with N=256 (two key tiles per group).
// bits 128 / 256 / 512: warps 2-9 loop like the softmax warps: 128 = tcgen05.ld of 2 x 32 columns + tcgen05.st of 32,
// 256 = 64 x (fma, ex2, add) + 32 packs per thread, 512 = a shared-memory exchange + 64-thread named barrier.
// bit 2048: warp 10 streams 16 KB bulk copies (global -> shared, two in flight) into a spare shared-memory region.
__global__ void __launch_bounds__(352, 1) k_attn_mma_bench(int pairs, int mode, long long *out, const uint8_t *src) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    __shared__ float xch[256];
    __shared__ uint64_t cbar[2];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&cbar[0], 1); mbar_init(&cbar[1], 1); mbar_fence_init(); stop = 0; }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 1) {
        const int nv = (mode & 2) ? 48 : 64;
        const uint32_t idesc_s = make_idesc_bf16(128, 128), idesc_s2 = make_idesc_bf16(128, 256);
        const uint32_t idesc_o = make_idesc_bf16(128, nv, 0, (mode & 1) ? 1 : 0);
        const uint32_t k_base = smem_u32(smem), v_base = smem_u32(smem + 65536);
        uint32_t ph = 0;
        long long t_issue = 0;
        const long long t0 = clock64();
        for (int n = 0; n < pairs; ++n) {
            const long long ti = clock64();
            if (elect_one()) {
                if (!(mode & 8)) {
                    if (mode & 64) {
                        if (!(n & 1))
                            for (int i = 0; i < 12; ++i)
                                umma_bf16_ts(tb, tb + 320u + (uint32_t)(i * 8), make_smem_desc_sw128(k_base + (uint32_t)(i >> 2) * 16384u + (i & 3) * 32, 1024), idesc_s2, i ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int i = 0; i < 12; ++i)
                            umma_bf16_ts(tb + (uint32_t)((n & 1) * 128), tb + 320u + (uint32_t)(i * 8),
                                         make_smem_desc_sw128(k_base + (uint32_t)(i >> 2) * 16384u + (i & 3) * 32, 1024), idesc_s, i ? 1u : 0u);
                    }
                }
                if (!(mode & 4)) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint64_t bd = (mode & 1) ? make_smem_desc_sw128(v_base + ks * 2048, 1024) : make_smem_desc_sw128(v_base + (ks & 3) * 32, 1024);
                        umma_bf16_ts(tb + 256u, tb + (uint32_t)(((n + 1) & 1) * 128) + (uint32_t)ks * 8u, bd, idesc_o, ks ? 1u : 0u);
                    }
                }
                if ((mode & 16) || (n & 3) == 3 || n == pairs - 1) umma_commit(&bar);
            }
            __syncwarp();
            t_issue += clock64() - ti;
            if ((mode & 16) || (n & 3) == 3 || n == pairs - 1) { mbar_wait(&bar, ph); ph ^= 1u; }
        }
        const long long t1 = clock64();
        if (threadIdx.x == 32) { if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; } stop = 1; }
    } else if (warp == 10) {
        if ((mode & 2048) && (threadIdx.x & 31) == 0) {
            long long n = 0;
            const uint8_t *my = src + (size_t)blockIdx.x * (64 * 1024);
            for (; !stop; ++n) {
                const int sl = (int)(n & 1);
                if (n >= 2) mbar_wait(&cbar[sl], (uint32_t)((n >> 1) - 1) & 1u);
                mbar_expect_tx(&cbar[sl], 16384);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(smem + 96 * 1024 + sl * 16384)), "l"(my + (n & 3) * 16384), "r"(16384), "r"(smem_u32(&cbar[sl])) : "memory");
            }
            // drain
            if (n >= 1) mbar_wait(&cbar[(n - 1) & 1], (uint32_t)(((n - 1) >> 1)) & 1u);
            if (n >= 2) mbar_wait(&cbar[(n - 2) & 1], (uint32_t)(((n - 2) >> 1)) & 1u);
            if (blockIdx.x == 0) out[3] = n;
        }
    } else if (warp >= 2 && (mode & (4096 | 8192))) {
        // pollers: every lane (4096) or lane 0 only (8192) spins on a barrier that never completes
        long long n = 0;
        if ((mode & 4096) || (threadIdx.x & 31) == 0)
            while (!stop) { if (mbar_try_wait(&cbar[1], 1u ^ 1u ^ 0u)) break; ++n; }
        if (threadIdx.x == 64 && blockIdx.x == 0) out[2] = n;
    } else if (warp >= 2 && (mode & (128 | 256 | 512))) {
        const int w = warp - 2, q = warp & 3, hf = w >> 2, r = q * 32 + (threadIdx.x & 31);
        const uint32_t base = tb + ((uint32_t)(q * 32) << 16) + 384u + (uint32_t)(hf * 64);   // columns the MMAs do not touch
        uint32_t v0[32], v1[32];
        for (int e = 0; e < 32; ++e) { v0[e] = __float_as_uint(0.001f * (threadIdx.x + e)); v1[e] = __float_as_uint(0.002f * (threadIdx.x - e)); }
        float acc = 0.f;
        long long n = 0;
        // with the named barrier both warps of a pair must run the same number of rounds: fixed count that outlasts the MMAs
        for (;;) {
            if ((mode & 512) ? (n >= 2 * pairs) : (stop != 0)) break;
            if (mode & 128) {
                tmem_ld_32x32b_x32(base, v0);
                tmem_ld_32x32b_x32(base + 32, v1);
                tmem_ld_wait_dep(v0); tmem_ld_wait_dep(v1);
            }
            float mb = acc * 1e-9f;
            if (mode & 512) {
                xch[hf * 128 + r] = __uint_as_float(v0[3]);
                named_bar_sync(1 + q, 64);
                mb += xch[r] + xch[128 + r];
            }
            uint32_t pk[32];
            if (mode & 256) {
                float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p0, p1;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v0[2 * e]), 0.3f, -mb)));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v0[2 * e + 1]), 0.3f, -mb)));
                    ps0 += p0; ps1 += p1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                    pk[e] = *reinterpret_cast<uint32_t *>(&h2);
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p0, p1;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v1[2 * e]), 0.3f, -mb)));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v1[2 * e + 1]), 0.3f, -mb)));
                    ps0 += p0; ps1 += p1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                    pk[16 + e] = *reinterpret_cast<uint32_t *>(&h2);
                }
                acc += ps0 + ps1;
            } else {
                for (int e = 0; e < 32; ++e) pk[e] = v0[e] ^ v1[e];
            }
            if (mode & 128) {
                tmem_st_32x32b_x32(base, pk);
                tmem_st_wait();
            } else {
                for (int e = 0; e < 32; ++e) acc += __uint_as_float(pk[e]);
            }
            ++n;
        }
        if (acc == 12345.678f) out[7] = 1;
        if (threadIdx.x == 64 && blockIdx.x == 0) out[2] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int selftest_attn_mma_bench(int device, int mode, int what, int grid, double *cycles_per_pair) {
    const int pairs = 400;
    long long *d = nullptr, hres[8] = {0};
    uint8_t *src = nullptr;
    if (grid < 1) grid = 1;
    cudaMalloc((void **)&d, 64);
    cudaMalloc((void **)&src, (size_t)grid * 64 * 1024);
    cudaMemset(d, 0, 64);
    cudaMemset(src, 0x3c, (size_t)grid * 64 * 1024);
    cudaFuncSetAttribute(k_attn_mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    k_attn_mma_bench<<<grid, 352, 128 * 1024>>>(pairs, mode, d, src);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); return set_error(nullptr, COCR_ERR_CUDA, "attention mma bench: %s", cudaGetErrorString(e)); }
    cudaMemcpy(hres, d, 64, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaFree(src);
    if (what == 3) { if (cycles_per_pair) *cycles_per_pair = 16384.0 * (double)hres[3] / (double)hres[0]; return COCR_OK; }   // copy bytes per cycle
    if (cycles_per_pair) *cycles_per_pair = what == 2 ? (double)hres[0] / (double)(hres[2] > 0 ? hres[2] : 1) : (double)hres[what ? 1 : 0] / pairs;
    return COCR_OK;
}

// mbarrier hand-shake latency: warps 0 and 1 play ping-pong over two mbarriers (`rounds` round trips); mode 0: every lane
// waits with try_wait, 1: lane 0 waits with try_wait then __syncwarp, 2: every lane spins on test_wait, 3: lane 0 spins on
// test_wait then __syncwarp.  +16: warps 2..13 sit suspended in try_wait on a barrier that completes only at the end.
// +32: the "pong" is a tcgen05.commit instead of a plain arrive (warp 1 elects one lane).
__global__ void __launch_bounds__(448, 1) k_mbar_pingpong(int rounds, int mode, long long *out) {
    using namespace tc;
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_fence_init(); }
    __syncthreads();
    const int wm = mode & 3;
    auto wait = [&](uint64_t *b, uint32_t par) {
        if (wm == 0) mbar_wait(b, par);
        else if (wm == 1) { if (lane == 0) mbar_wait(b, par); __syncwarp(); }
        else if (wm == 2) { while (!mbar_test(b, par)) {} }
        else { if (lane == 0) { while (!mbar_test(b, par)) {} } __syncwarp(); }
    };
    if (warp == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < rounds; ++i) {
            if (lane == 0) mbar_arrive(&bars[0]);
            wait(&bars[1], i & 1);
        }
        const long long t1 = clock64();
        if (lane == 0) { out[0] = t1 - t0; mbar_arrive(&bars[2]); }
    } else if (warp == 1) {
        for (int i = 0; i < rounds; ++i) {
            wait(&bars[0], i & 1);
            if (mode & 32) { if (elect_one()) umma_commit(&bars[1]); __syncwarp(); }
            else if (lane == 0) mbar_arrive(&bars[1]);
        }
    } else if (mode & 16) {
        mbar_wait(&bars[2], 0);
    }
}

int selftest_mbar_pingpong(int device, int mode, double *cycles_per_hop) {
    const int rounds = 2000;
    long long *d = nullptr, hres = 0;
    cudaMalloc((void **)&d, 8);
    cudaMemset(d, 0, 8);
    k_mbar_pingpong<<<1, 448>>>(rounds, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); return set_error(nullptr, COCR_ERR_CUDA, "mbarrier ping-pong: %s", cudaGetErrorString(e)); }
    cudaMemcpy(&hres, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (cycles_per_hop) *cycles_per_hop = (double)hres / (2.0 * rounds);
    return COCR_OK;
}

int selftest_mma_bench(int device, int N, int chains, int a_tmem_and_count, double *cycles_per_mma) {
    const int a_tmem = a_tmem_and_count % 10;
    const int count = a_tmem_and_count >= 10 ? a_tmem_and_count / 10 : 64;     // MMAs per commit group
    if (N < 16 || N > 256 || N % 16 || chains < 1 || chains * N > 448) return set_error(nullptr, COCR_ERR_BAD_ARG, "bad microbench shape");
    long long *d = nullptr, hres = 0;
    cudaMalloc((void **)&d, 8);
    cudaFuncSetAttribute(k_mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k_mma_bench<<<1, 128, 64 * 1024>>>(N, chains, a_tmem, count, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); return set_error(nullptr, COCR_ERR_CUDA, "mma bench: %s", cudaGetErrorString(e)); }
    cudaMemcpy(&hres, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (cycles_per_mma) *cycles_per_mma = (double)hres / count;
    return COCR_OK;
}

}  // namespace cocr

// MUFU throughput: 8 warps x `iters` x 16 independent ops per thread; kind 0 ex2.f32, 1 ex2.f16x2, 2 tanh.f32, 3 tanh.f16x2,
// 4 ex2.bf16x2, 5 rcp.f32.  Reports cycles per warp-instruction per SM sub-partition (2 warps per scheduler).
namespace cocr {
template <int kind>
__global__ void __launch_bounds__(256, 1) k_mufu_bench(int iters, long long *out, float *sink) {
    float f[16];
    uint32_t u[16];
    for (int i = 0; i < 16; ++i) { f[i] = 0.001f * (threadIdx.x + i); u[i] = 0x3c003c00u + threadIdx.x + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (kind == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            else if (kind == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
            else if (kind == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
            else if (kind == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u[i]));
            else if (kind == 4) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
            else asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += f[i] + __uint_as_float(u[i]);
    if (s == 12345.678f) sink[0] = s;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

int selftest_mufu_bench(int device, int kind, double *cycles_per_inst) {
    const int iters = 2000;
    long long *d = nullptr, hres = 0;
    float *sink = nullptr;
    cudaMalloc((void **)&d, 8); cudaMalloc((void **)&sink, 4);
    switch (kind) {
        case 0: k_mufu_bench<0><<<1, 256>>>(iters, d, sink); break;
        case 1: k_mufu_bench<1><<<1, 256>>>(iters, d, sink); break;
        case 2: k_mufu_bench<2><<<1, 256>>>(iters, d, sink); break;
        case 3: k_mufu_bench<3><<<1, 256>>>(iters, d, sink); break;
        case 4: k_mufu_bench<4><<<1, 256>>>(iters, d, sink); break;
        default: k_mufu_bench<5><<<1, 256>>>(iters, d, sink); break;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); cudaFree(sink); return set_error(nullptr, COCR_ERR_CUDA, "mufu bench: %s", cudaGetErrorString(e)); }
    cudaMemcpy(&hres, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d); cudaFree(sink);
    // 8 warps on 4 schedulers: 2 warps x 16 x iters instructions per scheduler
    if (cycles_per_inst) *cycles_per_inst = (double)hres / (2.0 * 16.0 * iters);
    return COCR_OK;
}
}  // namespace cocr

extern "C" const char *cocr_microbench_error(void) { return cocr::g_mb_err; }

// which: 40 tensor-pipe cycles per MMA (p0 = N, p1 = chains, p2 = A in TMEM); 41 TMEM read bandwidth in bytes per cycle per SM
// (depth, with_mma, shape); 42 12-MMA group cycles under TMEM traffic (mode, a_tmem); 43 MUFU cycles per warp instruction
// (kind); 44 attention S + P V MMA sequence, cycles per key tile (mode, what, grid); 45 mbarrier hand-shake cycles per hop (mode)
extern "C" int cocr_microbench(int device, int which, long long p0, long long p1, long long p2, double *out) {
    if (out) *out = -1.0;
    if (cudaSetDevice(device) != cudaSuccess) return cocr::mb_error(COCR_ERR_CUDA, "cocr_microbench: bad device %d", device);
    if (which == 45) return cocr::selftest_mbar_pingpong(device, (int)p0, out);
    if (which == 44) return cocr::selftest_attn_mma_bench(device, (int)p0, (int)p1, (int)p2, out);
    if (which == 43) return cocr::selftest_mufu_bench(device, (int)p0, out);
    if (which == 42) return cocr::selftest_mma_contention(device, (int)p0, (int)p1, out);
    if (which == 41) return cocr::selftest_tmem_bench(device, (int)p0, (int)p1, (int)p2, out);
    if (which == 40) return cocr::selftest_mma_bench(device, (int)p0, (int)p1, (int)p2, out);
    return cocr::mb_error(COCR_ERR_BAD_ARG, "cocr_microbench: unknown benchmark %d", which);
}
