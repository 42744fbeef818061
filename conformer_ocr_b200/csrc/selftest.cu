// Kernel self checks callable through the C ABI (cocr_selftest).
//   which = 10 + epilogue (0 bias->bf16, 1 bias+SiLU->bf16, 2 bias->fp32, 3 fused arg-max, 6 bias+GLU->bf16): tcgen05 GEMM vs a scalar
//           device reference.  p0 = M, p1 = N, p2 = K + 100000 * block_n_override (0 = production choice).
//   which = 20: residual + LayerNorm GEMM epilogue vs a host reference (see below).
//   which = 30: batched tcgen05 GEMM of the training step (bgemm_tc.cu): p0 = M, p1 = N, p2 = K + 100000 * flags
//           (1 A stored [k][m], 2 B stored [k][n], 4 batched 3 x 2, 8 split contraction + atomic output); max abs error.
//           *max_err: bf16 outputs -> max |got-ref| / (1+|ref|); fp32 / arg-max -> max abs error.
#include "common.h"

namespace cocr {
int selftest_gemm(int device, int epi, int M, int N, int K, int block_n, double *max_err);
int selftest_resid(int device, int M, int N, int K, int flags, double *max_err);
int selftest_ffn(int device, int M, int NC, int flags, double *max_err);
}

extern "C" int cocr_selftest(int device, int which, int64_t p0, int64_t p1, int64_t p2, double *max_err) {
    if (max_err) *max_err = -1.0;
    if (cudaSetDevice(device) != cudaSuccess) return cocr::set_error(nullptr, COCR_ERR_CUDA, "cocr_selftest: bad device %d", device);
    if ((which >= 10 && which <= 13) || which == 16) {
        if (p0 < 1 || p1 < 1 || p2 < 16) return cocr::set_error(nullptr, COCR_ERR_BAD_ARG, "cocr_selftest: bad GEMM shape");
        return cocr::selftest_gemm(device, which - 10, (int)p0, (int)p1, (int)(p2 % 100000), (int)(p2 / 100000), max_err);
    }
    if (which == 20) {   // residual + LayerNorm epilogue: p0 = M, p1 = N, p2 = K + 100000 * flags (1 init, 2 final LN, 4 next LN)
        if (p0 < 1 || p1 < 16 || p2 < 16) return cocr::set_error(nullptr, COCR_ERR_BAD_ARG, "cocr_selftest: bad GEMM shape");
        return cocr::selftest_resid(device, (int)p0, (int)p1, (int)(p2 % 100000), (int)(p2 / 100000), max_err);
    }
    if (which == 30) return cocr::selftest_bgemm(device, (int)p0, (int)p1, (int)(p2 % 100000), (int)(p2 / 100000), max_err);   // batched GEMM: p2 = K + 100000 * flags
    if (which == 21) return cocr::selftest_ffn(device, (int)p0, (int)p1, (int)p2, max_err);   // fused FFN module: p0 = M, p1 = ff / 96, p2 = flags
    return cocr::set_error(nullptr, COCR_ERR_BAD_ARG, "cocr_selftest: unknown test %d", which);
}
