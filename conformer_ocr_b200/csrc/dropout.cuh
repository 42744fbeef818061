// Dropout of the training step: Philox-4x32-10 keyed by (seed, site); used by the element-wise kernels of train_kernels.cuh.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cocr {
namespace train {

// ------------------------------------------------------------------------------------------
// dropout: Philox-4x32-10 keyed by (seed, site); element e uses word e & 3 of counter e >> 2.
// keep (scale 1/(1-p)) when the word >= thresh = p * 2^32; thresh 0 = dropout off.
// ------------------------------------------------------------------------------------------
struct Drop {
    const uint64_t *seedp;      // device word: the step's seed (read at run time so that a captured step replays with a new seed)
    uint32_t site;
    uint32_t thresh;
    float inv_keep;
};

__device__ __forceinline__ uint4 philox4(uint64_t seed, uint32_t site, uint64_t ctr) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = site, c3 = 0x5eed5eedu;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float drop_scale(const Drop &dr, uint64_t e) {
    if (dr.thresh == 0) return 1.0f;
    const uint4 r = philox4(__ldg(dr.seedp), dr.site, e >> 2);
    const uint32_t w = (e & 2) ? ((e & 1) ? r.w : r.z) : ((e & 1) ? r.y : r.x);
    return w >= dr.thresh ? dr.inv_keep : 0.0f;
}

// the factors of the 4 consecutive elements e0 .. e0+3 (e0 % 4 == 0): one Philox call
__device__ __forceinline__ float4 drop_scale4(const Drop &dr, uint64_t e0) {
    if (dr.thresh == 0) return make_float4(1.f, 1.f, 1.f, 1.f);
    const uint4 r = philox4(__ldg(dr.seedp), dr.site, e0 >> 2);
    return make_float4(r.x >= dr.thresh ? dr.inv_keep : 0.f, r.y >= dr.thresh ? dr.inv_keep : 0.f, r.z >= dr.thresh ? dr.inv_keep : 0.f,
                       r.w >= dr.thresh ? dr.inv_keep : 0.f);
}
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16 *p, float4 v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b));
}

}  // namespace train
}  // namespace cocr
