// bf16 tensor-core path (placeholder until the tcgen05 kernels land in this file).
#include "common.h"

namespace cocr {

struct Bf16State { int dummy; };

int bf16_init(cocr_handle *h) { return set_error(h, COCR_ERR_UNSUPPORTED, "bf16 path not built yet"); }
void bf16_destroy(cocr_handle *h) { delete h->bf; h->bf = nullptr; }
int bf16_encode(cocr_handle *h, const float *, int64_t, int64_t, int64_t, const int32_t *, int, int, cudaStream_t, float *, int32_t *) {
    return set_error(h, COCR_ERR_UNSUPPORTED, "bf16 path not built yet");
}
size_t bf16_workspace_bytes(const cocr_handle *, int, int) { return 0; }
int bf16_forward_decode(cocr_handle *h, const float *, int64_t, int64_t, int64_t, const int32_t *, int, int, cudaStream_t,
                        int32_t *, int32_t *, int32_t *, float *, int32_t *, int32_t *) {
    return set_error(h, COCR_ERR_UNSUPPORTED, "bf16 path not built yet");
}
int bf16_logits(cocr_handle *h, const float *, int64_t, int64_t, int64_t, const int32_t *, int, int, cudaStream_t, float *, int32_t *) {
    return set_error(h, COCR_ERR_UNSUPPORTED, "bf16 path not built yet");
}

}  // namespace cocr
