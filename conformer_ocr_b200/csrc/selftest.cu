// Kernel self checks callable through the C ABI (cocr_selftest); filled in per kernel family.
#include "common.h"

extern "C" int cocr_selftest(int device, int which, int64_t p0, int64_t p1, int64_t p2, double *max_err) {
    (void)device; (void)p0; (void)p1; (void)p2;
    if (max_err) *max_err = 0.0;
    return cocr::set_error(nullptr, COCR_ERR_BAD_ARG, "cocr_selftest: unknown test %d", which);
}
