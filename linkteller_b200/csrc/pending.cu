// Entry points declared in include/linkteller_b200.h whose kernels have not landed yet: they fail
// loudly (LT_ERR_INTERNAL) instead of silently falling back to anything.
#include "common.cuh"

#define LT_PENDING(name)                                     \
    do {                                                     \
        lt::set_error(name ": kernel not implemented yet");  \
        return LT_ERR_INTERNAL;                              \
    } while (0)

extern "C" int lt_dp_edgerand_apply(const int32_t*, const int32_t*, const double*, int32_t, const int64_t*,
                                    const uint8_t*, int64_t, int32_t*, int32_t*, double*, int64_t, int64_t*,
                                    lt_stream_t) { LT_PENDING("lt_dp_edgerand_apply"); }
extern "C" int lt_dp_lapgraph_select(const int32_t*, const int32_t*, const double*, int32_t, const double*, int64_t,
                                     int32_t*, int32_t*, float*, int64_t, int64_t*, lt_stream_t) {
    LT_PENDING("lt_dp_lapgraph_select");
}
