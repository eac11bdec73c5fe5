// call_freq aggregation -- placeholder until the segmented-replay kernels land.
#include "common.cuh"
extern "C" int dsp_freq_aggregate(int, const uint64_t*, const double*, const double*, const int32_t*, int64_t, double, int,
                                  uint64_t*, int64_t*, double*, double*, int32_t*, int32_t*, int32_t*, int64_t*, void*) {
    dsp::set_error("dsp_freq_aggregate is not built in this revision");
    return DSP_ERR_INVALID;
}
