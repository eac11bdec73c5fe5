// tcgen05 path -- placeholder until the tensor-core kernels land.
#include "tc.cuh"
namespace dsp {
int tc_create(Model*) { set_error("DSP_PRECISION_FP16 path is not built in this revision"); return DSP_ERR_INVALID; }
void tc_destroy(Model*) {}
int tc_pack_lstm_layer(Model*, LstmLayer&, const float*, const float*, const float*, const float*,
                       const float*, const float*, const float*, const float*) { return DSP_ERR_INVALID; }
int tc_pack_dense(Model*, DenseF32&, const float*, const float*) { return DSP_ERR_INVALID; }
int tc_finalize_pack(Model*) { return DSP_ERR_INVALID; }
int tc_forward_chunk(Model*, const float*, const float*, const float*, const float*, const float*,
                     const float* const*, const float* const*, const int64_t*, int64_t, float*, float*, int32_t*,
                     cudaStream_t) { return DSP_ERR_INVALID; }
}
