// Interface between api.cu and the tcgen05 (FP16-operand) path in kernels_tc.cu.
#pragma once
#include "common.cuh"

namespace dsp {

int tc_create(Model* m);
void tc_destroy(Model* m);
// raw torch-layout parameters of both directions of one layer (weight_ih (4H,K),
// weight_hh (4H,H), bias_ih (4H), bias_hh (4H)); d0 = forward, d1 = reverse
int tc_pack_lstm_layer(Model* m, LstmLayer& L,
                       const float* wih0, const float* whh0, const float* bih0, const float* bhh0,
                       const float* wih1, const float* whh1, const float* bih1, const float* bhh1,
                       int split_first = 0, int split_n = 0, int len_col = -1);
// true when the first lstm_seq layer takes its scalar features as FP16 value + residual columns
bool tc_seq_split(const Model* m);
int tc_pack_dense(Model* m, DenseF32& D, const float* w, const float* b);
int tc_pack_head(Model* m, DenseF32& fc1, const float* w, const float* b);
void tc_drop_packs(Model* m);      // forget the packed layers before a re-pack (their device memory is freed by the caller)
int tc_forward_chunk(Model* m, const float* kmer, const float* means, const float* stds, const float* lens,
                     const float* signals, const float* const* h0, const float* const* c0,
                     const int64_t* state_stride, uint64_t seed, uint64_t chunk_id, int64_t n,
                     float* logits, float* probs, int32_t* labels, cudaStream_t st);
// h0 == nullptr: the layer kernels draw N(0,1) initial states themselves (Philox keyed by seed)

}  // namespace dsp
