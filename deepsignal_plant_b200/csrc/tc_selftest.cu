// Known-answer tests of the tcgen05 building blocks (dsp_selftest, include/dsp_b200.h):
// a single-CTA GEMM D[128 x N] = A[128 x K] * B[N x K]^T with FP16 operands and FP32
// accumulation in TMEM, checked on the host against a double-precision product.
//   which = 0: A and B from shared memory (slab images moved with cp.async.bulk), N = 128, K = 64
//   which = 1: same, N = 256, K = 192 (three slabs, accumulate across slabs)
//   which = 2: A operand staged in TMEM with tcgen05.st (packed FP16 pairs), N = 128, K = 128
//   which = 3: like 2 with N = 256, K = 256, A at a non-zero TMEM column offset
#include "common.cuh"
#include "tc_prims.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>

namespace dsp {
namespace {

using namespace tc;

__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __half* __restrict__ a_slabs, const __half* __restrict__ a_rowmajor,
                const __half* __restrict__ b_slabs, float* __restrict__ d_out, int N, int KS, int a_in_tmem) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int K = KS * SLAB_K;
    uint8_t* sA = smem;                                  // KS slabs of 128 rows
    uint8_t* sB = sA + (size_t)KS * 128 * SLAB_ROW_BYTES;  // KS slabs of N rows
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_load = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    if (tid == 0) {
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t acc = tmem;                 // columns [0, N)
    const uint32_t a_t = tmem + 320;           // A operand columns (K/2 <= 128), deliberately offset

    if (tid == 0) {
        const uint32_t a_bytes = a_in_tmem ? 0u : (uint32_t)KS * 128 * SLAB_ROW_BYTES;
        const uint32_t b_bytes = (uint32_t)KS * N * SLAB_ROW_BYTES;
        mbar_arrive_expect_tx(bar_load, a_bytes + b_bytes);
        if (!a_in_tmem) bulk_g2s(smem_u32(sA), a_slabs, a_bytes, bar_load);
        bulk_g2s(smem_u32(sB), b_slabs, b_bytes, bar_load);
    }
    if (a_in_tmem) {
        // thread = TMEM lane = row; 8 packed FP16 pairs per store
        const __half* row = a_rowmajor + (size_t)tid * K;
        for (int k = 0; k < K; k += 16) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = pack_half2(__half2float(row[k + 2 * j]), __half2float(row[k + 2 * j + 1]));
            tmem_st8(a_t + ((uint32_t)(warp * 32) << 16) + (uint32_t)(k / 2), v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_wait(bar_load, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(128, N);
        for (int ks = 0; ks < KS; ++ks) {
            for (int k = 0; k < 4; ++k) {
                const uint64_t bd = make_smem_desc(smem_u32(sB + (size_t)ks * N * SLAB_ROW_BYTES) + k * 32);
                const uint32_t accum = (ks | k) ? 1u : 0u;
                if (a_in_tmem) {
                    mma_ts(acc, a_t + (uint32_t)((ks * SLAB_K + k * 16) / 2), bd, idesc, accum);
                } else {
                    const uint64_t ad = make_smem_desc(smem_u32(sA + (size_t)ks * 128 * SLAB_ROW_BYTES) + k * 32);
                    mma_ss(acc, ad, bd, idesc, accum);
                }
            }
        }
        mma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) d_out[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

void to_slabs(const std::vector<__half>& m, int rows, int K, std::vector<__half>& out) {
    const int KS = K / SLAB_K;
    out.assign((size_t)rows * K, __float2half(0.f));
    for (int ks = 0; ks < KS; ++ks)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < SLAB_K; ++c) {
                size_t off = (size_t)ks * rows * SLAB_ROW_BYTES + slab_offset_bytes(r, c);
                out[off / 2] = m[(size_t)r * K + ks * SLAB_K + c];
            }
}

}  // namespace
}  // namespace dsp

using namespace dsp;

extern "C" int dsp_selftest(int device, int which, double* max_abs_err) {
    DSP_REQUIRE(max_abs_err, DSP_ERR_INVALID, "dsp_selftest: null output");
    *max_abs_err = -1.0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("dsp_selftest: no CUDA device available");
        return DSP_ERR_CUDA;
    }
    DSP_REQUIRE(device >= 0 && device < ndev, DSP_ERR_INVALID, "dsp_selftest: bad device");
    DSP_REQUIRE(which >= 0 && which <= 3, DSP_ERR_INVALID, "dsp_selftest: unknown test %d", which);
    DSP_CUDA(cudaSetDevice(device));
    const int N = (which == 1 || which == 3) ? 256 : 128;
    const int KS = which == 0 ? 1 : which == 1 ? 3 : which == 2 ? 2 : 4;
    const int K = KS * tc::SLAB_K;
    const int a_in_tmem = which >= 2;
    std::vector<__half> A((size_t)128 * K), B((size_t)N * K), As, Bs;
    srand(1234 + which);
    for (auto& v : A) v = __float2half((float)(rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = __float2half((float)(rand() % 2001 - 1000) / 1000.f);
    to_slabs(A, 128, K, As);
    to_slabs(B, N, K, Bs);
    __half *dA, *dAr, *dB; float* dD;
    DSP_CUDA(cudaMalloc(&dA, As.size() * 2));
    DSP_CUDA(cudaMalloc(&dAr, A.size() * 2));
    DSP_CUDA(cudaMalloc(&dB, Bs.size() * 2));
    DSP_CUDA(cudaMalloc(&dD, (size_t)128 * N * 4));
    DSP_CUDA(cudaMemcpy(dA, As.data(), As.size() * 2, cudaMemcpyHostToDevice));
    DSP_CUDA(cudaMemcpy(dAr, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    DSP_CUDA(cudaMemcpy(dB, Bs.data(), Bs.size() * 2, cudaMemcpyHostToDevice));
    DSP_CUDA(cudaMemset(dD, 0xff, (size_t)128 * N * 4));
    const size_t smem = (size_t)KS * (128 + N) * tc::SLAB_ROW_BYTES + 1024;
    DSP_CUDA(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_kernel<<<1, 128, smem>>>(dA, dAr, dB, dD, N, KS, a_in_tmem);
    DSP_CUDA(cudaGetLastError());
    DSP_CUDA(cudaDeviceSynchronize());
    std::vector<float> D((size_t)128 * N);
    DSP_CUDA(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dAr); cudaFree(dB); cudaFree(dD);
    double worst = 0.0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0.0;
            for (int k = 0; k < K; ++k) ref += (double)__half2float(A[(size_t)m * K + k]) * (double)__half2float(B[(size_t)n * K + k]);
            double e = std::fabs(ref - (double)D[(size_t)m * N + n]);
            if (!(e <= worst)) worst = std::isnan(e) ? 1e30 : e;
        }
    *max_abs_err = worst;
    return DSP_OK;
}
