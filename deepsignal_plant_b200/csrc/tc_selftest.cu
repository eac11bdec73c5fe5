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

// ---- MMA issue-rate probe ---------------------------------------------------------------------
// One CTA per SM; one thread issues `iters` groups of 8 MMAs (operands resident, zero-filled),
// optionally while another warp streams 32 KB bulk copies into a different part of shared
// memory and/or four warps keep reading TMEM.  Reports SM cycles per MMA.
//   mode bit0: A from TMEM (TS) instead of shared memory (SS); bit1: N = 256 instead of 128;
//   bit2: concurrent bulk copies; bit3: concurrent tcgen05.ld traffic
__global__ void __launch_bounds__(256, 1)
mma_rate_kernel(const uint8_t* __restrict__ src, int iters, int mode, unsigned long long* __restrict__ cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // 16 KB
    uint8_t* sB = smem + 16384;               // 32 KB (N up to 256)
    uint8_t* sL = smem + 49152;               // 2 x 32 KB landing zone for the concurrent copies
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop_flag;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
        mbar_fence_init();
        stop_flag = 0;
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // mode bits 4-5: number of independent accumulator chains the MMAs rotate over (1,2,4);
    // bit 6: N = 64
    const int N = (mode & 64) ? 64 : (mode & 2) ? 256 : 128;
    const int chains = 1 << ((mode >> 4) & 3);
    const uint32_t idesc = make_idesc_f16(128, N);
    if (warp == 0) {
        // warp-uniform loop, one elected lane issues (same structure as the layer kernels)
        const bool leader = elect_one();
        const uint32_t a_lo = smem_desc_lo(smem_u32(sA)), b_lo = smem_desc_lo(smem_u32(sB));
        const uint32_t d0 = tmem, d1 = tmem + (uint32_t)((chains > 1) ? N : 0);
        const long long t0 = clock64();
        if (mode & 1) {
            for (int it = 0; it < iters; ++it) {
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma_ts_lo<1>((k & 1) ? d1 : d0, tmem + 384 + (k & 3) * 8, b_lo + (k & 3) * 2, idesc);
                }
                __syncwarp();
            }
        } else {
            for (int it = 0; it < iters; ++it) {
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma_ss_lo<1>((k & 1) ? d1 : d0, a_lo + (k & 3) * 2, b_lo + (k & 3) * 2, idesc);
                }
                __syncwarp();
            }
        }
        if (leader) mma_commit(smem_u32(&bars[0]));
        __syncwarp();
        mbar_wait(smem_u32(&bars[0]), 0);
        const long long t1 = clock64();
        if (leader) { cycles_out[blockIdx.x] = (unsigned long long)(t1 - t0); stop_flag = 1; }
    } else if (warp == 1 && (mode & 4)) {
        if (tid == 32) {
            uint32_t ph[2] = {0, 0};
            int buf = 0;
            size_t off = (size_t)blockIdx.x * 65536;
            while (!stop_flag) {
                const uint32_t bar = smem_u32(&bars[1 + buf]);
                mbar_arrive_expect_tx(bar, 32768);
                bulk_g2s(smem_u32(sL + buf * 32768), src + (off & ((64u << 20) - 1)), 32768, bar);
                off += 32768;
                if (buf == 1) {   // keep two copies in flight: wait for the older one
                    mbar_wait(smem_u32(&bars[1]), ph[0]); ph[0] ^= 1;
                    mbar_wait(smem_u32(&bars[2]), ph[1]); ph[1] ^= 1;
                }
                buf ^= 1;
            }
            if (buf == 1) { mbar_wait(smem_u32(&bars[1]), ph[0]); }
        }
    } else if (warp >= 4 && (mode & 8)) {
        uint32_t acc = 0;
        while (!stop_flag) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256, v);
            tmem_ld_wait();
            acc += v[0] + v[31];
        }
        if (acc == 0x12345678u) cycles_out[0] = 0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- pipe-overlap probe ----------------------------------------------------------------------------
// 16 warps per SM run `iters` iterations of 8 independent chains of: bit0 MUFU.EX2, bit1 packed FFMA2,
// bit2 scalar FFMA (x2, same FLOPs as one FFMA2), bit3 MUFU.RCP.  Reports SM cycles per iteration per
// warp-instruction slot, so that the cost of a mix can be compared with the sum / max of its parts.
__global__ void __launch_bounds__(512, 1)
pipe_probe_kernel(int iters, int mode, float seed, unsigned long long* __restrict__ cycles_out, float* __restrict__ sink) {
    float2 a[8];
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i, seed - i); e[i] = seed * 0.001f * (i + 1); }
    const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (mode & 1) e[i] = ex2_approx(e[i]);
            if (mode & 8) e[i] = rcp_approx(e[i]);
            if (mode & 2) a[i] = fma2(a[i], m, c);
            if (mode & 4) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + e[i];
    if (acc == 12345.678f) sink[0] = acc;
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = (unsigned long long)(t1 - t0);
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
    DSP_REQUIRE((which >= 0 && which <= 3) || (which >= 100 && which < 228) || (which >= 300 && which < 316), DSP_ERR_INVALID,
                "dsp_selftest: unknown test %d", which);
    DSP_CUDA(cudaSetDevice(device));
    if (which >= 300) {
        // pipe-overlap probe: returns SM cycles per loop iteration (8 chains x the selected instructions, 16 warps)
        const int mode = which - 300, iters = 4000;
        int nsm = 0;
        DSP_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
        unsigned long long* cyc; float* sink;
        DSP_CUDA(cudaMalloc(&cyc, sizeof(unsigned long long) * nsm));
        DSP_CUDA(cudaMalloc(&sink, sizeof(float)));
        for (int rep = 0; rep < 2; ++rep) {
            pipe_probe_kernel<<<nsm, 512>>>(iters, mode, 0.5f, cyc, sink);
            DSP_CUDA(cudaGetLastError());
            DSP_CUDA(cudaDeviceSynchronize());
        }
        std::vector<unsigned long long> h(nsm);
        DSP_CUDA(cudaMemcpy(h.data(), cyc, sizeof(unsigned long long) * nsm, cudaMemcpyDeviceToHost));
        cudaFree(cyc); cudaFree(sink);
        double sum = 0;
        for (auto v : h) sum += (double)v;
        *max_abs_err = sum / nsm / (double)iters;
        return DSP_OK;
    }
    if (which >= 100) {
        // MMA rate probe: returns average SM cycles per tcgen05.mma over all SMs
        const int mode = which - 100, iters = 2000;
        int nsm = 0;
        DSP_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
        uint8_t* src; unsigned long long* cyc;
        DSP_CUDA(cudaMalloc(&src, (size_t)64 << 20));
        DSP_CUDA(cudaMemset(src, 0, (size_t)64 << 20));
        DSP_CUDA(cudaMalloc(&cyc, sizeof(unsigned long long) * nsm));
        const size_t smem = 49152 + 65536 + 1024;
        DSP_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int rep = 0; rep < 2; ++rep) {
            mma_rate_kernel<<<nsm, 256, smem>>>(src, iters, mode, cyc);
            DSP_CUDA(cudaGetLastError());
            DSP_CUDA(cudaDeviceSynchronize());
        }
        std::vector<unsigned long long> h(nsm);
        DSP_CUDA(cudaMemcpy(h.data(), cyc, sizeof(unsigned long long) * nsm, cudaMemcpyDeviceToHost));
        cudaFree(src); cudaFree(cyc);
        double sum = 0;
        for (auto v : h) sum += (double)v;
        *max_abs_err = sum / nsm / ((double)iters * 8);
        return DSP_OK;
    }
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
