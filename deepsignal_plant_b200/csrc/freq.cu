// call_freq per-site aggregation on the GPU (dsp_freq_aggregate, include/dsp_b200.h).
//
// Reference semantics (deepsignal_plant/call_mods_freq.py:29-74, utils/txt_formater.py:8-26):
// records are visited in file order; a record is dropped when |p0 - p1| < prob_cf; per site
// key the probabilities are accumulated in float64 left to right, counts are integers, and
// the first callable record of a key supplies its metadata.  Floating-point addition is
// not associative, so this is NOT a tree reduction: records are brought together with a
// STABLE radix sort on the key (equal keys keep file order) and each key's run is replayed
// sequentially by one thread -- bit-identical sums to the reference's Python loop.
//
// HBM-bound integer/byte work: compaction, one 64-bit-key radix sort of (key, index) pairs
// restricted to the significant key bits, run-length encode, replay, optional re-order by
// first appearance.
#include "common.cuh"
#include <cub/cub.cuh>
#include <mutex>

namespace dsp {
namespace {

// Scratch memory comes from a small per-process cache of device blocks (grow-only, best fit):
// a call makes ~20 allocations, and cudaMalloc/cudaFree around every call cost several times the
// kernels themselves.  dsp_freq_release_cache() returns the blocks to the driver.
struct Block { void* p; size_t bytes; int device; bool busy; };
std::vector<Block> g_blocks;
std::mutex g_blocks_mu;

struct Scratch {
    std::vector<size_t> held;
    cudaStream_t st;
    int device;
    Scratch(cudaStream_t s, int dev) : st(s), device(dev) {}
    ~Scratch() {
        std::lock_guard<std::mutex> lk(g_blocks_mu);
        for (size_t i : held) g_blocks[i].busy = false;
    }
    template <typename T> int alloc(T** p, size_t count) {
        const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
        std::lock_guard<std::mutex> lk(g_blocks_mu);
        size_t best = (size_t)-1;
        for (size_t i = 0; i < g_blocks.size(); ++i)
            if (!g_blocks[i].busy && g_blocks[i].device == device && g_blocks[i].bytes >= bytes &&
                (best == (size_t)-1 || g_blocks[i].bytes < g_blocks[best].bytes)) best = i;
        if (best == (size_t)-1 || g_blocks[best].bytes > 2 * bytes + (1 << 20)) {
            void* q = nullptr;
            cudaError_t e = cudaMalloc(&q, bytes);
            if (e != cudaSuccess) { set_error("dsp_freq_aggregate: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return DSP_ERR_NOMEM; }
            g_blocks.push_back({q, bytes, device, false});
            best = g_blocks.size() - 1;
        }
        g_blocks[best].busy = true;
        held.push_back(best);
        *p = (T*)g_blocks[best].p;
        return DSP_OK;
    }
};

__global__ void callable_flags_kernel(const double* __restrict__ p0, const double* __restrict__ p1, int64_t n,
                                      double prob_cf, uint8_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = !(fabs(p0[i] - p1[i]) < prob_cf);     // txt_formater.py:23-26
}

__global__ void gather_keys_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ idx, int64_t m,
                                   uint64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = key[idx[i]];
}

// one thread per site: sequential float64 replay of its records in file order
__global__ void replay_kernel(const uint32_t* __restrict__ sorted_idx, const int64_t* __restrict__ offsets,
                              const int32_t* __restrict__ counts, int64_t nseg,
                              const double* __restrict__ p0, const double* __restrict__ p1,
                              const int32_t* __restrict__ label,
                              uint32_t* __restrict__ first, double* __restrict__ s0, double* __restrict__ s1,
                              int32_t* __restrict__ met, int32_t* __restrict__ unmet) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const int64_t off = offsets[s];
    const int32_t cnt = counts[s];
    double a0 = 0.0, a1 = 0.0;
    int32_t m1 = 0, m0 = 0;
    for (int32_t j = 0; j < cnt; ++j) {
        const uint32_t r = sorted_idx[off + j];
        a0 = __dadd_rn(a0, p0[r]);      // call_mods_freq.py:60-61, float64, file order
        a1 = __dadd_rn(a1, p1[r]);
        if (label[r] == 1) ++m1; else ++m0;
    }
    first[s] = sorted_idx[off];
    s0[s] = a0; s1[s] = a1; met[s] = m1; unmet[s] = m0;
}

__global__ void iota_kernel(uint32_t* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

__global__ void emit_kernel(const uint32_t* __restrict__ perm, int64_t nseg, const uint64_t* __restrict__ ukey,
                            const uint32_t* __restrict__ first, const double* __restrict__ s0,
                            const double* __restrict__ s1, const int32_t* __restrict__ met,
                            const int32_t* __restrict__ unmet, const int32_t* __restrict__ counts,
                            uint64_t* __restrict__ out_key, int64_t* __restrict__ out_first,
                            double* __restrict__ out_p0, double* __restrict__ out_p1,
                            int32_t* __restrict__ out_met, int32_t* __restrict__ out_unmet,
                            int32_t* __restrict__ out_cov) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    const uint32_t s = perm ? perm[i] : (uint32_t)i;
    out_key[i] = ukey[s];
    out_first[i] = (int64_t)first[s];
    out_p0[i] = s0[s];
    out_p1[i] = s1[s];
    out_met[i] = met[s];
    out_unmet[i] = unmet[s];
    out_cov[i] = counts[s];
}

inline unsigned blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace
}  // namespace dsp

using namespace dsp;

extern "C" int dsp_freq_aggregate(int device, const uint64_t* key, const double* p0, const double* p1,
                                  const int32_t* label, int64_t n, double prob_cf, int sort_by_key,
                                  uint64_t* out_key, int64_t* out_first, double* out_p0, double* out_p1,
                                  int32_t* out_met, int32_t* out_unmet, int32_t* out_cov,
                                  int64_t* n_sites_host, void* stream) {
    DSP_REQUIRE(n_sites_host, DSP_ERR_INVALID, "dsp_freq_aggregate: n_sites_host is null");
    *n_sites_host = 0;
    DSP_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, DSP_ERR_INVALID,
                "dsp_freq_aggregate: n=%lld out of range (shard the records: < 2^31 per call)", (long long)n);
    if (n == 0) return DSP_OK;
    DSP_REQUIRE(key && p0 && p1 && label && out_key && out_first && out_p0 && out_p1 && out_met && out_unmet && out_cov,
                DSP_ERR_INVALID, "dsp_freq_aggregate: null pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("dsp_freq_aggregate: no CUDA device available; this library has no CPU path");
        return DSP_ERR_CUDA;
    }
    DSP_REQUIRE(device >= 0 && device < ndev, DSP_ERR_INVALID, "dsp_freq_aggregate: bad device %d", device);
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    struct Restore { int p, d; ~Restore() { if (p != d && p >= 0) cudaSetDevice(p); } } restore{prev, device};
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(st, device);
    int rc;

    // 1. callable filter -> compacted record indices (file order preserved)
    uint8_t* flag; uint32_t* idx; int64_t* d_m;
    if ((rc = sc.alloc(&flag, n)) || (rc = sc.alloc(&idx, n)) || (rc = sc.alloc(&d_m, 2))) return rc;
    callable_flags_kernel<<<blocks(n), 256, 0, st>>>(p0, p1, n, prob_cf, flag);
    DSP_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    cub::CountingInputIterator<uint32_t> counting(0);
    DSP_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, counting, flag, idx, d_m, (int)n, st));
    void* tmp; size_t tmp_cap = tmp_bytes;
    if ((rc = sc.alloc((uint8_t**)&tmp, tmp_cap))) return rc;
    DSP_CUDA(cub::DeviceSelect::Flagged(tmp, tmp_bytes, counting, flag, idx, d_m, (int)n, st));
    int64_t m = 0;
    DSP_CUDA(cudaMemcpyAsync(&m, d_m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DSP_CUDA(cudaStreamSynchronize(st));
    if (m == 0) return DSP_OK;

    // 2. stable sort of (key, index) on the significant key bits
    uint64_t *kc, *ks; uint32_t* is; uint64_t* d_max;
    if ((rc = sc.alloc(&kc, m)) || (rc = sc.alloc(&ks, m)) || (rc = sc.alloc(&is, m)) || (rc = sc.alloc(&d_max, 1))) return rc;
    gather_keys_kernel<<<blocks(m), 256, 0, st>>>(key, idx, m, kc);
    DSP_CUDA(cudaGetLastError());
    auto ensure_tmp = [&](size_t need) -> int {
        if (need <= tmp_cap) return DSP_OK;
        tmp_cap = need;
        return sc.alloc((uint8_t**)&tmp, tmp_cap);
    };
    DSP_CUDA(cub::DeviceReduce::Max(nullptr, tmp_bytes, kc, d_max, (int)m, st));
    if ((rc = ensure_tmp(tmp_bytes))) return rc;
    DSP_CUDA(cub::DeviceReduce::Max(tmp, tmp_bytes, kc, d_max, (int)m, st));
    uint64_t kmax = 0;
    DSP_CUDA(cudaMemcpyAsync(&kmax, d_max, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    DSP_CUDA(cudaStreamSynchronize(st));
    int end_bit = 1;
    while (end_bit < 64 && (kmax >> end_bit)) ++end_bit;
    DSP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kc, ks, idx, is, (int)m, 0, end_bit, st));
    if ((rc = ensure_tmp(tmp_bytes))) return rc;
    DSP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kc, ks, idx, is, (int)m, 0, end_bit, st));

    // 3. runs of equal keys = sites
    uint64_t* ukey; int32_t* counts; int64_t* offsets; int32_t* d_runs;
    if ((rc = sc.alloc(&ukey, m)) || (rc = sc.alloc(&counts, m)) || (rc = sc.alloc(&offsets, m)) || (rc = sc.alloc(&d_runs, 1))) return rc;
    DSP_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, ks, ukey, counts, d_runs, (int)m, st));
    if ((rc = ensure_tmp(tmp_bytes))) return rc;
    DSP_CUDA(cub::DeviceRunLengthEncode::Encode(tmp, tmp_bytes, ks, ukey, counts, d_runs, (int)m, st));
    int32_t nseg32 = 0;
    DSP_CUDA(cudaMemcpyAsync(&nseg32, d_runs, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DSP_CUDA(cudaStreamSynchronize(st));
    const int64_t nseg = nseg32;
    DSP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offsets, (int)nseg, st));
    if ((rc = ensure_tmp(tmp_bytes))) return rc;
    DSP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, offsets, (int)nseg, st));

    // 4. ordered float64 replay per site
    uint32_t* first; double *s0, *s1; int32_t *met, *unmet;
    if ((rc = sc.alloc(&first, nseg)) || (rc = sc.alloc(&s0, nseg)) || (rc = sc.alloc(&s1, nseg)) ||
        (rc = sc.alloc(&met, nseg)) || (rc = sc.alloc(&unmet, nseg))) return rc;
    replay_kernel<<<blocks(nseg), 256, 0, st>>>(is, offsets, counts, nseg, p0, p1, label, first, s0, s1, met, unmet);
    DSP_CUDA(cudaGetLastError());

    // 5. output order: by key (already) or by first callable appearance (dict insertion order)
    uint32_t* perm = nullptr;
    if (!sort_by_key) {
        uint32_t *seg_ids, *first_sorted;
        if ((rc = sc.alloc(&seg_ids, nseg)) || (rc = sc.alloc(&perm, nseg)) || (rc = sc.alloc(&first_sorted, nseg))) return rc;
        iota_kernel<<<blocks(nseg), 256, 0, st>>>(seg_ids, nseg);
        DSP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, first, first_sorted, seg_ids, perm, (int)nseg, 0, 32, st));
        if ((rc = ensure_tmp(tmp_bytes))) return rc;
        DSP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, first, first_sorted, seg_ids, perm, (int)nseg, 0, 32, st));
    }
    emit_kernel<<<blocks(nseg), 256, 0, st>>>(perm, nseg, ukey, first, s0, s1, met, unmet, counts,
                                             out_key, out_first, out_p0, out_p1, out_met, out_unmet, out_cov);
    DSP_CUDA(cudaGetLastError());
    DSP_CUDA(cudaStreamSynchronize(st));
    *n_sites_host = nseg;
    return DSP_OK;
}

extern "C" int dsp_freq_release_cache(void) {
    std::lock_guard<std::mutex> lk(g_blocks_mu);
    std::vector<Block> keep;
    int prev = -1;
    cudaGetDevice(&prev);
    for (Block& b : g_blocks) {
        if (b.busy) { keep.push_back(b); continue; }
        cudaSetDevice(b.device);
        cudaFree(b.p);
    }
    if (prev >= 0) cudaSetDevice(prev);
    g_blocks.swap(keep);
    cudaGetLastError();
    return DSP_OK;
}
