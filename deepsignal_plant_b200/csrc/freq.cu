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
// HBM-bound integer/byte work: callable records are compacted (file order kept) and packed into aligned
// 32-byte structs, the sort key is the site key with its never-set bits squeezed out (order
// preserving; 32-bit keys when they fit), one stable radix sort of (key, position) pairs on exactly the bits
// in use, run-length encode, replay (one 32-byte sector per record), optional re-order by first appearance.
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

// A callable record, packed: one aligned 32-byte sector holds everything the replay needs, so the gather by
// sorted position costs one sector per record instead of three scattered ones.
struct __align__(32) Rec { uint64_t key; double p0, p1; uint32_t idx; int32_t label; };

__global__ void callable_flags_kernel(const double* __restrict__ p0, const double* __restrict__ p1, int64_t n,
                                      double prob_cf, uint8_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = !(fabs(p0[i] - p1[i]) < prob_cf);     // txt_formater.py:23-26
}

// Which key bits are ever set (bitwise OR of all keys): the sort key is the key with every never-set bit
// squeezed out (an order-preserving "parallel bit extract"), which for chrom<<40|pos keys turns a 43-bit sort
// into a ~22-bit one.  The set bits are described as up to MAX_RUNS runs of consecutive ones.
constexpr int MAX_RUNS = 8;
struct BitRuns {
    int n, bits;
    uint8_t lo[MAX_RUNS], len[MAX_RUNS], out_lo[MAX_RUNS];
    __host__ __device__ __forceinline__ uint64_t squeeze(uint64_t k) const {
        uint64_t r = 0;
        for (int i = 0; i < n; ++i) r |= ((k >> lo[i]) & ((len[i] >= 64 ? 0 : (1ull << len[i])) - 1ull)) << out_lo[i];
        return r;
    }
    __host__ __device__ __forceinline__ uint64_t spread(uint64_t k) const {
        uint64_t r = 0;
        for (int i = 0; i < n; ++i) r |= ((k >> out_lo[i]) & ((len[i] >= 64 ? 0 : (1ull << len[i])) - 1ull)) << lo[i];
        return r;
    }
};

__global__ void key_bits_kernel(const uint64_t* __restrict__ key, int64_t n, unsigned long long* __restrict__ out) {
    unsigned lo = 0, hi = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = key[i];
        lo |= (unsigned)k; hi |= (unsigned)(k >> 32);
    }
    lo = __reduce_or_sync(0xffffffffu, lo); hi = __reduce_or_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0 && (lo | hi)) atomicOr(out, ((unsigned long long)hi << 32) | lo);
}

// Callable record c (c-th in file order, original index idx[c]) -> its packed Rec, its squeezed sort key and its
// position.  idx is increasing, so the four input streams are read almost sequentially.
template <typename K>
__global__ void pack_kernel(const uint32_t* __restrict__ idx, int64_t m, const uint64_t* __restrict__ key,
                            const double* __restrict__ p0, const double* __restrict__ p1, const int32_t* __restrict__ label,
                            BitRuns runs, Rec* __restrict__ rec, K* __restrict__ kc, uint32_t* __restrict__ pc) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const uint32_t i = idx[c];
    Rec r; r.key = key[i]; r.p0 = p0[i]; r.p1 = p1[i]; r.idx = i; r.label = label[i];
    rec[c] = r;
    kc[c] = (K)runs.squeeze(r.key);
    pc[c] = (uint32_t)c;
}

// one thread per site: sequential float64 replay of its records in file order
__global__ void replay_kernel(const uint32_t* __restrict__ sorted_pos, const int64_t* __restrict__ offsets,
                              const int32_t* __restrict__ counts, int64_t nseg, const Rec* __restrict__ rec,
                              uint32_t* __restrict__ first, double* __restrict__ s0, double* __restrict__ s1,
                              int32_t* __restrict__ met, int32_t* __restrict__ unmet) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const int64_t off = offsets[s];
    const int32_t cnt = counts[s];
    double a0 = 0.0, a1 = 0.0;
    int32_t m1 = 0, m0 = 0;
    uint32_t f = 0;
    for (int32_t j = 0; j < cnt; ++j) {
        const Rec r = rec[sorted_pos[off + j]];
        if (j == 0) f = r.idx;
        a0 = __dadd_rn(a0, r.p0);       // call_mods_freq.py:60-61, float64, file order
        a1 = __dadd_rn(a1, r.p1);
        if (r.label == 1) ++m1; else ++m0;
    }
    first[s] = f;
    s0[s] = a0; s1[s] = a1; met[s] = m1; unmet[s] = m0;
}

__global__ void iota_kernel(uint32_t* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

template <typename K>
__global__ void emit_kernel(const uint32_t* __restrict__ perm, int64_t nseg, const K* __restrict__ ukey, BitRuns runs,
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
    out_key[i] = runs.spread((uint64_t)ukey[s]);
    out_first[i] = (int64_t)first[s];
    out_p0[i] = s0[s];
    out_p1[i] = s1[s];
    out_met[i] = met[s];
    out_unmet[i] = unmet[s];
    out_cov[i] = counts[s];
}

inline unsigned blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

// Steps 2-5 on squeezed sort keys of type K (32-bit when the keys in use fit, else 64-bit).
template <typename K>
int aggregate_sorted(Scratch& sc, cudaStream_t st, const uint32_t* idx, int64_t m, const uint64_t* key, const double* p0,
                     const double* p1, const int32_t* label, const BitRuns& runs, int sort_by_key,
                     void* tmp, size_t tmp_cap,
                     uint64_t* out_key, int64_t* out_first, double* out_p0, double* out_p1,
                     int32_t* out_met, int32_t* out_unmet, int32_t* out_cov, int64_t* n_sites_host) {
    int rc;
    size_t tmp_bytes = 0;
    auto ensure_tmp = [&](size_t need) -> int {
        if (need <= tmp_cap) return DSP_OK;
        tmp_cap = need;
        return sc.alloc((uint8_t**)&tmp, tmp_cap);
    };
    // 2. stable sort of (squeezed key, position in rec) on the bits in use: equal keys keep file order
    K *kc, *ks; uint32_t *pc, *ps; Rec* rec;
    if ((rc = sc.alloc(&rec, m)) || (rc = sc.alloc(&kc, m)) || (rc = sc.alloc(&ks, m)) || (rc = sc.alloc(&pc, m)) || (rc = sc.alloc(&ps, m))) return rc;
    pack_kernel<K><<<blocks(m), 256, 0, st>>>(idx, m, key, p0, p1, label, runs, rec, kc, pc);
    DSP_CUDA(cudaGetLastError());
    DSP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kc, ks, pc, ps, (int)m, 0, runs.bits, st));
    if ((rc = ensure_tmp(tmp_bytes))) return rc;
    DSP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kc, ks, pc, ps, (int)m, 0, runs.bits, st));

    // 3. runs of equal keys = sites
    K* ukey; int32_t* counts; int64_t* offsets; int32_t* d_runs;
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
    replay_kernel<<<blocks(nseg), 256, 0, st>>>(ps, offsets, counts, nseg, rec, first, s0, s1, met, unmet);
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
    emit_kernel<K><<<blocks(nseg), 256, 0, st>>>(perm, nseg, ukey, runs, first, s0, s1, met, unmet, counts,
                                                out_key, out_first, out_p0, out_p1, out_met, out_unmet, out_cov);
    DSP_CUDA(cudaGetLastError());
    DSP_CUDA(cudaStreamSynchronize(st));
    *n_sites_host = nseg;
    return DSP_OK;
}

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

    // 0. which key bits are in use
    unsigned long long* d_bits;
    if ((rc = sc.alloc(&d_bits, 1))) return rc;
    DSP_CUDA(cudaMemsetAsync(d_bits, 0, sizeof(unsigned long long), st));
    key_bits_kernel<<<148 * 8, 256, 0, st>>>(key, n, d_bits);
    DSP_CUDA(cudaGetLastError());

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
    unsigned long long bits_used = 0;
    DSP_CUDA(cudaMemcpyAsync(&m, d_m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DSP_CUDA(cudaMemcpyAsync(&bits_used, d_bits, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    DSP_CUDA(cudaStreamSynchronize(st));
    if (m == 0) return DSP_OK;
    BitRuns runs{};
    {
        int out = 0;
        bool ok = true;
        for (int b = 0; b < 64 && ok;) {
            if (!((bits_used >> b) & 1ull)) { ++b; continue; }
            int e = b;
            while (e < 64 && ((bits_used >> e) & 1ull)) ++e;
            if (runs.n == MAX_RUNS) { ok = false; break; }
            runs.lo[runs.n] = (uint8_t)b; runs.len[runs.n] = (uint8_t)(e - b); runs.out_lo[runs.n] = (uint8_t)out;
            out += e - b; ++runs.n; b = e;
        }
        if (!ok) {                                   // too fragmented: sort on the key as it is
            int top = 64; while (top > 1 && !((bits_used >> (top - 1)) & 1ull)) --top;
            runs.n = 1; runs.lo[0] = 0; runs.len[0] = (uint8_t)top; runs.out_lo[0] = 0; out = top;
        }
        if (out == 0) { runs.n = 1; runs.lo[0] = 0; runs.len[0] = 1; runs.out_lo[0] = 0; out = 1; }   // every key is 0
        runs.bits = out;
    }
    return runs.bits <= 32
        ? aggregate_sorted<uint32_t>(sc, st, idx, m, key, p0, p1, label, runs, sort_by_key, tmp, tmp_cap, out_key, out_first, out_p0, out_p1, out_met, out_unmet, out_cov, n_sites_host)
        : aggregate_sorted<uint64_t>(sc, st, idx, m, key, p0, p1, label, runs, sort_by_key, tmp, tmp_cap, out_key, out_first, out_p0, out_p1, out_met, out_unmet, out_cov, n_sites_host);
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
