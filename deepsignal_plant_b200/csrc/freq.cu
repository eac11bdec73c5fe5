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
// HBM-bound integer/byte work, in five steps:
//   1. pack_callable     callable filter + stable compaction into aligned 32-byte Rec (route.cuh with one
//                        destination: count per block, scan, scatter), OR of the keys in use on the way;
//   2. sort              sort key = site key with its never-set bits squeezed out (order preserving; 32-bit
//                        keys when they fit); one stable CUB radix sort of (key, position) on exactly the bits
//                        in use;
//   3. run-length encode + exclusive scan = the sites and where their records are;
//   4. replay            a block of 256 sites stages the records of its span through shared memory with
//                        coalesced index reads and one 32-byte sector per record, then every thread adds
//                        its own site's records in file order (__dadd_rn, no contraction);
//   5. order + emit      rows by first callable appearance (dict insertion order) or by key.
// Steps 2-4 (sites_from_records) are shared with the multi-GPU path (comm.cu), whose records arrive
// already packed through the NVLink exchange.
#include "freq.cuh"
#include "route.cuh"
#include <cub/cub.cuh>
#include <mutex>
#include <list>
#include <cstdlib>

namespace dsp {

// ---- scratch cache --------------------------------------------------------------------------------
namespace {
struct Block { void* p; size_t bytes; int device; bool busy; };
std::list<Block> g_blocks;                       // a list: live Scratch objects hold pointers into it
std::mutex g_blocks_mu;
}  // namespace

Scratch::~Scratch() {
    std::lock_guard<std::mutex> lk(g_blocks_mu);
    for (void* h : held) static_cast<Block*>(h)->busy = false;
}

int Scratch::alloc_bytes(void** p, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    std::lock_guard<std::mutex> lk(g_blocks_mu);
    Block* best = nullptr;
    for (Block& b : g_blocks)
        if (!b.busy && b.device == device && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
    if (!best || best->bytes > 2 * bytes + (1 << 20)) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) { set_error("call_freq: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return DSP_ERR_NOMEM; }
        g_blocks.push_back({q, bytes, device, false});
        best = &g_blocks.back();
    }
    best->busy = true;
    held.push_back(best);
    *p = best->p;
    return DSP_OK;
}

BitRuns make_bit_runs(unsigned long long bits_used) {
    BitRuns runs{};
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
    return runs;
}

namespace {

inline unsigned blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

template <typename K>
__global__ void sort_keys_kernel(const Rec* __restrict__ rec, int64_t m, BitRuns runs, K* __restrict__ kc, uint32_t* __restrict__ pc) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    kc[c] = (K)runs.squeeze(rec[c].key);
    pc[c] = (uint32_t)c;
}

// Ordered float64 replay.  Block = RS sites (one per thread) whose records are the span
// [offsets[first site], offsets[last site] + counts[last site]) of the sorted order.  The span goes through
// shared memory RC entries at a time -- sorted positions read coalesced, one 32-byte sector per record --
// and every thread then consumes the entries of ITS site in order.  A site of any coverage works (its thread
// keeps consuming across chunks); the sums are the reference's left-to-right float64 additions.
constexpr int RS = 256, RC = 1024;
template <typename K>
__global__ void __launch_bounds__(RS) replay_kernel(const uint32_t* __restrict__ sorted_pos, const int64_t* __restrict__ offsets,
                                                   const int32_t* __restrict__ counts, int64_t nseg, const Rec* __restrict__ rec,
                                                   const K* __restrict__ ukey, BitRuns runs, SiteRow* __restrict__ rows) {
    __shared__ double s_p0[RC], s_p1[RC];
    __shared__ uint64_t s_gl[RC];
    __shared__ int64_t s_span[2];
    const int64_t s = (int64_t)blockIdx.x * RS + threadIdx.x;
    const bool valid = s < nseg;
    const int64_t my_off = valid ? offsets[s] : 0;
    const int64_t my_end = valid ? my_off + counts[s] : 0;
    if (threadIdx.x == 0) s_span[0] = my_off;
    if (valid && (s == nseg - 1 || threadIdx.x == RS - 1)) s_span[1] = my_end;
    __syncthreads();
    const int64_t lo = s_span[0], hi = s_span[1];
    double a0 = 0.0, a1 = 0.0;
    int32_t m1 = 0;
    uint64_t first = 0;
    for (int64_t base = lo; base < hi; base += RC) {
        const int len = (int)min((int64_t)RC, hi - base);
        for (int i = threadIdx.x; i < len; i += RS) {
            const Rec r = load_rec(rec + sorted_pos[base + i]);
            s_p0[i] = r.p0; s_p1[i] = r.p1; s_gl[i] = r.gl;
        }
        __syncthreads();
        const int64_t j0 = max(my_off, base), j1 = min(my_end, base + (int64_t)len);
        for (int64_t j = j0; j < j1; ++j) {
            const int i = (int)(j - base);
            const uint64_t gl = s_gl[i];
            if (j == my_off) first = gl & ~REC_LABEL_BIT;
            a0 = __dadd_rn(a0, s_p0[i]);          // call_mods_freq.py:60-61, float64, file order
            a1 = __dadd_rn(a1, s_p1[i]);
            m1 += (int32_t)(gl >> 63);
        }
        __syncthreads();
    }
    if (valid) {
        SiteRow r;
        r.key = runs.spread((uint64_t)ukey[s]); r.first = first; r.s0 = a0; r.s1 = a1;
        r.cov = (int32_t)(my_end - my_off); r.met = m1; r.unmet = r.cov - m1; r.pad = 0;
        rows[s] = r;
    }
}

template <typename F>
__global__ void row_field_kernel(const SiteRow* __restrict__ rows, int64_t n, int by_key, F* __restrict__ f, uint32_t* __restrict__ id) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { f[i] = (F)(by_key ? rows[i].key : rows[i].first); id[i] = (uint32_t)i; }
}
__global__ void gather_rows_kernel(const SiteRow* __restrict__ rows, const uint32_t* __restrict__ perm, int64_t n, SiteRow* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rows[perm[i]];
}

__global__ void emit_kernel(const SiteRow* __restrict__ rows, const uint32_t* __restrict__ perm, int64_t n,
                            uint64_t* __restrict__ out_key, int64_t* __restrict__ out_first,
                            double* __restrict__ out_p0, double* __restrict__ out_p1,
                            int32_t* __restrict__ out_met, int32_t* __restrict__ out_unmet, int32_t* __restrict__ out_cov) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SiteRow r = rows[perm ? perm[i] : (uint32_t)i];
    out_key[i] = r.key; out_first[i] = (int64_t)r.first; out_p0[i] = r.s0; out_p1[i] = r.s1;
    out_met[i] = r.met; out_unmet[i] = r.unmet; out_cov[i] = r.cov;
}

template <typename K>
int sites_typed(Scratch& sc, cudaStream_t st, const Rec* rec, int64_t m, const BitRuns& runs, SiteRow* rows, int64_t* nseg_host) {
    int rc;
    size_t tmp_bytes = 0, tmp_cap = 0;
    void* tmp = nullptr;
    auto ensure_tmp = [&](size_t need) -> int {
        if (need <= tmp_cap) return DSP_OK;
        tmp_cap = need;
        return sc.alloc((uint8_t**)&tmp, tmp_cap);
    };
    K *kc, *ks; uint32_t *pc, *ps;
    if ((rc = sc.alloc(&kc, m)) || (rc = sc.alloc(&ks, m)) || (rc = sc.alloc(&pc, m)) || (rc = sc.alloc(&ps, m))) return rc;
    sort_keys_kernel<K><<<blocks(m), 256, 0, st>>>(rec, m, runs, kc, pc);
    DSP_CUDA(cudaGetLastError());
    DSP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kc, ks, pc, ps, (int)m, 0, runs.bits, st));
    if ((rc = ensure_tmp(tmp_bytes))) return rc;
    DSP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kc, ks, pc, ps, (int)m, 0, runs.bits, st));

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
    replay_kernel<K><<<(unsigned)((nseg + RS - 1) / RS), RS, 0, st>>>(ps, offsets, counts, nseg, rec, ukey, runs, rows);
    DSP_CUDA(cudaGetLastError());
    *nseg_host = nseg;
    return DSP_OK;
}

}  // namespace

int sites_from_records(Scratch& sc, cudaStream_t st, const Rec* rec, int64_t m, unsigned long long bits_used,
                       SiteRow* rows, int64_t* nseg_host) {
    *nseg_host = 0;
    if (m == 0) return DSP_OK;
    {
        // the replay gathers one aligned 32-byte record per sorted position: ask L2 not to fetch the neighbouring
        // sectors of the 128-byte line (measured: 125 B of DRAM reads per record without this)
        static bool hinted[64] = {};
        if (sc.device >= 0 && sc.device < 64 && !hinted[sc.device]) {
            hinted[sc.device] = true;
            if (!getenv("DSP_B200_KEEP_L2_GRANULARITY")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
            cudaGetLastError();
        }
    }
    DSP_REQUIRE(m < (int64_t)0x7fffffff, DSP_ERR_INVALID, "call_freq: %lld records in one aggregation pass (< 2^31; shard by key)", (long long)m);
    const BitRuns runs = make_bit_runs(bits_used);
    return runs.bits <= 32 ? sites_typed<uint32_t>(sc, st, rec, m, runs, rows, nseg_host)
                           : sites_typed<uint64_t>(sc, st, rec, m, runs, rows, nseg_host);
}

int pack_callable(Scratch& sc, cudaStream_t st, const uint64_t* key, const double* p0, const double* p1,
                  const int32_t* label, int64_t n, uint64_t gidx_base, double prob_cf, Rec* rec,
                  int64_t* m_host, unsigned long long* bits_host) {
    using namespace route;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, sc.device);
    const Plan plan = make_plan(n, n_sm);
    int32_t* blk_counts; int64_t* blk_off; int64_t* totals; unsigned long long* d_bits; Targets* d_tg;
    int rc;
    if ((rc = sc.alloc(&blk_counts, (size_t)plan.blocks * MAXD)) || (rc = sc.alloc(&blk_off, (size_t)plan.blocks * MAXD)) ||
        (rc = sc.alloc(&totals, MAXD)) || (rc = sc.alloc(&d_bits, 1)) || (rc = sc.alloc(&d_tg, 1))) return rc;
    RecFromColumns src{key, p0, p1, label, gidx_base, prob_cf, 1};
    Targets tg{};
    tg.dst[0] = rec; tg.base[0] = 0;
    DSP_CUDA(cudaMemsetAsync(d_bits, 0, sizeof(unsigned long long), st));
    DSP_CUDA(cudaMemcpyAsync(d_tg, &tg, sizeof(tg), cudaMemcpyHostToDevice, st));
    count_kernel<RecFromColumns><<<plan.blocks, RT, 0, st>>>(src, plan, 1, blk_counts);
    scan_kernel<<<1, MAXD * 32, 0, st>>>(blk_counts, plan.blocks, 1, blk_off, totals);
    scatter_kernel<RecFromColumns><<<plan.blocks, RT, 0, st>>>(src, plan, 1, blk_off, d_tg, nullptr, d_bits);
    DSP_CUDA(cudaGetLastError());
    int64_t m = 0;
    DSP_CUDA(cudaMemcpyAsync(&m, totals, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DSP_CUDA(cudaMemcpyAsync(bits_host, d_bits, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    DSP_CUDA(cudaStreamSynchronize(st));           // also: `tg` must outlive the copy above
    *m_host = m;
    return DSP_OK;
}

template <typename F>
int order_rows_typed(Scratch& sc, cudaStream_t st, const SiteRow* rows, int64_t n, int by_key, uint32_t** perm_out, int end_bit) {
    F *f, *fs; uint32_t *id, *perm;
    int rc;
    if ((rc = sc.alloc(&f, n)) || (rc = sc.alloc(&fs, n)) || (rc = sc.alloc(&id, n)) || (rc = sc.alloc(&perm, n))) return rc;
    row_field_kernel<F><<<blocks(n), 256, 0, st>>>(rows, n, by_key, f, id);
    DSP_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    void* tmp;
    DSP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, f, fs, id, perm, (int)n, 0, end_bit, st));
    if ((rc = sc.alloc((uint8_t**)&tmp, tmp_bytes))) return rc;
    DSP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, f, fs, id, perm, (int)n, 0, end_bit, st));
    *perm_out = perm;
    return DSP_OK;
}

// rows (n) -> permutation that orders them by `first` (by_key == 0) or by key; end_bit = significant bits of the field
int order_rows(Scratch& sc, cudaStream_t st, const SiteRow* rows, int64_t n, int by_key, uint32_t** perm_out, int end_bit = 64) {
    return end_bit <= 32 ? order_rows_typed<uint32_t>(sc, st, rows, n, by_key, perm_out, end_bit)
                         : order_rows_typed<uint64_t>(sc, st, rows, n, by_key, perm_out, end_bit);
}

int sort_rows(Scratch& sc, cudaStream_t st, const SiteRow* rows, int64_t n, int by_key, SiteRow* out, int end_bit) {
    if (n == 0) return DSP_OK;
    uint32_t* perm;
    int rc = order_rows(sc, st, rows, n, by_key, &perm, end_bit);
    if (rc) return rc;
    gather_rows_kernel<<<blocks(n), 256, 0, st>>>(rows, perm, n, out);
    DSP_CUDA(cudaGetLastError());
    return DSP_OK;
}

}  // namespace dsp

using namespace dsp;

namespace dsp {
namespace {

// the whole aggregation on device columns (current device = sc.device); outputs on the device, capacity n
int aggregate_columns(Scratch& sc, cudaStream_t st, const uint64_t* key, const double* p0, const double* p1, const int32_t* label,
                      int64_t n, double prob_cf, int sort_by_key, uint64_t* out_key, int64_t* out_first, double* out_p0,
                      double* out_p1, int32_t* out_met, int32_t* out_unmet, int32_t* out_cov, int64_t* n_sites_host) {
    int rc;
    Rec* rec;
    if ((rc = sc.alloc(&rec, n))) return rc;
    int64_t m = 0;
    unsigned long long bits = 0;
    if ((rc = pack_callable(sc, st, key, p0, p1, label, n, 0, prob_cf, rec, &m, &bits))) return rc;
    if (m == 0) return DSP_OK;
    SiteRow* rows;
    if ((rc = sc.alloc(&rows, m))) return rc;
    int64_t nseg = 0;
    if ((rc = sites_from_records(sc, st, rec, m, bits, rows, &nseg))) return rc;
    uint32_t* perm = nullptr;                         // rows are in key order; dict insertion order = by first appearance
    int first_bits = 1;
    while (first_bits < 63 && (1ll << first_bits) < n) ++first_bits;              // `first` is a record index < n
    if (!sort_by_key && (rc = order_rows(sc, st, rows, nseg, 0, &perm, first_bits))) return rc;
    emit_kernel<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(rows, perm, nseg, out_key, out_first, out_p0, out_p1, out_met, out_unmet, out_cov);
    DSP_CUDA(cudaGetLastError());
    DSP_CUDA(cudaStreamSynchronize(st));
    *n_sites_host = nseg;
    return DSP_OK;
}

// key = (rank of the chromosome code << 40) | pos; flags positions outside [0, 2^40)
__global__ void make_keys_kernel(const int32_t* __restrict__ code, const int64_t* __restrict__ rank, const int64_t* __restrict__ pos,
                                 int64_t n, uint64_t* __restrict__ key, int* __restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = pos[i];
    if (p < 0 || p >= (1ll << 40)) *bad = 1;
    key[i] = ((uint64_t)rank[code[i]] << 40) | (uint64_t)p;
}

int check_device(int device, const char* who) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("%s: no CUDA device available; this library has no CPU path", who);
        return DSP_ERR_CUDA;
    }
    DSP_REQUIRE(device >= 0 && device < ndev, DSP_ERR_INVALID, "%s: bad device %d", who, device);
    return DSP_OK;
}

}  // namespace
}  // namespace dsp

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
    int rc = check_device(device, "dsp_freq_aggregate");
    if (rc) return rc;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    struct Restore { int p, d; ~Restore() { if (p != d && p >= 0) cudaSetDevice(p); } } restore{prev, device};
    Scratch sc(device);
    return aggregate_columns(sc, (cudaStream_t)stream, key, p0, p1, label, n, prob_cf, sort_by_key, out_key, out_first, out_p0,
                             out_p1, out_met, out_unmet, out_cov, n_sites_host);
}

extern "C" int dsp_freq_aggregate_host(int device, const int32_t* chrom_code, const int64_t* code_rank, int32_t n_codes,
                                       const int64_t* pos, const double* p0, const double* p1, const int32_t* label,
                                       int64_t n, double prob_cf, int sort_by_key,
                                       uint64_t* out_key, int64_t* out_first, double* out_p0, double* out_p1,
                                       int32_t* out_met, int32_t* out_unmet, int32_t* out_cov, int64_t out_cap,
                                       int64_t* n_sites_host) {
    DSP_REQUIRE(n_sites_host, DSP_ERR_INVALID, "dsp_freq_aggregate_host: n_sites_host is null");
    *n_sites_host = 0;
    DSP_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, DSP_ERR_INVALID,
                "dsp_freq_aggregate_host: n=%lld out of range (shard the records: < 2^31 per call)", (long long)n);
    if (n == 0) return DSP_OK;
    DSP_REQUIRE(chrom_code && code_rank && n_codes > 0 && pos && p0 && p1 && label && out_key && out_first && out_p0 && out_p1 &&
                out_met && out_unmet && out_cov, DSP_ERR_INVALID, "dsp_freq_aggregate_host: null pointer");
    int rc = check_device(device, "dsp_freq_aggregate_host");
    if (rc) return rc;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    struct Restore { int p, d; ~Restore() { if (p != d && p >= 0) cudaSetDevice(p); } } restore{prev, device};
    cudaStream_t st = nullptr;
    Scratch sc(device);
    int32_t* d_code; int64_t *d_rank, *d_pos; double *d_p0, *d_p1; int32_t* d_lab; uint64_t* d_key; int* d_bad;
    if ((rc = sc.alloc(&d_code, n)) || (rc = sc.alloc(&d_rank, n_codes)) || (rc = sc.alloc(&d_pos, n)) || (rc = sc.alloc(&d_p0, n)) ||
        (rc = sc.alloc(&d_p1, n)) || (rc = sc.alloc(&d_lab, n)) || (rc = sc.alloc(&d_key, n)) || (rc = sc.alloc(&d_bad, 1))) return rc;
    DSP_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    DSP_CUDA(cudaMemcpyAsync(d_code, chrom_code, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    DSP_CUDA(cudaMemcpyAsync(d_rank, code_rank, sizeof(int64_t) * n_codes, cudaMemcpyHostToDevice, st));
    DSP_CUDA(cudaMemcpyAsync(d_pos, pos, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    make_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_code, d_rank, d_pos, n, d_key, d_bad);
    DSP_CUDA(cudaGetLastError());
    DSP_CUDA(cudaMemcpyAsync(d_p0, p0, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    DSP_CUDA(cudaMemcpyAsync(d_p1, p1, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    DSP_CUDA(cudaMemcpyAsync(d_lab, label, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    int bad = 0;
    DSP_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    DSP_CUDA(cudaStreamSynchronize(st));
    DSP_REQUIRE(!bad, DSP_ERR_INVALID, "positions must be in [0, 2^40)");
    uint64_t* o_key; int64_t* o_first; double *o_p0, *o_p1; int32_t *o_met, *o_unmet, *o_cov;
    if ((rc = sc.alloc(&o_key, n)) || (rc = sc.alloc(&o_first, n)) || (rc = sc.alloc(&o_p0, n)) || (rc = sc.alloc(&o_p1, n)) ||
        (rc = sc.alloc(&o_met, n)) || (rc = sc.alloc(&o_unmet, n)) || (rc = sc.alloc(&o_cov, n))) return rc;
    int64_t ns = 0;
    if ((rc = aggregate_columns(sc, st, d_key, d_p0, d_p1, d_lab, n, prob_cf, sort_by_key, o_key, o_first, o_p0, o_p1, o_met, o_unmet,
                                o_cov, &ns))) return rc;
    *n_sites_host = ns;
    DSP_REQUIRE(ns <= out_cap, DSP_ERR_NOMEM, "dsp_freq_aggregate_host: %lld sites, output capacity %lld", (long long)ns, (long long)out_cap);
    DSP_CUDA(cudaMemcpy(out_key, o_key, sizeof(uint64_t) * ns, cudaMemcpyDeviceToHost));
    DSP_CUDA(cudaMemcpy(out_first, o_first, sizeof(int64_t) * ns, cudaMemcpyDeviceToHost));
    DSP_CUDA(cudaMemcpy(out_p0, o_p0, sizeof(double) * ns, cudaMemcpyDeviceToHost));
    DSP_CUDA(cudaMemcpy(out_p1, o_p1, sizeof(double) * ns, cudaMemcpyDeviceToHost));
    DSP_CUDA(cudaMemcpy(out_met, o_met, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost));
    DSP_CUDA(cudaMemcpy(out_unmet, o_unmet, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost));
    DSP_CUDA(cudaMemcpy(out_cov, o_cov, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost));
    return DSP_OK;
}

extern "C" int dsp_device_warmup(int device) {
    int rc = check_device(device, "dsp_device_warmup");
    if (rc) return rc;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    DSP_CUDA(cudaFree(nullptr));                      // creates the primary context
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return DSP_OK;
}

extern "C" int dsp_freq_release_cache(void) {
    std::lock_guard<std::mutex> lk(g_blocks_mu);
    int prev = -1;
    cudaGetDevice(&prev);
    for (auto it = g_blocks.begin(); it != g_blocks.end();) {
        if (it->busy) { ++it; continue; }           // held by a running call on another thread: stays
        cudaSetDevice(it->device);
        cudaFree(it->p);
        it = g_blocks.erase(it);                    // list: pointers held by live Scratch objects stay valid
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
    return DSP_OK;
}
