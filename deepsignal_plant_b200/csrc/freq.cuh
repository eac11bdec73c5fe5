// Internal interface of the call_freq device code: record / site-row formats shared by the
// aggregation (freq.cu) and the multi-GPU routing (comm.cu).  Not part of the C ABI.
#pragma once
#include "common.cuh"

namespace dsp {

// A callable per-read call, packed: one aligned 32-byte sector holds everything the ordered replay
// needs.  `gl` = record index (file order: the original index on one GPU, the GLOBAL record index
// across ranks) in bits [0,63) and (called_label == 1) in bit 63.
struct __align__(32) Rec { uint64_t key; double p0, p1; uint64_t gl; };
constexpr uint64_t REC_LABEL_BIT = 1ull << 63;

// One 256-bit access per record (LDG/STG.E.256 on sm_100): a record is one aligned 32-byte sector, and one store
// instruction per record is also one NVLink write instead of two half-sector ones when the destination is a peer.
__device__ __forceinline__ void store_rec(Rec* dst, const Rec& r) {
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(r.key), "l"((uint64_t)__double_as_longlong(r.p0)),
                 "l"((uint64_t)__double_as_longlong(r.p1)), "l"(r.gl) : "memory");
}
__device__ __forceinline__ Rec load_rec(const Rec* src) {
    uint64_t a, b, c, d;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(src));
    Rec r; r.key = a; r.p0 = __longlong_as_double((long long)b); r.p1 = __longlong_as_double((long long)c); r.gl = d;
    return r;
}

// One site of the frequency table (call_mods_freq.py:55-66): 48 bytes.
struct __align__(16) SiteRow { uint64_t key; uint64_t first; double s0, s1; int32_t met, unmet, cov, pad; };

// Scratch memory: a per-process cache of device blocks (grow-only, best fit); a call makes ~20
// allocations and cudaMalloc/cudaFree around every call cost several times the kernels.
struct Scratch {
    std::vector<void*> held;
    int device;
    explicit Scratch(int dev) : device(dev) {}
    ~Scratch();
    int alloc_bytes(void** p, size_t bytes);
    template <typename T> int alloc(T** p, size_t count) { return alloc_bytes((void**)p, (count ? count : 1) * sizeof(T)); }
};

// Which key bits are ever set, as runs of consecutive ones: the sort key is the site key with every
// never-set bit squeezed out (order preserving), which turns the 43-bit sort of chrom<<40|pos keys
// into a ~22-bit one.
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
BitRuns make_bit_runs(unsigned long long bits_used);

// m packed records (file order within equal keys is their order in `rec`) -> one SiteRow per
// distinct key, in ascending key order.  rows has capacity m.  Synchronises `st` once.
int sites_from_records(Scratch& sc, cudaStream_t st, const Rec* rec, int64_t m, unsigned long long bits_used,
                       SiteRow* rows, int64_t* nseg_host);

// Filter (|p0-p1| >= prob_cf) + pack into `rec` (capacity n), file order kept; *m_host, *bits_host =
// callable count and the OR of their keys.  Synchronises `st` once.
int pack_callable(Scratch& sc, cudaStream_t st, const uint64_t* key, const double* p0, const double* p1,
                  const int32_t* label, int64_t n, uint64_t gidx_base, double prob_cf, Rec* rec,
                  int64_t* m_host, unsigned long long* bits_host);

}  // namespace dsp
