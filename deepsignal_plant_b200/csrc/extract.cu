// Feature extraction from decoded re-squiggled reads:
//   deepsignal_plant/extract_features.py  _rescale_signals (:276-277), _normalize_signals (:179-190),
//   the per-site body of _extract_features (:343-372) and _get_signals_rect (:232-251).
// Integer / float64 work, bit-exact against numpy's arithmetic:
//   read_scale_kernel    one CTA per read: two passes over the int16 samples (min/max, then a shared-memory
//                        histogram of the DAC levels); median and MAD are radix-selected over the histogram
//                        bins, exact because equal DAC values give equal float64 values (reads spanning more
//                        than 16 384 levels run the same select over the samples);
//   read_zscore_kernel   normalize_method 'zscore': whole-read np.mean / np.std in numpy's pairwise order, one
//                        warp per read;
//   site_features_kernel one warp per site: the site's window of samples is normalised once into shared
//                        memory, lane j owns base j (len, np.mean, np.std in numpy's pairwise order, ordered
//                        subsample -- caller-supplied offsets in parity mode, Philox selection sampling
//                        otherwise), all lanes emit the 13 x 16 rectangle; float32 outputs are the five tensors
//                        dsp_forward takes, float64 outputs what the feature file prints;
//   dsp_find_sites       motif search + the reference's site filters as one DeviceSelect over the event table.
// Every double operation is an explicit round-to-nearest intrinsic so that nvcc cannot contract a
// multiply and an add into an FMA (numpy does not).
#include "common.cuh"
#include <cub/block/block_scan.cuh>
#include <cub/block/block_reduce.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <math_constants.h>

namespace dsp {
namespace {

constexpr int SEL_THREADS = 512;
constexpr int SEL_BITS = 11;
constexpr int SEL_BINS = 1 << SEL_BITS;
constexpr int BINS_PER_THREAD = SEL_BINS / SEL_THREADS;
constexpr int HIST_BINS = 16384;                // DAC levels a read may span on the histogram path (64 KB)
constexpr double MAD_C = 0.6744897501960817;   // scipy.stats.norm.ppf(3/4.), statsmodels.robust.mad's default c

__device__ __forceinline__ uint64_t key_of(double v) {
    const uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(uint64_t k) {
    const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// _rescale_signals: scaling * (raw + offset) in float64; a NaN scaling marks a read without channel
// info (extract_features.py:313-315 leaves the DAC values as they are).
struct Rescale {
    double scaling, offset;
    bool on;
    __device__ __forceinline__ double operator()(int16_t r) const {
        const double x = (double)r;
        return on ? __dmul_rn(scaling, __dadd_rn(x, offset)) : x;
    }
};

struct SelectShared {
    int hist[SEL_BINS];
    uint64_t prefix;
    int64_t k;
    typename cub::BlockScan<int, SEL_THREADS>::TempStorage scan;
    union {
        typename cub::BlockReduce<unsigned long long, SEL_THREADS>::TempStorage red_u;
        typename cub::BlockReduce<int, SEL_THREADS>::TempStorage red_i;
    } red;
    uint64_t max_less;
    int count_less;
};

// Key of the element of rank k (0-based) in the multiset {keyfn(i) repeated wfn(i) times, i < n_items}:
// most-significant-digit radix select, 11-bit digits, shared-memory histogram per pass.  The items are
// either the samples themselves (weight 1) or the bins of a per-read histogram of DAC values.
template <typename KeyFn, typename WFn>
__device__ uint64_t block_select(SelectShared& sh, KeyFn keyfn, WFn wfn, int64_t n_items, int64_t k) {
    const int tid = threadIdx.x;
    if (tid == 0) { sh.prefix = 0; sh.k = k; }
    int hi = 64;                                     // bits [hi, 64) of the answer are known (= sh.prefix)
    while (hi > 0) {
        const int width = hi >= SEL_BITS ? SEL_BITS : hi;
        const int lo = hi - width;
        for (int b = tid; b < SEL_BINS; b += SEL_THREADS) sh.hist[b] = 0;
        __syncthreads();
        const uint64_t prefix = sh.prefix;
        for (int64_t i0 = 0; i0 < n_items; i0 += SEL_THREADS) {
            const int64_t i = i0 + tid;
            int w = i < n_items ? wfn(i) : 0;
            uint64_t key = 0;
            if (w > 0) {
                key = keyfn(i);
                if (hi != 64 && (key >> hi) != prefix) w = 0;
            }
            const int bin = (int)((key >> lo) & ((1u << width) - 1));
            // the high digits of neighbouring items mostly agree: one atomic per warp when they all do
            const unsigned act = __ballot_sync(0xffffffffu, w > 0);
            if (act == 0) continue;
            const int leader = __ffs(act) - 1;
            const int lbin = __shfl_sync(0xffffffffu, bin, leader);
            const unsigned same = __ballot_sync(0xffffffffu, w > 0 && bin == lbin);
            if (same == act) {
                const int tot = __reduce_add_sync(0xffffffffu, w);
                if ((tid & 31) == leader) atomicAdd(&sh.hist[lbin], tot);
            } else if (w > 0) {
                atomicAdd(&sh.hist[bin], w);
            }
        }
        __syncthreads();
        int c[BINS_PER_THREAD], tsum = 0;
#pragma unroll
        for (int j = 0; j < BINS_PER_THREAD; ++j) { c[j] = sh.hist[tid * BINS_PER_THREAD + j]; tsum += c[j]; }
        int before;
        cub::BlockScan<int, SEL_THREADS>(sh.scan).ExclusiveSum(tsum, before);
        const int64_t kk = sh.k;
        __syncthreads();
        if (kk >= before && kk < (int64_t)before + tsum) {   // exactly one thread
            int64_t rem = kk - before;
            int j = 0;
            while (rem >= c[j]) { rem -= c[j]; ++j; }
            sh.prefix = (prefix << width) | (uint64_t)(tid * BINS_PER_THREAD + j);
            sh.k = rem;
        }
        __syncthreads();
        hi = lo;
    }
    const uint64_t res = sh.prefix;
    __syncthreads();                                 // the next select re-initialises sh.prefix
    return res;
}

// numpy's median of an even count is the mean of the two middle order statistics: given the key of
// rank k, the value of rank k-1 is either the same (ties) or the largest key below it.
template <typename KeyFn, typename WFn>
__device__ uint64_t block_rank_below(SelectShared& sh, KeyFn keyfn, WFn wfn, int64_t n_items, int64_t k, uint64_t key_k) {
    unsigned long long mx = 0;
    int cnt = 0;
    for (int64_t i = threadIdx.x; i < n_items; i += SEL_THREADS) {
        const int w = wfn(i);
        if (w <= 0) continue;
        const uint64_t key = keyfn(i);
        if (key < key_k) { cnt += w; mx = key > mx ? key : mx; }
    }
    const unsigned long long bmx = cub::BlockReduce<unsigned long long, SEL_THREADS>(sh.red.red_u).Reduce(mx, cub::Max());
    __syncthreads();
    const int bcnt = cub::BlockReduce<int, SEL_THREADS>(sh.red.red_i).Sum(cnt);
    if (threadIdx.x == 0) { sh.max_less = bmx; sh.count_less = bcnt; }
    __syncthreads();
    const uint64_t r = ((int64_t)sh.count_less <= k - 1) ? key_k : sh.max_less;
    __syncthreads();
    return r;
}

// median (np.median: mean of the two middle values for an even count) of valfn over the weighted items,
// then optionally divided by `div` element-wise before the mean (statsmodels' mad divides by c first)
template <typename ValFn, typename WFn>
__device__ double block_median(SelectShared& sh, ValFn valfn, WFn wfn, int64_t n_items, int64_t n, double div) {
    auto keyfn = [&](int64_t i) { return key_of(valfn(i)); };
    const int64_t k = n / 2;
    const uint64_t key_hi = block_select(sh, keyfn, wfn, n_items, k);
    const double v_hi = div == 1.0 ? value_of(key_hi) : __ddiv_rn(value_of(key_hi), div);
    if (n & 1) return v_hi;
    const uint64_t key_lo = block_rank_below(sh, keyfn, wfn, n_items, k, key_hi);
    const double v_lo = div == 1.0 ? value_of(key_lo) : __ddiv_rn(value_of(key_lo), div);
    return __ddiv_rn(__dadd_rn(v_lo, v_hi), 2.0);
}

// every sample of a read, 8 per 16-byte load once the pointer is aligned
template <typename F>
__device__ __forceinline__ void for_each_sample(const int16_t* __restrict__ x, int64_t n, F fn) {
    const int tid = threadIdx.x;
    int64_t head = (int64_t)(((16 - ((uintptr_t)x & 15)) & 15) >> 1);
    if (head > n) head = n;
    for (int64_t i = tid; i < head; i += SEL_THREADS) fn((int)x[i]);
    const uint4* v = reinterpret_cast<const uint4*>(x + head);
    const int64_t nv = (n - head) >> 3;
    for (int64_t i = tid; i < nv; i += SEL_THREADS) {
        const uint4 q = v[i];
        fn((int)(int16_t)(q.x & 0xffff)); fn((int)(int16_t)(q.x >> 16));
        fn((int)(int16_t)(q.y & 0xffff)); fn((int)(int16_t)(q.y >> 16));
        fn((int)(int16_t)(q.z & 0xffff)); fn((int)(int16_t)(q.z >> 16));
        fn((int)(int16_t)(q.w & 0xffff)); fn((int)(int16_t)(q.w >> 16));
    }
    for (int64_t i = head + (nv << 3) + tid; i < n; i += SEL_THREADS) fn((int)x[i]);
}

// One CTA per read.  Fast path (DAC values of the read span <= HIST_BINS levels, i.e. every real read):
// two passes over the samples -- min/max, then a shared-memory histogram of the DAC values -- and both
// medians are selected over the histogram bins (value of a bin = the rescaled level, weight = its count),
// which is exact because equal DAC values give equal float64 values.  Otherwise the same selects run over
// the samples themselves.
__global__ void __launch_bounds__(SEL_THREADS)
read_scale_kernel(const int16_t* __restrict__ raw, const int64_t* __restrict__ raw_off,
                  const double* __restrict__ scaling, const double* __restrict__ offset, int64_t n_reads,
                  double* __restrict__ shift_out, double* __restrict__ scale_out) {
    __shared__ SelectShared sh;
    __shared__ int s_mn, s_mx;
    extern __shared__ int level_count[];               // HIST_BINS
    for (int64_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const int16_t* x = raw + raw_off[r];
        const int64_t n = raw_off[r + 1] - raw_off[r];
        if (n <= 0) {                                   // np.median of nothing is nan
            if (threadIdx.x == 0) { shift_out[r] = CUDART_NAN; scale_out[r] = CUDART_NAN; }
            continue;
        }
        Rescale f;
        f.scaling = scaling[r]; f.offset = offset[r]; f.on = !(f.scaling != f.scaling);
        int mn = 32767, mx = -32768;
        for_each_sample(x, n, [&](int v) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; });
        mn = cub::BlockReduce<int, SEL_THREADS>(sh.red.red_i).Reduce(mn, cub::Min());
        __syncthreads();
        mx = cub::BlockReduce<int, SEL_THREADS>(sh.red.red_i).Reduce(mx, cub::Max());
        if (threadIdx.x == 0) { s_mn = mn; s_mx = mx; }
        __syncthreads();
        mn = s_mn; mx = s_mx;
        const int levels = mx - mn + 1;
        double med, mad;
        if (levels <= HIST_BINS) {
            for (int b = threadIdx.x; b < levels; b += SEL_THREADS) level_count[b] = 0;
            __syncthreads();
            for_each_sample(x, n, [&](int v) { atomicAdd(&level_count[v - mn], 1); });
            __syncthreads();
            auto w = [&](int64_t i) { return level_count[i]; };
            med = block_median(sh, [&](int64_t i) { return f((int16_t)(mn + (int)i)); }, w, levels, n, 1.0);
            // statsmodels.robust.mad: median(|a - median(a)| / c).  Dividing by the positive constant is
            // monotone, so the order statistics are selected on |a - median| and divided afterwards.
            mad = block_median(sh, [&](int64_t i) { return fabs(__dsub_rn(f((int16_t)(mn + (int)i)), med)); }, w, levels, n, MAD_C);
        } else {
            auto w = [](int64_t) { return 1; };
            med = block_median(sh, [&](int64_t i) { return f(x[i]); }, w, n, n, 1.0);
            mad = block_median(sh, [&](int64_t i) { return fabs(__dsub_rn(f(x[i]), med)); }, w, n, n, MAD_C);
        }
        if (threadIdx.x == 0) { shift_out[r] = med; scale_out[r] = mad; }
        __syncthreads();
    }
}

// np.around(v, 6): rint(v * 1e6) / 1e6 in float64
__device__ __forceinline__ double around6(double v) { return __ddiv_rn(rint(__dmul_rn(v, 1e6)), 1e6); }

// One base's samples as _normalize_signals leaves them: around((x - shift) / scale, 6), or around(x, 6)
// when the scale is 0 (:186-189).
struct NormSamples {
    const int16_t* x;
    Rescale f;
    double shift, scale;
    __device__ __forceinline__ double operator()(int64_t i) const {
        const double v = f(x[i]);
        return around6(scale == 0.0 ? v : __ddiv_rn(__dsub_rn(v, shift), scale));
    }
};

// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum) of elem(0..n-1):
// n < 8 a running sum from 0; n <= 128 eight interleaved accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail; larger n split at n/2 rounded down to a multiple of 8.
template <typename I, typename Elem>
__device__ double pairwise_leaf(Elem elem, I off, I n) {
    if (n < 8) {
        double res = 0.0;
        for (I i = 0; i < n; ++i) res = __dadd_rn(res, elem(off + i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = elem(off + (I)j);
    I i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], elem(off + i + (I)j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, elem(off + i));
    return res;
}

template <typename I, typename Elem>
__device__ double pairwise_sum(Elem elem, I n) {
    if (n <= 128) return pairwise_leaf<I>(elem, (I)0, n);
    // explicit stack instead of recursion: frames are (offset, count, phase, left value)
    constexpr int DEPTH = sizeof(I) == 4 ? 26 : 40;
    I off_s[DEPTH], n_s[DEPTH];
    double left_s[DEPTH];
    int phase_s[DEPTH];
    int sp = 0;
    off_s[0] = 0; n_s[0] = n; phase_s[0] = 0; left_s[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        const I o = off_s[sp], c = n_s[sp];
        if (phase_s[sp] == 0) {
            if (c <= 128) { ret = pairwise_leaf<I>(elem, o, c); --sp; continue; }
            I n2 = c / 2; n2 -= n2 % 8;
            phase_s[sp] = 1;
            ++sp; off_s[sp] = o; n_s[sp] = n2; phase_s[sp] = 0;
        } else if (phase_s[sp] == 1) {
            I n2 = c / 2; n2 -= n2 % 8;
            left_s[sp] = ret; phase_s[sp] = 2;
            ++sp; off_s[sp] = o + n2; n_s[sp] = c - n2; phase_s[sp] = 0;
        } else {
            ret = __dadd_rn(left_s[sp], ret); --sp;
        }
    }
    return ret;
}

// The same summation evaluated by one warp for whole-read sums (normalize_method 'zscore'): the stack walk
// is uniform across the warp, lanes 0..7 own the eight accumulators of a leaf.
template <typename Elem>
__device__ double warp_pairwise_leaf(Elem elem, int64_t off, int64_t n) {
    const int lane = threadIdx.x & 31;
    double res = 0.0;
    if (n < 8) {
        if (lane == 0) for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, elem(off + i));
    } else {
        const int64_t body = n - (n % 8);
        double r = 0.0;
        if (lane < 8) {
            r = elem(off + lane);
            for (int64_t i = 8; i < body; i += 8) r = __dadd_rn(r, elem(off + i + lane));
        }
        r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1));      // lanes 0,2,4,6: r0+r1, r2+r3, r4+r5, r6+r7
        r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 2));      // lanes 0,4
        r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 4));      // lane 0
        res = r;
        if (lane == 0) for (int64_t i = body; i < n; ++i) res = __dadd_rn(res, elem(off + i));
    }
    return __shfl_sync(0xffffffffu, res, 0);
}

template <typename Elem>
__device__ double warp_pairwise_sum(Elem elem, int64_t n) {
    if (n <= 128) return warp_pairwise_leaf(elem, 0, n);
    constexpr int DEPTH = 64;
    int64_t off_s[DEPTH], n_s[DEPTH];
    double left_s[DEPTH];
    int phase_s[DEPTH];
    int sp = 0;
    off_s[0] = 0; n_s[0] = n; phase_s[0] = 0; left_s[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        const int64_t o = off_s[sp], c = n_s[sp];
        if (phase_s[sp] == 0) {
            if (c <= 128) { ret = warp_pairwise_leaf(elem, o, c); --sp; continue; }
            int64_t n2 = c / 2; n2 -= n2 % 8;
            phase_s[sp] = 1;
            ++sp; off_s[sp] = o; n_s[sp] = n2; phase_s[sp] = 0;
        } else if (phase_s[sp] == 1) {
            int64_t n2 = c / 2; n2 -= n2 % 8;
            left_s[sp] = ret; phase_s[sp] = 2;
            ++sp; off_s[sp] = o + n2; n_s[sp] = c - n2; phase_s[sp] = 0;
        } else {
            ret = __dadd_rn(left_s[sp], ret); --sp;
        }
    }
    return ret;
}

// _normalize_signals with 'zscore' (:180-181): shift = np.mean, scale = np.std of the rescaled read, both in
// numpy's summation order.  One warp per read.
__global__ void __launch_bounds__(128)
read_zscore_kernel(const int16_t* __restrict__ raw, const int64_t* __restrict__ raw_off,
                   const double* __restrict__ scaling, const double* __restrict__ offset, int64_t n_reads,
                   double* __restrict__ shift_out, double* __restrict__ scale_out) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_reads) return;
    const int16_t* x = raw + raw_off[r];
    const int64_t n = raw_off[r + 1] - raw_off[r];
    Rescale f;
    f.scaling = scaling[r]; f.offset = offset[r]; f.on = !(f.scaling != f.scaling);
    const double dn = (double)n;
    const double mean = __ddiv_rn(warp_pairwise_sum([&](int64_t i) { return f(x[i]); }, n), dn);
    const double var = __ddiv_rn(warp_pairwise_sum([&](int64_t i) { const double d = __dsub_rn(f(x[i]), mean); return __dmul_rn(d, d); }, n), dn);
    if ((threadIdx.x & 31) == 0) { shift_out[r] = mean; scale_out[r] = __dsqrt_rn(var); }
}

// Philox4x32-10 (Salmon et al., SC'11): one 128-bit block per counter.
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__constant__ uint8_t c_base2code[256];

struct SiteParams {
    const int16_t* raw; const int64_t* raw_off; const double* scaling; const double* offset;
    const int64_t* ev_start; const int64_t* ev_len; const uint8_t* ev_base;
    const int32_t* site_read; const int64_t* site_ev; int64_t n_sites;
    int T, S, round_stats;
    const int32_t* drawn; uint64_t seed;
    const double* shift; const double* scale;
    void *kmer, *means, *stds, *lens, *signals;       // float (what the classifier reads) or double (what the feature file prints)
};

constexpr int SITE_WARPS = 8;                      // sites per CTA (one warp each)
constexpr int WIN_MAX = 448;                       // samples of a site's window held in shared memory (3.5 KB per warp; 4 CTAs per SM)
constexpr int RND_MAX = WIN_MAX + 4 * 32;          // uniforms for the subsample draws of a windowed site (one per sample)

// ordered uniform S-subset of 0..n-1 by selection sampling (Knuth 3.4.2 S): offset i is taken with
// probability (still needed) / (still available) -- the distribution of sorted(random.sample(range(n), S))
// uniform i of row `row` = word (i & 3) of Philox block (i >> 2) of that row
__device__ __forceinline__ uint4 subset_block(int64_t row, uint32_t block, uint64_t seed) {
    return philox4x32((uint32_t)row, (uint32_t)(row >> 32), block, 0x65787472u, (uint32_t)seed, (uint32_t)(seed >> 32));
}

template <typename Put>
__device__ __forceinline__ void draw_ordered_subset(int64_t n, int S, int64_t row, uint64_t seed, Put put) {
    int need = S, s = 0;
    uint4 blk = make_uint4(0, 0, 0, 0);
    for (int64_t i = 0; i < n && need > 0; ++i) {
        if ((i & 3) == 0) blk = subset_block(row, (uint32_t)(i >> 2), seed);
        const uint32_t u = (i & 3) == 0 ? blk.x : (i & 3) == 1 ? blk.y : (i & 3) == 2 ? blk.z : blk.w;
        // u / 2^32 < need / (n - i)   <=>   u * (n - i) < need * 2^32
        if ((uint64_t)u * (uint64_t)(n - i) < ((uint64_t)need << 32)) { put(s++, i); --need; }
    }
}

// the same draw with the uniforms already in shared memory (drawn by the whole warp)
template <typename Put>
__device__ __forceinline__ void draw_ordered_subset_from(const uint32_t* u, int n, int S, Put put) {
    int need = S, s = 0;
    for (int i = 0; i < n && need > 0; ++i)
        if ((uint64_t)u[i] * (uint64_t)(n - i) < ((uint64_t)need << 32)) { put(s++, i); --need; }
}

__host__ __device__ inline size_t site_smem_per_warp(int T, int S) {
    return (size_t)WIN_MAX * 8 + (size_t)T * 16 + (size_t)RND_MAX * 4 + 32 * 4 + (((size_t)T * S * 2 + 15) & ~(size_t)15);
}

// One warp per site.  The seq_len events of a site are neighbours in the raw signal, so the warp first
// normalises the whole window (lane-strided, coalesced, two float64 divisions per sample, each sample
// once) into shared memory; lane j then owns base j: len, np.mean, np.std (numpy's pairwise order, read
// from shared memory) and, for a base longer than the rectangle, its subsample offsets; finally all lanes
// emit the T x S rectangle with coalesced stores.  A site whose window does not fit (a stalled base)
// takes the same steps with the samples recomputed from the raw signal instead of read from the window.
template <typename OutT>
__global__ void __launch_bounds__(SITE_WARPS * 32, 4) site_features_kernel(SiteParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = p.T, S = p.S;
    const size_t per_warp = site_smem_per_warp(T, S);
    unsigned char* base = smem_raw + warp * per_warp;
    double* win = reinterpret_cast<double*>(base);
    int64_t* b_off = reinterpret_cast<int64_t*>(base + (size_t)WIN_MAX * 8);       // [T] first sample of base j (read-relative)
    int64_t* b_len = b_off + T;                                                     // [T]
    uint32_t* rnd = reinterpret_cast<uint32_t*>(b_len + T);                         // [RND_MAX] uniforms of the long bases
    int* b_blk = reinterpret_cast<int*>(rnd + RND_MAX);                             // [32] first Philox block of base j in rnd
    uint16_t* sel = reinterpret_cast<uint16_t*>(b_blk + 32);                        // [T][S] subsample offsets
    const int shift_S = (S & (S - 1)) == 0 ? __ffs(S) - 1 : -1;
    OutT* const o_kmer = static_cast<OutT*>(p.kmer);
    OutT* const o_means = static_cast<OutT*>(p.means);
    OutT* const o_stds = static_cast<OutT*>(p.stds);
    OutT* const o_lens = static_cast<OutT*>(p.lens);
    OutT* const o_signals = static_cast<OutT*>(p.signals);

    for (int64_t site = (int64_t)blockIdx.x * SITE_WARPS + warp; site < p.n_sites; site += (int64_t)gridDim.x * SITE_WARPS) {
        const int32_t rd = p.site_read[site];
        const int64_t ev0 = p.site_ev[site] - (T - 1) / 2;
        NormSamples v;
        v.x = p.raw + p.raw_off[rd];
        v.f.scaling = p.scaling[rd]; v.f.offset = p.offset[rd]; v.f.on = !(v.f.scaling != v.f.scaling);
        v.shift = p.shift[rd]; v.scale = p.scale[rd];
        int64_t st = 0x7fffffffffffffffll, en = 0, n = 0, off = 0;
        if (lane < T) {
            off = p.ev_start[ev0 + lane]; n = p.ev_len[ev0 + lane];
            st = off; en = off + n;
            b_off[lane] = off; b_len[lane] = n;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int64_t a = __shfl_xor_sync(0xffffffffu, st, d), b = __shfl_xor_sync(0xffffffffu, en, d);
            st = a < st ? a : st; en = b > en ? b : en;
        }
        const int64_t wlen = en - st;
        const bool windowed = wlen <= WIN_MAX;
        if (windowed) for (int64_t i = lane; i < wlen; i += 32) win[i] = v(st + i);
        __syncwarp();
        const int64_t row = site * T + lane;
        if (lane < T) {
            // np.mean / np.std of the slice (extract_features.py:363-364): sum / n, sqrt(sum((x - mean)^2) / n)
            const double dn = (double)n;
            const double* a = win + (off - st);
            double mean, sd;
            if (windowed) {
                mean = __ddiv_rn(pairwise_sum<int>([&](int i) { return a[i]; }, (int)n), dn);
                const double m0 = mean;
                sd = __dsqrt_rn(__ddiv_rn(pairwise_sum<int>([&](int i) { const double d = __dsub_rn(a[i], m0); return __dmul_rn(d, d); }, (int)n), dn));
            } else {
                mean = __ddiv_rn(pairwise_sum<int64_t>([&](int64_t i) { return v(off + i); }, n), dn);
                const double m0 = mean;
                sd = __dsqrt_rn(__ddiv_rn(pairwise_sum<int64_t>([&](int64_t i) { const double d = __dsub_rn(v(off + i), m0); return __dmul_rn(d, d); }, n), dn));
            }
            if (p.round_stats) { mean = around6(mean); sd = around6(sd); }   // _features_to_str, :388-389
            o_kmer[row] = (OutT)c_base2code[p.ev_base[ev0 + lane]];
            o_means[row] = (OutT)mean;
            o_stds[row] = (OutT)sd;
            o_lens[row] = (OutT)n;
        }
        if (!p.drawn) {
            // ordered subsamples of the bases longer than the rectangle.  Windowed site: the Philox blocks of
            // all its long bases are drawn by the whole warp (base j's blocks at b_blk[j] in rnd), then each
            // owner lane only compares; otherwise each owner lane draws its own.
            const bool lng = lane < T && n > S;
            const int nblk = (lng && windowed) ? (int)((n + 3) >> 2) : 0;
            int incl = nblk;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const bool coop = windowed && total > 0 && total * 4 <= RND_MAX;
            if (coop) {
                b_blk[lane] = incl - nblk;
                __syncwarp();
                for (int b = lane; b < total; b += 32) {
                    int j = 0;
                    for (int t = 1; t < T; ++t) if (b_blk[t] <= b) j = t;
                    const uint4 blk = subset_block(site * T + j, (uint32_t)(b - b_blk[j]), p.seed);
                    *reinterpret_cast<uint4*>(rnd + 4 * b) = blk;
                }
                __syncwarp();
                if (lng) draw_ordered_subset_from(rnd + 4 * (incl - nblk), (int)n, S, [&](int s, int i) { sel[lane * S + s] = (uint16_t)i; });
            } else if (lng) {
                if (n <= 65536) draw_ordered_subset(n, S, row, p.seed, [&](int s, int64_t i) { sel[lane * S + s] = (uint16_t)i; });
                else draw_ordered_subset(n, S, row, p.seed, [&](int s, int64_t i) { o_signals[row * S + s] = (OutT)v(off + i); });
            }
        }
        __syncwarp();
        // _get_signals_rect (:232-251): centred zero pad, or the ordered subsample
        OutT* out = o_signals + site * T * S;
        for (int e = lane; e < T * S; e += 32) {
            const int j = shift_S >= 0 ? e >> shift_S : e / S, sidx = e - j * S;
            const int64_t nj = b_len[j];
            int64_t i;
            if (nj <= S) {
                i = sidx - (S - nj) / 2;
                if (i < 0 || i >= nj) { out[e] = (OutT)0; continue; }
            } else if (p.drawn) {
                i = p.drawn[site * T * S + e];         // parity mode: replay the reference's random.sample offsets
            } else if (nj <= 65536) {
                i = sel[e];
            } else {
                continue;                              // written by the base's own lane above
            }
            out[e] = (OutT)(windowed ? win[b_off[j] - st + i] : v(b_off[j] + i));
        }
        __syncwarp();
    }
}

// ---- site search: get_refloc_of_methysite_in_motif (utils/process_utils.py:97-112) + the site filters of
// _extract_features (:341-352), evaluated per event of the concatenated event table
constexpr int MAX_MOTIFS = 64, MAX_MOTIF_LEN = 8;

struct SitePredicate {
    const uint8_t* ev_base; const int64_t* ev_off; int64_t n_reads;
    const int64_t* chrom_start; const uint8_t* minus; const int64_t* region_start; const int64_t* region_end;
    int n_motifs, motif_len, methyloc, num_bases;
    unsigned long long motif[MAX_MOTIFS];              // up to 8 letters, first letter in the low byte

    __device__ __forceinline__ int64_t read_of(int64_t e) const {     // largest r with ev_off[r] <= e
        int64_t lo = 0, hi = n_reads;
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (ev_off[mid] <= e) lo = mid; else hi = mid; }
        return lo;
    }
    __device__ __forceinline__ int64_t pos_of(int64_t r, int64_t loc, int64_t rlen) const {
        return minus[r] ? chrom_start[r] + rlen - 1 - loc : chrom_start[r] + loc;
    }
    __device__ bool operator()(int64_t e) const {
        const int64_t r = read_of(e);
        const int64_t r0 = ev_off[r], rlen = ev_off[r + 1] - r0, loc = e - r0;
        const int64_t i = loc - methyloc;                              // where the motif would start
        if (i < 0 || i + motif_len > rlen) return false;
        if (loc < num_bases || loc >= rlen - num_bases) return false;  // :341
        unsigned long long w = 0;
        for (int k = 0; k < motif_len; ++k) w |= (unsigned long long)ev_base[r0 + i + k] << (8 * k);
        bool hit = false;
        for (int m = 0; m < n_motifs; ++m) hit |= (w == motif[m]);
        if (!hit) return false;
        if (region_start) {                                            // :350-351
            const int64_t pos = pos_of(r, loc, rlen);
            if (pos < region_start[r] || pos >= region_end[r]) return false;
        }
        return true;
    }
};

__global__ void site_columns_kernel(SitePredicate q, const int64_t* __restrict__ chrom_len, const int64_t* __restrict__ site_ev,
                                    const int64_t* __restrict__ n_sel, int64_t cap, int32_t* __restrict__ site_read,
                                    int64_t* __restrict__ pos, int64_t* __restrict__ pos_in_strand) {
    const int64_t n = *n_sel < cap ? *n_sel : cap;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = site_ev[s], r = q.read_of(e);
        const int64_t rlen = q.ev_off[r + 1] - q.ev_off[r], loc = e - q.ev_off[r];
        const int64_t p = q.pos_of(r, loc, rlen);
        const int64_t cl = chrom_len ? chrom_len[r] : -1;
        site_read[s] = (int32_t)r;
        pos[s] = p;
        pos_in_strand[s] = cl < 0 ? -1 : (q.minus[r] ? cl - 1 - p : p);   // :343-348
    }
}

const uint8_t kBase2Code[16] = {'A', 'C', 'G', 'T', 'N', 'W', 'S', 'M', 'K', 'R', 'Y', 'B', 'V', 'D', 'H', 'Z'};

}  // namespace
}  // namespace dsp

using namespace dsp;

namespace dsp {
namespace {
int extract_features_impl(int device, bool f64,
                          const int16_t* raw, const int64_t* raw_off, const double* scaling, const double* offset,
                          int64_t n_reads,
                          const int64_t* ev_start, const int64_t* ev_len, const uint8_t* ev_base,
                          const int32_t* site_read, const int64_t* site_ev, int64_t n_sites,
                          int32_t seq_len, int32_t signal_len, int32_t normalize_method, int32_t round_stats,
                          const int32_t* drawn, uint64_t seed,
                          double* read_shift, double* read_scale,
                          void* kmer, void* base_means, void* base_stds, void* base_signal_lens,
                          void* signals, void* stream) {
    DSP_REQUIRE(n_reads >= 0 && n_sites >= 0, DSP_ERR_INVALID, "dsp_extract_features: negative count");
    DSP_REQUIRE(seq_len > 0 && (seq_len & 1), DSP_ERR_INVALID, "kmer_len must be odd");
    DSP_REQUIRE(seq_len <= 31, DSP_ERR_INVALID, "dsp_extract_features: kmer_len must be at most 31 (one lane per base)");
    DSP_REQUIRE(signal_len > 0 && signal_len <= 128, DSP_ERR_INVALID, "dsp_extract_features: signal_len must be in 1..128");
    DSP_REQUIRE(normalize_method == 0 || normalize_method == 1, DSP_ERR_INVALID, "dsp_extract_features: normalize_method must be 0 (mad) or 1 (zscore)");
    DSP_REQUIRE(read_shift && read_scale, DSP_ERR_INVALID, "dsp_extract_features: read_shift / read_scale buffers are required");
    DSP_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    static int table_device = -1;
    if (table_device != device) {
        uint8_t table[256];
        for (int i = 0; i < 256; ++i) table[i] = 4;           // callers validate the alphabet; 'N' otherwise
        for (int c = 0; c < 16; ++c) table[kBase2Code[c]] = (uint8_t)c;
        DSP_CUDA(cudaMemcpyToSymbolAsync(c_base2code, table, sizeof(table), 0, cudaMemcpyHostToDevice, st));
        DSP_CUDA(cudaStreamSynchronize(st));
        table_device = device;
    }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    if (n_reads > 0 && normalize_method == 1) {
        read_zscore_kernel<<<(unsigned)((n_reads + 3) / 4), 128, 0, st>>>(raw, raw_off, scaling, offset, n_reads, read_shift, read_scale);
        DSP_CUDA(cudaGetLastError());
    } else if (n_reads > 0) {
        const int64_t grid = n_reads < (int64_t)n_sm * 3 ? n_reads : (int64_t)n_sm * 3;   // 3 CTAs of 72 KB per SM
        static int attr_device = -1;                   // function attributes are per device
        if (attr_device != device) {
            DSP_CUDA(cudaFuncSetAttribute(read_scale_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HIST_BINS * (int)sizeof(int)));
            attr_device = device;
        }
        read_scale_kernel<<<(unsigned)grid, SEL_THREADS, HIST_BINS * sizeof(int), st>>>(raw, raw_off, scaling, offset, n_reads, read_shift, read_scale);
        DSP_CUDA(cudaGetLastError());
    }
    if (n_sites > 0) {
        SiteParams p;
        p.raw = raw; p.raw_off = raw_off; p.scaling = scaling; p.offset = offset;
        p.ev_start = ev_start; p.ev_len = ev_len; p.ev_base = ev_base;
        p.site_read = site_read; p.site_ev = site_ev; p.n_sites = n_sites;
        p.T = seq_len; p.S = signal_len; p.round_stats = round_stats;
        p.drawn = drawn; p.seed = seed; p.shift = read_shift; p.scale = read_scale;
        p.kmer = kmer; p.means = base_means; p.stds = base_stds; p.lens = base_signal_lens; p.signals = signals;
        const size_t smem = site_smem_per_warp(seq_len, signal_len) * SITE_WARPS;
        static size_t smem_allowed[2] = {0, 0};
        static int smem_device = -1;
        if (smem_device != device) { smem_allowed[0] = smem_allowed[1] = 0; smem_device = device; }
        auto kern = f64 ? site_features_kernel<double> : site_features_kernel<float>;
        if (smem > smem_allowed[f64]) {
            DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
            DSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            smem_allowed[f64] = smem;
        }
        int64_t grid = (n_sites + SITE_WARPS - 1) / SITE_WARPS;
        if (grid > (int64_t)n_sm * 64) grid = (int64_t)n_sm * 64;
        kern<<<(unsigned)grid, SITE_WARPS * 32, smem, st>>>(p);
        DSP_CUDA(cudaGetLastError());
    }
    return DSP_OK;
}
}  // namespace
}  // namespace dsp

extern "C" int dsp_extract_features(int device,
                                    const int16_t* raw, const int64_t* raw_off, const double* scaling, const double* offset,
                                    int64_t n_reads,
                                    const int64_t* ev_start, const int64_t* ev_len, const uint8_t* ev_base,
                                    const int32_t* site_read, const int64_t* site_ev, int64_t n_sites,
                                    int32_t seq_len, int32_t signal_len, int32_t normalize_method, int32_t round_stats,
                                    const int32_t* drawn, uint64_t seed,
                                    double* read_shift, double* read_scale,
                                    float* kmer, float* base_means, float* base_stds, float* base_signal_lens,
                                    float* signals, void* stream) {
    return extract_features_impl(device, false, raw, raw_off, scaling, offset, n_reads, ev_start, ev_len, ev_base, site_read, site_ev,
                                 n_sites, seq_len, signal_len, normalize_method, round_stats, drawn, seed, read_shift, read_scale,
                                 kmer, base_means, base_stds, base_signal_lens, signals, stream);
}

extern "C" int dsp_extract_features_f64(int device,
                                        const int16_t* raw, const int64_t* raw_off, const double* scaling, const double* offset,
                                        int64_t n_reads,
                                        const int64_t* ev_start, const int64_t* ev_len, const uint8_t* ev_base,
                                        const int32_t* site_read, const int64_t* site_ev, int64_t n_sites,
                                        int32_t seq_len, int32_t signal_len, int32_t normalize_method, int32_t round_stats,
                                        const int32_t* drawn, uint64_t seed,
                                        double* read_shift, double* read_scale,
                                        double* kmer, double* base_means, double* base_stds, double* base_signal_lens,
                                        double* signals, void* stream) {
    return extract_features_impl(device, true, raw, raw_off, scaling, offset, n_reads, ev_start, ev_len, ev_base, site_read, site_ev,
                                 n_sites, seq_len, signal_len, normalize_method, round_stats, drawn, seed, read_shift, read_scale,
                                 kmer, base_means, base_stds, base_signal_lens, signals, stream);
}

extern "C" int dsp_find_sites(int device, const uint8_t* ev_base, const int64_t* ev_off, int64_t n_reads, int64_t n_events,
                              const char* motifs, int32_t n_motifs, int32_t motif_len, int32_t methyloc, int32_t seq_len,
                              const int64_t* chrom_start, const uint8_t* minus_strand, const int64_t* chrom_len,
                              const int64_t* region_start, const int64_t* region_end, int64_t max_sites,
                              int32_t* site_read, int64_t* site_ev, int64_t* pos, int64_t* pos_in_strand,
                              int64_t* n_sites_host, void* stream) {
    DSP_REQUIRE(n_sites_host, DSP_ERR_INVALID, "dsp_find_sites: n_sites_host is required");
    *n_sites_host = 0;
    DSP_REQUIRE(seq_len > 0 && (seq_len & 1), DSP_ERR_INVALID, "kmer_len must be odd");
    DSP_REQUIRE(n_motifs >= 1 && n_motifs <= MAX_MOTIFS && motif_len >= 1 && motif_len <= MAX_MOTIF_LEN && motifs, DSP_ERR_INVALID,
                "dsp_find_sites: 1..%d motifs of 1..%d letters", MAX_MOTIFS, MAX_MOTIF_LEN);
    DSP_REQUIRE(n_reads >= 0 && n_events >= 0 && max_sites >= 0 && n_events < (1ll << 31), DSP_ERR_INVALID, "dsp_find_sites: bad count");
    DSP_REQUIRE((region_start == nullptr) == (region_end == nullptr), DSP_ERR_INVALID, "dsp_find_sites: region_start and region_end go together");
    if (n_reads == 0 || n_events == 0) return DSP_OK;
    DSP_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    SitePredicate q;
    q.ev_base = ev_base; q.ev_off = ev_off; q.n_reads = n_reads; q.chrom_start = chrom_start; q.minus = minus_strand;
    q.region_start = region_start; q.region_end = region_end;
    q.n_motifs = n_motifs; q.motif_len = motif_len; q.methyloc = methyloc; q.num_bases = (seq_len - 1) / 2;
    for (int m = 0; m < MAX_MOTIFS; ++m) q.motif[m] = 0;
    for (int m = 0; m < n_motifs; ++m)
        for (int k = 0; k < motif_len; ++k) q.motif[m] |= (unsigned long long)(uint8_t)motifs[m * motif_len + k] << (8 * k);
    // events in ascending order = reads in order, positions ascending within a read: the reference's site order
    cub::CountingInputIterator<int64_t> events(0);
    int64_t* d_count = nullptr;
    int64_t* d_sel = nullptr;                           // all hits (n_events at most) before the max_sites cut
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    DSP_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, events, d_sel, d_count, (int)n_events, q, st));
    DSP_CUDA(cudaMallocAsync((void**)&d_count, sizeof(int64_t), st));
    DSP_CUDA(cudaMallocAsync((void**)&d_sel, sizeof(int64_t) * (size_t)n_events, st));
    DSP_CUDA(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 1, st));
    cudaError_t e = cub::DeviceSelect::If(d_tmp, tmp_bytes, events, d_sel, d_count, (int)n_events, q, st);
    int64_t n = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&n, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && n > 0 && n <= max_sites) {
        e = cudaMemcpyAsync(site_ev, d_sel, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) {
            site_columns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, chrom_len, site_ev, d_count, max_sites, site_read, pos, pos_in_strand);
            e = cudaGetLastError();
        }
    }
    cudaFreeAsync(d_tmp, st); cudaFreeAsync(d_sel, st); cudaFreeAsync(d_count, st);
    if (e != cudaSuccess) { set_error("dsp_find_sites: %s", cudaGetErrorString(e)); return DSP_ERR_CUDA; }
    *n_sites_host = n;
    DSP_REQUIRE(n <= max_sites, DSP_ERR_NOMEM, "dsp_find_sites: %lld sites found, buffers hold %lld", (long long)n, (long long)max_sites);
    return DSP_OK;
}
