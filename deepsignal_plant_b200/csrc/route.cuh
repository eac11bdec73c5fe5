// Stable multi-way partition of fixed-size items into up to 16 destination buffers, as three
// kernels (count per block -> scan -> scatter).  Used with ONE destination as the stream compaction
// of callable records on a single GPU (freq.cu) and with `world` destinations -- peer memory windows
// mapped over NVLink -- as the fused partition + all-to-all of the multi-GPU call_freq (comm.cu):
// the scatter kernel stores every item straight into the receiving GPU's window at its final
// position, so the exchange overlaps the partition store by store and no send buffer exists.
//
// Stability is the point: items bound for the same destination keep their order (file order of the
// per-read calls), because the float64 sums of call_mods_freq.py:60-61 are order sensitive.
// HBM/NVLink-bound byte work: every item is read twice (count, scatter) and written once.
#pragma once
#include "freq.cuh"

namespace dsp {
namespace route {

constexpr int MAXW = 16;                      // ranks of a communicator (comm.cu)
constexpr int MAXD = 32;                      // destinations of one partition; = warp size (one lane scans one destination)
constexpr int RT = 256;                       // threads per block
constexpr int RWARPS = RT / 32;

struct Plan {
    int64_t n = 0;
    int64_t per_block = 0;                    // items per block (multiple of RT)
    int blocks = 0;
};
inline Plan make_plan(int64_t n, int n_sm) {
    Plan p;
    p.n = n;
    const int64_t want = (int64_t)(n_sm > 0 ? n_sm : 148) * 8;
    int64_t per = (n + want - 1) / want;
    if (per < 4096) per = 4096;
    per = (per + RT - 1) / RT * RT;
    p.per_block = per;
    p.blocks = (int)((n + per - 1) / per);
    if (p.blocks < 1) p.blocks = 1;
    return p;
}

__host__ __device__ __forceinline__ int owner_of_key(uint64_t key, int world) {
    // multiplicative hash of the 64-bit site key, top 31 bits, mapped onto [0, world) by multiply-shift (no division);
    // call_mods_freq.owner_of_key is the same function on the host
    return (int)((((key * 0x9E3779B97F4A7C15ull) >> 33) * (uint64_t)world) >> 31);
}

// ---- sources --------------------------------------------------------------------------------------
// A source turns item index i into (keep?, destination, payload).
struct RecFromColumns {                       // call_mods columns -> packed Rec, by key hash
    typedef Rec Item;
    const uint64_t* key; const double* p0; const double* p1; const int32_t* label;
    uint64_t gidx_base; double prob_cf; int world;
    __device__ __forceinline__ bool dest_of(int64_t i, int& d, uint64_t& k) const {
        if (fabs(p0[i] - p1[i]) < prob_cf) return false;               // txt_formater.py:23-26
        if (world == 1) { d = 0; k = 0; return true; }                // one destination: the key is not needed to count
        k = key[i];
        d = owner_of_key(k, world);
        return true;
    }
    __device__ __forceinline__ void load(int64_t i, Item& it) const {
        it.key = key[i]; it.p0 = p0[i]; it.p1 = p1[i];
        it.gl = (gidx_base + (uint64_t)i) | (label[i] == 1 ? REC_LABEL_BIT : 0ull);
    }
    __device__ __forceinline__ void store(Item* dst, int64_t pos, const Item& it) const { store_rec(dst + pos, it); }
    // everything of item i in one go (one round of loads): keep?, destination, payload
    __device__ __forceinline__ bool load_all(int64_t i, Item& it, int& d) const {
        it.key = key[i]; it.p0 = p0[i]; it.p1 = p1[i];
        it.gl = (gidx_base + (uint64_t)i) | (label[i] == 1 ? REC_LABEL_BIT : 0ull);
        d = world == 1 ? 0 : owner_of_key(it.key, world);
        return !(fabs(it.p0 - it.p1) < prob_cf);
    }
};

template <int UNITS>                           // rows of UNITS x 16 bytes, by range of a 64-bit field
struct RowsByRange {
    struct __align__(16) Item { uint4 u[UNITS]; };
    const Item* rows; int field_word;           // index of the 64-bit field inside the row
    const uint64_t* bounds; int world;          // world + 1 ascending values in DEVICE memory
    __device__ __forceinline__ bool dest_of(int64_t i, int& d, uint64_t& k) const {
        k = reinterpret_cast<const uint64_t*>(rows + i)[field_word];
        int lo = 0;
        for (int w = 1; w < world; ++w) lo += (k >= bounds[w]) ? 1 : 0;     // bounds ascending
        d = lo;
        return true;
    }
    __device__ __forceinline__ void load(int64_t i, Item& it) const { it = rows[i]; }
    __device__ __forceinline__ void store(Item* dst, int64_t pos, const Item& it) const { dst[pos] = it; }
    __device__ __forceinline__ bool load_all(int64_t i, Item& it, int& d) const {
        it = rows[i];
        const uint64_t k = reinterpret_cast<const uint64_t*>(&it)[field_word];
        int lo = 0;
        for (int w = 1; w < world; ++w) lo += (k >= bounds[w]) ? 1 : 0;
        d = lo;
        return true;
    }
};

// ---- kernels ----------------------------------------------------------------------------------------
template <typename Src>
__global__ void __launch_bounds__(RT) count_kernel(Src src, Plan plan, int world, int32_t* __restrict__ blk_counts) {
    __shared__ int s_cnt[MAXD];
    if (threadIdx.x < MAXD) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * plan.per_block;
    const int64_t hi = min(plan.n, lo + plan.per_block);
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = lo; i0 < hi; i0 += RT) {
        const int64_t i = i0 + threadIdx.x;
        int d = -1; uint64_t k;
        if (i < hi) { int dd; if (src.dest_of(i, dd, k)) d = dd; }
        // one vote per warp: the lanes bound for the same destination find each other, their leader adds the group
        const unsigned m = __match_any_sync(0xffffffffu, d);
        if (d >= 0 && lane == __ffs(m) - 1) atomicAdd(&s_cnt[d], __popc(m));
    }
    __syncthreads();
    if (threadIdx.x < world) blk_counts[(size_t)blockIdx.x * MAXD + threadIdx.x] = s_cnt[threadIdx.x];
}

// one warp per destination: exclusive scan of the per-block counts; totals[d] = items bound for d
static __global__ void __launch_bounds__(MAXD * 32) scan_kernel(const int32_t* __restrict__ blk_counts, int blocks, int world,
                                                        int64_t* __restrict__ blk_off, int64_t* __restrict__ totals) {
    const int d = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (d >= world) return;
    int64_t carry = 0;
    for (int b0 = 0; b0 < blocks; b0 += 32) {
        const int b = b0 + lane;
        const int64_t v = b < blocks ? blk_counts[(size_t)b * MAXD + d] : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (b < blocks) blk_off[(size_t)b * MAXD + d] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) totals[d] = carry;
}

struct Targets {
    void* dst[MAXD];                          // destination buffers (this GPU's or a peer's window)
    int64_t base[MAXD];                       // first position of THIS source's segment in dst[d]
};

// Items of block b bound for d land at dst[d][base[d] + blk_off[b][d] + rank inside the block], in order.
// `tg` is read from device memory: on the multi-GPU path the bases are computed on the device from the
// count matrix the ranks publish to each other (no host round trip between count and scatter).
template <typename Src>
__global__ void __launch_bounds__(RT) scatter_kernel(Src src, Plan plan, int world, const int64_t* __restrict__ blk_off,
                                                    const Targets* __restrict__ tg, const int* __restrict__ abort_flag,
                                                    unsigned long long* __restrict__ key_bits_out) {
    typedef typename Src::Item Item;
    __shared__ int s_warp[RWARPS][MAXD];
    __shared__ int64_t s_run[MAXD];
    __shared__ Item* s_dst[MAXD];
    if (abort_flag && *abort_flag) return;                               // a window would overflow: nobody writes
    if (threadIdx.x < MAXD) {
        const int d = threadIdx.x;
        s_run[d] = d < world ? tg->base[d] + blk_off[(size_t)blockIdx.x * MAXD + d] : 0;
        s_dst[d] = d < world ? reinterpret_cast<Item*>(tg->dst[d]) : nullptr;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t lo = (int64_t)blockIdx.x * plan.per_block;
    const int64_t hi = min(plan.n, lo + plan.per_block);
    uint64_t bits = 0;
    for (int64_t t0 = lo; t0 < hi; t0 += RT) {
        const int64_t i = t0 + threadIdx.x;
        int d = -1; uint64_t k;
        bool keep = false;
        if (i < hi) { int dd; keep = src.dest_of(i, dd, k); if (keep) d = dd; }
        for (int e = lane; e < world; e += 32) s_warp[warp][e] = 0;
        __syncwarp();
        const unsigned m = __match_any_sync(0xffffffffu, d);       // the lanes of this warp bound for the same destination
        const int rank = __popc(m & lt);
        if (d >= 0 && rank == 0) s_warp[warp][d] = __popc(m);
        __syncthreads();
        if (keep) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += s_warp[w][d];
            Item it;
            src.load(i, it);
            bits |= *reinterpret_cast<const uint64_t*>(&it);            // first word of every item is its key
            src.store(s_dst[d], s_run[d] + before + rank, it);
        }
        __syncthreads();
        if (threadIdx.x < world) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < RWARPS; ++w) tot += s_warp[w][threadIdx.x];
            s_run[threadIdx.x] += tot;
        }
        __syncthreads();
    }
    if (key_bits_out) {
        const unsigned blo = __reduce_or_sync(0xffffffffu, (unsigned)bits), bhi = __reduce_or_sync(0xffffffffu, (unsigned)(bits >> 32));
        if (lane == 0 && (blo | bhi)) atomicOr(key_bits_out, ((unsigned long long)bhi << 32) | blo);
    }
}

// The same scatter for 32-byte records / 48-byte rows bound for PEER windows: a tile of ST x RT items is ranked, staged in shared
// memory grouped by destination, and copied out with consecutive threads writing consecutive records -- every warp
// store covers 1 KB of one destination window instead of ~4 records for each of `world` windows, which is what NVLink
// wants (few large writes instead of many 32-byte ones).  Same positions, same order as scatter_kernel.
template <typename Src>
__global__ void __launch_bounds__(RT, 4) scatter_staged_kernel(Src src, Plan plan, int world, const int64_t* __restrict__ blk_off,
                                                           const Targets* __restrict__ tg, const int* __restrict__ abort_flag,
                                                           unsigned long long* __restrict__ key_bits_out) {
    typedef typename Src::Item Item;
    constexpr int ST = sizeof(Item) <= 32 ? 4 : 3;          // sub-tiles of RT items per staged tile (stage <= 36 KB)
    static_assert(sizeof(Item) <= 48, "the stage of ST x RT items must fit the static shared memory");
    __shared__ __align__(32) Item s_stage[ST * RT];
    __shared__ int s_warp[ST][RWARPS][MAXD];
    __shared__ int s_sub[ST][MAXD];            // items of sub-tile j bound for d
    __shared__ int s_seg[MAXD + 1];            // first stage slot of destination d in this tile
    __shared__ uint8_t s_sd[ST * RT];          // destination of every stage slot
    __shared__ int64_t s_run[MAXD];
    __shared__ Item* s_dst[MAXD];
    if (abort_flag && *abort_flag) return;
    if (threadIdx.x < MAXD) {
        const int d = threadIdx.x;
        s_run[d] = d < world ? tg->base[d] + blk_off[(size_t)blockIdx.x * MAXD + d] : 0;
        s_dst[d] = d < world ? reinterpret_cast<Item*>(tg->dst[d]) : nullptr;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t lo = (int64_t)blockIdx.x * plan.per_block;
    const int64_t hi = min(plan.n, lo + plan.per_block);
    uint64_t bits = 0;
    for (int64_t t0 = lo; t0 < hi; t0 += (int64_t)ST * RT) {
        // phase 1: all loads of the tile in flight together (no synchronisation between them)
        Item item[ST];
        int dest[ST], rank[ST];
#pragma unroll
        for (int j = 0; j < ST; ++j) {
            const int64_t i = t0 + (int64_t)j * RT + threadIdx.x;
            int d = -1;
            if (i < hi) { int dd; if (src.load_all(i, item[j], dd)) d = dd; }
            dest[j] = d;
        }
        for (int e = threadIdx.x; e < ST * RWARPS * MAXD; e += RT) (&s_warp[0][0][0])[e] = 0;
        __syncthreads();
        // phase 2: stable ranks -- lanes bound for the same destination find each other, their first lane counts the group
#pragma unroll
        for (int j = 0; j < ST; ++j) {
            const unsigned m = __match_any_sync(0xffffffffu, dest[j]);
            rank[j] = __popc(m & lt);
            if (dest[j] >= 0 && rank[j] == 0) s_warp[j][warp][dest[j]] = __popc(m);
        }
        __syncthreads();
        if (threadIdx.x < ST * MAXD) {
            const int j = threadIdx.x / MAXD, d = threadIdx.x % MAXD;
            int tot = 0;
            if (d < world)
#pragma unroll
                for (int w = 0; w < RWARPS; ++w) tot += s_warp[j][w][d];
            s_sub[j][d] = tot;
        }
        __syncthreads();
        if (warp == 0) {                                       // exclusive scan over the destinations, one lane each
            int tot = 0;
            if (lane < world)
#pragma unroll
                for (int j = 0; j < ST; ++j) tot += s_sub[j][lane];
            int x = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            s_seg[lane] = x - tot;
            if (lane == 31) s_seg[MAXD] = x;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ST; ++j) {
            const int d = dest[j];
            if (d >= 0) {
                int slot = s_seg[d] + rank[j];
                for (int jj = 0; jj < j; ++jj) slot += s_sub[jj][d];
                for (int w = 0; w < warp; ++w) slot += s_warp[j][w][d];
                bits |= *reinterpret_cast<const uint64_t*>(&item[j]);
                s_stage[slot] = item[j];
                s_sd[slot] = (uint8_t)d;
            }
        }
        __syncthreads();
        const int total = s_seg[MAXD];
        for (int sl = threadIdx.x; sl < total; sl += RT) {
            const int d = s_sd[sl];
            src.store(s_dst[d], s_run[d] + (sl - s_seg[d]), s_stage[sl]);
        }
        __syncthreads();
        if (threadIdx.x < world) s_run[threadIdx.x] += s_seg[threadIdx.x + 1] - s_seg[threadIdx.x];
        __syncthreads();
    }
    if (key_bits_out) {
        const unsigned blo = __reduce_or_sync(0xffffffffu, (unsigned)bits), bhi = __reduce_or_sync(0xffffffffu, (unsigned)(bits >> 32));
        if (lane == 0 && (blo | bhi)) atomicOr(key_bits_out, ((unsigned long long)bhi << 32) | blo);
    }
}

template <typename Src, bool STAGED = (sizeof(typename Src::Item) <= 48)>
struct ScatterLaunch {
    static void run(const Src& src, const Plan& plan, int world, const int64_t* blk_off, const Targets* tg, const int* abort_flag,
                    unsigned long long* bits, cudaStream_t st) {
        scatter_kernel<Src><<<plan.blocks, RT, 0, st>>>(src, plan, world, blk_off, tg, abort_flag, bits);
    }
};
template <typename Src>
struct ScatterLaunch<Src, true> {
    static void run(const Src& src, const Plan& plan, int world, const int64_t* blk_off, const Targets* tg, const int* abort_flag,
                    unsigned long long* bits, cudaStream_t st) {
        if (world > 1) scatter_staged_kernel<Src><<<plan.blocks, RT, 0, st>>>(src, plan, world, blk_off, tg, abort_flag, bits);
        else scatter_kernel<Src><<<plan.blocks, RT, 0, st>>>(src, plan, world, blk_off, tg, abort_flag, bits);
    }
};

}  // namespace route
}  // namespace dsp
