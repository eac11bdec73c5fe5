// sm_100a building blocks used by the tensor-core kernels: mbarrier, bulk async copy (TMA
// engine, cp.async.bulk), tcgen05 (TMEM alloc, MMA with shared-memory or TMEM A operand,
// commit, ld/st), UMMA descriptors and the 128-byte swizzle.  Inline PTX only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace dsp {
namespace tc {

// ---- operand tile images -------------------------------------------------------------------
// Every MMA operand is stored, in global memory and in shared memory alike, as "slabs":
// a slab is R rows x 64 fp16 (128 bytes per row, K-major) laid out as consecutive 8-row
// groups of 1024 bytes with the 128-byte swizzle applied inside each group: the 16-byte
// chunk c (0..7) of row r sits at chunk position c ^ (r & 7).  This is exactly the canonical
// SWIZZLE_128B K-major layout tcgen05.mma reads, so a slab travels from HBM to shared
// memory with one linear cp.async.bulk and needs no tensor map.
constexpr int SLAB_K = 64;                       // fp16 elements per slab row
constexpr int SLAB_ROW_BYTES = 128;

__host__ __device__ inline uint32_t slab_offset_bytes(uint32_t row, uint32_t col /* fp16 index 0..63 */) {
    const uint32_t chunk = col >> 3;
    return row * SLAB_ROW_BYTES + (((chunk ^ (row & 7u)) << 4) | ((col & 7u) << 1));
}

// ---- shared-memory address helpers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// ---- bulk async copy global -> shared (TMA engine), completion on an mbarrier ---------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}

// same copy delivered to the same shared-memory offset of every CTA in cta_mask; each
// destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar,
                                                   uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}

// ---- thread-block cluster ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- programmatic dependent launch: a kernel launched with programmaticStreamSerialization may start
// (prologue, TMEM allocation, initial-state draws) while the previous kernel of the stream drains;
// griddep_wait() blocks until that kernel has completed and its writes are visible.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- proxy / tcgen05 fences ----------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these) -----------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
// (stride byte offset), leading byte offset unused for swizzled K-major, descriptor version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)(1024u >> 4) << 32;                     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // version, bits [46,48)
    d |= (uint64_t)2 << 61;                                // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// The same descriptor split in halves: only the start-address field (low word) changes
// between MMAs, so issue loops keep the high word constant and add to the low word
// (+2 per 32 bytes of K, +1024 per 16 KB slab).
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr) { return (smem_addr & 0x3FFFFu) >> 4; }
constexpr uint32_t SMEM_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
static_assert(((uint64_t)SMEM_DESC_HI << 32) == (((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61)),
              "descriptor high word");

// Instruction descriptor for kind::f16: FP16 A/B (K-major both), FP32 accumulate, M x N.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4)                       // c_format = F32
           | (0u << 7) | (0u << 10)        // a_format = b_format = F16
           | (0u << 15) | (0u << 16)       // a_major = b_major = K
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA ---------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T  (A: lane = row, 32-bit column = two consecutive K values)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// descriptor-low-word forms with a compile-time accumulate flag (no predicate register traffic)
template <int ACCUM>
__device__ __forceinline__ void mma_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %4};\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(SMEM_DESC_HI), "n"(ACCUM) : "memory");
}
template <int ACCUM>
__device__ __forceinline__ void mma_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(SMEM_DESC_HI), "n"(ACCUM) : "memory");
}
// mbarrier arrives when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// same, arriving on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on one TPC execute one MMA of M = 256 ---------
// Each CTA contributes its own 128 A rows (shared memory or TMEM, same offsets in both CTAs) and
// HALF of the B rows (N/2 each, same shared-memory offset in both CTAs); each CTA receives its
// 128 accumulator rows x N columns in its own TMEM.  Only the leader CTA (cluster rank 0) issues.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int ACCUM>
__device__ __forceinline__ void mma2_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %4};\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(SMEM_DESC_HI), "n"(ACCUM) : "memory");
}
template <int ACCUM>
__device__ __forceinline__ void mma2_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(SMEM_DESC_HI), "n"(ACCUM) : "memory");
}
// completion of all MMAs issued so far by this thread -> one arrival on the barrier at this
// shared-memory offset in every CTA of cta_mask
__device__ __forceinline__ void mma2_commit(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
// address of the same shared-memory location in another CTA of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
// arrive on a barrier that may live in another CTA (address from mapa_u32)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local barrier whose arrivals may come from another CTA of the cluster.  What these
// barriers order is tensor-memory and async-proxy traffic (fenced with tcgen05.fence / completed
// through the barrier itself), not generic-proxy stores, so CTA-scope semantics suffice and no
// cluster-scope MEMBAR lands in the hot loops.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }

__device__ __forceinline__ float4 lds128(uint32_t smem_addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr));
    return v;
}

// ---- packed fp32 pairs (FADD2 / FMUL2 / FFMA2): one issue slot for two lanes of work ----------------
#ifdef DSP_NO_PACK   // measurement build: scalar FADD/FMUL/FFMA (either FMA pipe) instead of the packed forms (heavy pipe only)
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
#endif

// ---- TMEM <-> registers (32 lanes x 32-bit, N consecutive columns per thread) -----------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace tc
}  // namespace dsp
