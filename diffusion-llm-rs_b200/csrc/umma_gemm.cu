// umma_gemm.cu — tcgen05 / TMEM dequant-GEMM (K5) and skinny dequant-GEMV (K4 fast path).
//
//   y[M,N] = x[M,K] · dequant(W) + b            (diffuse-llm-rs/src/lib.rs:812 composed with
//                                                 dequantize_tensor, quantization.rs:81-85)
//
// Formulated "weights-as-A":   D[n, tok] += A[n, k] · B[tok, k]^T
//   A (128 output columns x 64 k per step): 2/4/8-bit codes streamed from HBM with one bulk
//      async copy per tile (tile-major layout, wlayout.cuh), dequantized IN REGISTERS to bf16
//      ((q - zp) * scale, per group of 128 k) by the dequant warps and written straight into
//      TENSOR MEMORY with tcgen05.st — the dequantized weights never touch shared memory or HBM.
//   B (NTOK tokens x 64 k): bf16 activations, TMA-loaded (SWIZZLE_128B, K-major) into smem.
//   D: f32 accumulators in TMEM (double buffered), read back with tcgen05.ld by the epilogue warps.
// tcgen05.mma.cta_group::1.kind::f16 with A from TMEM ("TS" form), M=128, N=NTOK, K=16.
//
// Warp roles (one persistent CTA per SM, (8 + 4*NDQ) warps):
//   warp 0      TMA / bulk-copy producer (one elected lane): x tile (TMA), packed weight tile and the
//               tile's 128 (zero-point, scale) dequant operands (bulk copies) per stage
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 4..   NDQ dequant groups of 4 warps (warp%4 = TMEM lane quarter); group g takes the
//               pipeline stages with stage % NDQ == g
//   last 4      epilogue: TMEM -> registers -> (+bias) -> global (f32 and/or bf16), coalesced
// Two rings: shared-memory stages (deep: they cover the HBM latency; KBS k-blocks per stage) and
//   TMEM A slots (shallow: they cover dequant -> MMA).  full[s] (TMA bytes landed) ->
//   afull[a] (A slot written) -> MMA -> tcgen05.commit -> sempty[s] + aempty[a];
//   tmem_full / tmem_empty between MMA and epilogue.
// Scheduling: dense shapes take whole output tiles round-robin; skinny shapes use stream-K — every CTA
//   gets the same number of contiguous k-block units, so all 148 SMs stream weights for the same
//   time; tiles cut by a CTA boundary emit f32 partials that a tiny fix-up kernel sums in CTA order
//   (deterministic, no atomics).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "wlayout.cuh"

namespace {

#ifndef DLLM_NDQ
#define DLLM_NDQ 3
#endif
constexpr int kNDQ = DLLM_NDQ;       // dequant groups of 4 warps
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of a TPC run ONE 256-row MMA.  Each CTA dequantizes its
// own 128 output columns into its own tensor memory and loads HALF of the activation tile (the B operand is read from both
// CTAs' shared memory), so the activation traffic per SM halves.  Barriers the leader (rank 0) waits on collect arrivals
// from both CTAs; completions it produces (tcgen05.commit) are multicast to both.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p`'s twin in the leader CTA
__device__ __forceinline__ uint32_t leader_addr(const void *p) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(a) : "r"(smem_u32(p)));
    return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_bar) : "memory");
}
// The same arrive without the cluster-scope release (which compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive): for hand-offs
// whose payload lives in tensor memory only — the tcgen05.wait / tcgen05.fence::before_thread_sync in front of it order the tensor-
// memory accesses, and no generic-proxy write has to be published.  (Both variants of the 256-token CTA-pair kernel.)
__device__ __forceinline__ void mbar_arrive_cluster_tmem(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_bar) : "memory");
}
// both CTAs load their half of the tile into their own shared memory; the bytes are reported to the LEADER's barrier
__device__ __forceinline__ void tma_load_3d_pair(void *smem_dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
// D[tmem of both CTAs, 256 rows] (+)= A[tmem of both CTAs] * B[smem of both CTAs]
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with kind::i8 (u8 x s8 operands, s32 accumulate: exact)
__device__ __forceinline__ void umma_ts_pair_i8(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}\n" :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::i8 with A from shared memory (both operands through descriptors)
[[maybe_unused]] __device__ __forceinline__ void umma_ss_pair_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// one lane of a fully converged warp (keeps the surrounding code warp-uniform so that descriptors and
// addresses live in uniform registers: no per-instruction R2UR waterfall in the MMA / TMA issue loops)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]      (kind::f16: bf16 operands, f32 accumulate)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]      (kind::i8: u8 x s8 operands, s32 accumulate: exact)
__device__ __forceinline__ void umma_ts_i8(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}\n" :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
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
template <int CH>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t *r) {
    static_assert(CH == 16 || CH == 32 || CH == 64, "chunk");
    if (CH == 16) tmem_ld16(taddr, r);
    else if (CH == 32) tmem_ld32(taddr, r);
    else { tmem_ld32(taddr, r); tmem_ld32(taddr + 32, r + 32); }   // two loads in flight, one wait
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// smem matrix descriptor of a K-major, SWIZZLE_128B bf16 tile (rows of 128 bytes, 8-row groups of
// 1024 bytes): start>>4 | LBO(=1)<<16 | SBO(=1024>>4)<<32 | version(=1)<<46 | layout(=2: SW128)<<61
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), K-major A and B, N>>3 at 17, M>>4 at 24.
// (Measured: kind::f16 traps with cudaErrorIllegalInstruction when the A and B formats differ, so fp16 weights —
// whose unpack is cheaper, see gemv_mma.cu — would need fp16 activations too; the stack keeps bf16 for its range.)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// the same for the 256-row MMA of a CTA pair
__host__ __device__ constexpr uint32_t make_idesc_pair(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
// kind::i8: c = s32 (2<<4), a = u8 (0<<7: the codes as they are), b = s8 (1<<10), K-major A and B
__host__ __device__ constexpr uint32_t make_idesc_i8(int n) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// dequantize one k-block (64 k) of one output column to 32 packed bf16 pairs: (q - zp) * scale.
// q - zp is exact in bf16 (small integers); the product rounds once (the scale is bf16-rounded).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf2_sub_mul(uint32_t a, uint32_t zb, uint32_t s2) {
    __nv_bfloat162 d = __hsub2(*reinterpret_cast<__nv_bfloat162 *>(&a), *reinterpret_cast<__nv_bfloat162 *>(&zb));
    __nv_bfloat162 r = __hmul2(d, *reinterpret_cast<__nv_bfloat162 *>(&s2));
    return *reinterpret_cast<uint32_t *>(&r);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
// (a & mask) | magic in ONE LOP3 (with two immediates the compiler emits two; +3.5% on the whole denoise step)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t mask, uint32_t magic) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(magic));
    return d;
}
// `zterm`/`s2` are the per-(group, column) operands prepared by wdq_params_kernel.
template <int CB>
__device__ __forceinline__ void dequant_kblock(const uint4 *wpk, int n_local, uint32_t zterm, uint32_t s2, uint32_t *out) {
    constexpr int CH = CB / 2;
    const uint32_t wbase = smem_u32(wpk) + (uint32_t)n_local * 16u;      // explicit ld.shared (the pointer form compiled to generic LD.E.128)
    if (CB == 4 || CB == 2) {
        // magic: 0x4300 | q is the bf16 number 128 + q (exact for q < 128); subtract bf16x2(128 + zp)
        const uint32_t zb = zterm;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const uint4 c = lds128(wbase + (uint32_t)(j * 128) * 16u);
            const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int wd = 0; wd < 4; ++wd) {
                if (CB == 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        out[j * 16 + wd * 4 + i] = bf2_sub_mul(and_or(w[wd] >> (4 * i), 0x000f000fu, 0x43004300u), zb, s2);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        out[wd * 8 + i] = bf2_sub_mul(and_or(w[wd] >> (2 * i), 0x00030003u, 0x43004300u), zb, s2);
                }
            }
        }
    } else {
        // 8-bit: q and zp up to 255 — exact in f32, (q - zp) exact in bf16 after the subtraction
        const float zp = __uint_as_float(zterm);
        const __nv_bfloat162 sb = *reinterpret_cast<const __nv_bfloat162 *>(&s2);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const uint4 c = lds128(wbase + (uint32_t)(j * 128) * 16u);
            const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int wd = 0; wd < 4; ++wd) {
                // 0x4B000000 | byte == 8388608.0f + q
                const float q0 = __uint_as_float(__byte_perm(w[wd], 0x4B000000u, 0x7650)) - 8388608.0f;
                const float q1 = __uint_as_float(__byte_perm(w[wd], 0x4B000000u, 0x7651)) - 8388608.0f;
                const float q2 = __uint_as_float(__byte_perm(w[wd], 0x4B000000u, 0x7652)) - 8388608.0f;
                const float q3 = __uint_as_float(__byte_perm(w[wd], 0x4B000000u, 0x7653)) - 8388608.0f;
                __nv_bfloat162 d01 = __floats2bfloat162_rn(q0 - zp, q1 - zp);
                __nv_bfloat162 d23 = __floats2bfloat162_rn(q2 - zp, q3 - zp);
                __nv_bfloat162 r01 = __hmul2(d01, sb), r23 = __hmul2(d23, sb);
                out[j * 8 + wd * 2 + 0] = *reinterpret_cast<uint32_t *>(&r01);
                out[j * 8 + wd * 2 + 1] = *reinterpret_cast<uint32_t *>(&r23);
            }
        }
    }
}

// int8 mode: one k-block (64 k) of one output column as 16 words of four u8 codes in k order (the A operand of
// kind::i8 reads its 32-byte K chunk from 8 consecutive TMEM columns of the row) — no zero-point, no scale: the integer
// GEMM is exact and the zero-point leaves in the epilogue as zp * rowsum(x).
template <int CB>
__device__ __forceinline__ void unpack_kblock_u8(const uint4 *wpk, int n_local, uint32_t *out) {
    constexpr int CH = CB / 2;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const uint4 c = wpk[j * 128 + n_local];
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) {
            if (CB == 4) {
                // codes 0..7 at nibbles {0,4,1,5,2,6,3,7}: w & 0x0f.. = bytes (c0,c4,c1,c5), (w >> 4) & 0x0f.. = (c2,c6,c3,c7)
                const uint32_t lo = w[wd] & 0x0f0f0f0fu, hi = (w[wd] >> 4) & 0x0f0f0f0fu;
                out[j * 8 + wd * 2 + 0] = __byte_perm(lo, hi, 0x6420);
                out[j * 8 + wd * 2 + 1] = __byte_perm(lo, hi, 0x7531);
            } else if (CB == 2) {
                // code i at field (i >> 1) + 8 (i & 1): (w >> 2 s) & 0x03.. = bytes (2s, 8+2s, 2s+1, 9+2s)
                const uint32_t r0 = w[wd] & 0x03030303u, r1 = (w[wd] >> 2) & 0x03030303u;
                const uint32_t r2 = (w[wd] >> 4) & 0x03030303u, r3 = (w[wd] >> 6) & 0x03030303u;
                out[wd * 4 + 0] = __byte_perm(r0, r1, 0x6420);
                out[wd * 4 + 1] = __byte_perm(r2, r3, 0x6420);
                out[wd * 4 + 2] = __byte_perm(r0, r1, 0x7531);
                out[wd * 4 + 3] = __byte_perm(r2, r3, 0x7531);
            } else {
                out[j * 4 + wd] = w[wd];
            }
        }
    }
}

struct UmmaArgs {
    const uint8_t *packed;
    const uint2 *dqparams;     // [G][Npad] {zero-point term, bf16x2 scale}
    const float *bias;
    float *y_f32;              // final f32 output [M,N] or null
    __nv_bfloat16 *y_bf16;     // final bf16 output [M,N] or null
    float *partial;            // stream-K partial tiles: [2 * grid][NTOK][128] f32
    uint32_t M, N, Npad, k_blocks, n_tiles, m_tiles, group_kb;
    uint32_t x3d;              // activation tensor map is the 3-D {64, M, K/64} view (K % 64 == 0)
    uint32_t stream_k;         // 0: whole tiles round-robin; 1: contiguous unit ranges per CTA
    long long *trace;          // dbg & 128: per-stage clock64 stamps of CTA 0: [role 0..7][256]
    uint32_t dbg;              // timing experiments only (DLLM_UMMA_DBG): 1 skip MMAs, 2 skip dequant math, 4 skip TMEM stores
    uint64_t units;            // n_tiles * m_tiles * k_blocks
    // int8 mode reuses three fields (a larger parameter block costs the bf16 instances registers: they start to spill):
    //   y_f32 = the exact int32 output [M,N], partial = row sums of the int8 activations [M] (int32), x3d = the zero-point
};

// One piece of work of a CTA: k-blocks [kb0, kb1) of output tile `tile`.
#define TRACE(role, idx) do { if ((a.dbg & 128) && blockIdx.x == 0 && (idx) < 256) a.trace[(role) * 256 + (idx)] = clock64(); } while (0)

struct Item {
    uint32_t tile, kb0, kb1;
    int32_t slot;              // -1: complete tile, written directly; else partial slot index
};

// Every warp role walks the same item sequence.
struct ItemIter {
    uint32_t stream_k, KB, tiles, cur_tile, first;
    uint64_t u, u1;
    uint32_t stride;
    // pair: the two CTAs of a cluster walk the same sequence of (token tile, PAIR of column tiles)
    __device__ ItemIter(const UmmaArgs &a, bool pair = false) {
        stream_k = a.stream_k; KB = a.k_blocks; tiles = a.n_tiles * a.m_tiles; first = 1;
        stride = gridDim.x;
        if (pair) {
            tiles = ((a.n_tiles + 1) / 2) * a.m_tiles; stride = gridDim.x / 2;
            cur_tile = blockIdx.x / 2; u = u1 = 0;
        } else if (stream_k) {
            u = a.units * blockIdx.x / gridDim.x;
            u1 = a.units * (blockIdx.x + 1) / gridDim.x;
        } else {
            cur_tile = blockIdx.x; u = u1 = 0;
        }
    }
    __device__ bool next(Item &it) {
        if (!stream_k) {
            if (cur_tile >= tiles) return false;
            it.tile = cur_tile; it.kb0 = 0; it.kb1 = KB; it.slot = -1;
            cur_tile += stride;
            return true;
        }
        if (u >= u1) return false;
        it.tile = (uint32_t)(u / KB);
        it.kb0 = (uint32_t)(u - (uint64_t)it.tile * KB);
        const uint64_t left = u1 - u;
        it.kb1 = (uint64_t)(KB - it.kb0) <= left ? KB : (uint32_t)(it.kb0 + left);
        it.slot = (it.kb0 == 0 && it.kb1 == KB) ? -1 : (int32_t)(blockIdx.x * 2 + (first ? 0 : 1));
        first = 0;
        u += it.kb1 - it.kb0;
        return true;
    }
};

// Ring sizing.  Three resources, two rings:
//   W ring  (kWStages deep): packed weight tiles + their dequant operands in smem.  Filled by the weight
//           producer, consumed by the dequant warps (generic-proxy reads), released by their arrival.
//           It is deep and cheap (4-bit: 5 KB per k-block), so the dequant never waits on HBM latency.
//   X ring  (kXSlots deep): activation tiles in smem;  A ring (kSlots deep): the dequantized A tiles in TMEM, k-block
//           for k-block.  Each is released by its own tcgen05.commit once the stage's MMAs completed.
// Safety rule for the parity waits (verified with an interleaving model of the protocol, scripts/
// pipeline_model.py): the dequant group that owns stage `it` first waits for MMA(it - A) [previous user
// of its TMEM slot], then for the stage's weight bytes.  Those waits cannot alias if the A ring has
// at least as many slots as there are dequant groups and the W ring is at least as deep as the A ring.
template <int CB, int NTOK, int KBS, int NDQ = 4, bool I8 = false, bool PAIR = false>
struct Cfg {
    // one k-block of activations: bf16 = a SW128 tile of NTOK rows x 128 B; int8 = half of one (the stage's two k-blocks
    // share the 128-byte rows of ONE tile)
    // (CTA pair: this CTA's half of the tile's tokens)
    static constexpr int kXBytes = I8 ? NTOK * 64 : (PAIR ? NTOK * 64 : NTOK * 128);
    static constexpr int kACols = I8 ? 16 : 32;                      // TMEM columns of one k-block of A: 64 k x (1 | 2) B / 4
    static constexpr int kWBytes = WL_TILE_N * WL_TILE_K * CB / 8;   // one packed weight tile
    static constexpr int kPBytes = 128 * 8;                          // dequant operands of 128 columns
    static constexpr int kXStage = KBS * kXBytes;
    static constexpr int kWStage = KBS * (kWBytes + kPBytes);
    static constexpr int kSlotCols = KBS * kACols;
    static constexpr int kSlotsRaw = (kTmemCols - kAccStages * NTOK) / kSlotCols;
    static constexpr int kSlotsCap = I8 ? 4 : 8;                     // (int8: A slots are half as wide; keep the W ring the deeper one)
    static constexpr int kSlots = kSlotsRaw > kSlotsCap ? kSlotsCap : kSlotsRaw;     // XA ring depth
    static constexpr int kSmemBudget = 216 * 1024;
    // The activation ring (shared memory, kXSlots) and the A ring (tensor memory, kSlots) have their own barriers and may
    // differ in depth.  Measured (profiles/README.md, round-1 notes): a deeper activation ring (5-8 stages, at the price
    // of W stages) does NOT help — the stage cadence (~900 cycles for 8 MMAs of 128x128x16) is the pace of the tensor
    // pipe itself in this configuration (A from tensor memory + concurrent tcgen05.st of the next A slots), not the
    // activation latency — so both rings stay 4 deep and the W ring keeps the rest of the shared memory.
    static constexpr int kXSlots = kSlots;
    static constexpr int kWStagesRaw = (kSmemBudget - kXSlots * kXStage) / kWStage;
    static constexpr int kWStages = (kWStagesRaw > 24 ? 24 : kWStagesRaw) & ~1;   // W ring depth (even: two producer warps alternate)
    static constexpr int kWOffset = kXSlots * kXStage;
    static constexpr int kBarOffset = kWOffset + kWStages * kWStage;
    static constexpr int kNumBars = 2 * kWStages + 2 * kXSlots + 2 * kSlots + 2 * kAccStages;
    static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16 + 1024;     // + tmem slot + alignment slack
    static_assert(kXStage % 1024 == 0 && kWStage % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte aligned stages");
    static_assert(!I8 || KBS == 2, "int8 mode: one 128-byte-row tile per stage = two k-blocks");
    static_assert(kSlots >= NDQ, "XA ring must have at least as many slots as dequant groups");
    static_assert(kWStages >= kSlots, "W ring must be at least as deep as the A ring");
    static_assert(kTotal <= 227 * 1024, "shared memory over-subscribed");
    static_assert(kAccStages * NTOK + kSlots * kSlotCols <= kTmemCols, "TMEM over-subscribed");
};

// ------------------------------------------------------------------------------------------
// The bf16 kernel of the denoise step (the headline path).  It is kept as its own function, textually apart from the
// generalized body below (int8 mode, CTA pairs, separate activation / A rings): folding those variants into one template
// cost this instance ~4 % of the step rate through different register allocation and scheduling (measured A/B on one box:
// 55.7 vs 53.5 steps/s), although every added branch was `if constexpr`.
// ------------------------------------------------------------------------------------------
template <int CB, int NTOK, int KBS, int NDQ>
__global__ void __launch_bounds__((8 + 4 * NDQ) * 32, 1)
umma_qlinear_kernel(const __grid_constant__ CUtensorMap tmap_x, const UmmaArgs a) {
    using C = Cfg<CB, NTOK, KBS, NDQ>;
    constexpr int kACols = C::kACols;
    constexpr int SW = C::kWStages, A = C::kSlots;
    constexpr int kEpiWarp0 = 4 + 4 * NDQ;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_w = smem + C::kWOffset;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::kBarOffset);
    uint64_t *wfull = bars, *wempty = bars + SW;                     // W ring
    uint64_t *xfull = bars + 2 * SW, *xaempty = xfull + A, *afull = xaempty + A;   // XA ring
    uint64_t *tfull = afull + A, *tempty = tfull + kAccStages;       // accumulators
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + kAccStages);

    // warp index via shuffle: provably warp-uniform, so the role branches below are uniform branches
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // programmatic dependent launch: the next kernel of the stream may be scheduled as SMs free up; its barrier / TMEM set-up
    // and its weight prefetch (which depend on nothing) overlap this kernel's tail.  Everything that reads activations or
    // writes outputs waits for this kernel's predecessors at griddepcontrol.wait.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        for (int s = 0; s < SW; ++s) { mbar_init(wfull + s, 1); mbar_init(wempty + s, 4); }
        for (int s = 0; s < A; ++s) { mbar_init(xfull + s, 1); mbar_init(xaempty + s, 1); mbar_init(afull + s, 4); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_col0 = 0;                          // kAccStages x NTOK accumulator columns
    const uint32_t a_col0 = kAccStages * NTOK;            // then A slots of KBS x 32 columns

    if (warp == 0) {
        // ===================== activation producer (warp-converged; one elected lane issues TMA) =====================
        asm volatile("griddepcontrol.wait;" ::: "memory");     // the activations are the previous kernel's output
        ItemIter iter(a);
        Item item;
        uint32_t it = 0;                               // stage counter of this CTA
        while (iter.next(item)) {
            const uint32_t mt = item.tile / a.n_tiles;
            for (uint32_t kb = item.kb0; kb < item.kb1; kb += KBS, ++it) {
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                const uint32_t s = it % A, ph = (it / A) & 1;
                mbar_wait(xaempty + s, ph ^ 1);
                if (elect_one()) {
                    TRACE(0, it);
                    uint8_t *stage = smem + s * C::kXStage;
                    if (!(a.dbg & 8)) {
                        if (a.x3d) {
                            // one TMA for all KBS k-blocks: box {64 k, NTOK tokens, KBS k-blocks}; k-blocks past K are zero-filled
                            mbar_arrive_expect_tx(xfull + s, KBS * C::kXBytes);
                            tma_load_3d(stage, &tmap_x, xfull + s, 0, (int)(mt * NTOK), (int)kb);
                        } else {
                            mbar_arrive_expect_tx(xfull + s, nk * C::kXBytes);
                            for (uint32_t sub = 0; sub < nk; ++sub)
                                tma_load_2d(stage + sub * C::kXBytes, &tmap_x, xfull + s, (int)((kb + sub) * WL_TILE_K), (int)(mt * NTOK));
                        }
                    } else {
                        mbar_arrive(xfull + s);        // timing experiment: no activation traffic
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===================== weight producers (two warps alternate stages; they run ahead through the deep W ring) =====================
        const uint32_t me = (uint32_t)(warp - 2);
        ItemIter iter(a);
        Item item;
        uint32_t it = 0;
        while (iter.next(item)) {
            const uint32_t nt = item.tile % a.n_tiles;
            const uint8_t *wsrc = a.packed + ((size_t)nt * a.k_blocks) * C::kWBytes;
            const uint2 *psrc = a.dqparams + (size_t)nt * 128;
            for (uint32_t kb = item.kb0; kb < item.kb1; kb += KBS, ++it) {
                if ((it & 1) != me) continue;
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                const uint32_t s = it % SW, ph = (it / SW) & 1;
                mbar_wait(wempty + s, ph ^ 1);
                if (elect_one()) {
                    TRACE(2, it);                      // weight producer: W stage free, issuing loads
                    uint8_t *stage = smem_w + s * C::kWStage;
                    mbar_arrive_expect_tx(wfull + s, nk * (C::kWBytes + C::kPBytes));
                    bulk_load(stage, wsrc + (size_t)kb * C::kWBytes, nk * C::kWBytes, wfull + s);
                    for (uint32_t sub = 0; sub < nk; ++sub) {
                        const uint32_t g = (kb + sub) / a.group_kb;
                        bulk_load(stage + KBS * C::kWBytes + sub * C::kPBytes, psrc + (size_t)g * a.Npad, C::kPBytes, wfull + s);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-converged; one elected lane issues) =====================
        constexpr uint32_t idesc = make_idesc(NTOK);
        ItemIter iter(a);
        Item item;
        uint32_t it = 0, n_item = 0;
        while (iter.next(item)) {
            const uint32_t acc = n_item % kAccStages, aph = (n_item / kAccStages) & 1;
            ++n_item;
            mbar_wait(tempty + acc, aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc_col0 + acc * NTOK;
            // Software-pipelined: the barriers of stage it+1 are polled after the first k-block of stage `it`
            // has been issued, so the tensor pipe never drains while this warp sits in a try_wait.
            bool ready = false;
            for (uint32_t kb = item.kb0; kb < item.kb1; kb += KBS, ++it) {
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                const uint32_t s = it % A, ph = (it / A) & 1;
                if (!ready) {
                    if (lane == 0) TRACE(1, it);
                    mbar_wait(xfull + s, ph);      // activation tiles landed
                    mbar_wait(afull + s, ph);      // A slot written to TMEM
                    tc_fence_after();
                }
                if (lane == 0) TRACE(3, it);
                const uint32_t stage_addr = smem_u32(smem + s * C::kXStage);
                const uint32_t a_tmem = tmem_base + a_col0 + s * C::kSlotCols;
                const bool first = kb == item.kb0;
                if (elect_one() && !(a.dbg & 1)) {             // k-block 0 of the stage
                    const uint64_t bdesc = make_b_desc(stage_addr);
#pragma unroll
                    for (int k4 = 0; k4 < WL_TILE_K / 16; ++k4)
                        umma_ts(d_tmem, a_tmem + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, (!first || k4 > 0) ? 1u : 0u);
                }
                __syncwarp();
                ready = false;
                if (kb + KBS < item.kb1) {                     // peek at the next stage while those MMAs run
                    const uint32_t s2 = (it + 1) % A, ph2 = ((it + 1) / A) & 1;
                    mbar_wait(xfull + s2, ph2);
                    mbar_wait(afull + s2, ph2);
                    tc_fence_after();
                    ready = true;
                }
                if (elect_one()) {
                    for (uint32_t sub = 1; sub < nk && !(a.dbg & 1); ++sub) {
                        const uint64_t bdesc = make_b_desc(stage_addr + sub * C::kXBytes);
#pragma unroll
                        for (int k4 = 0; k4 < WL_TILE_K / 16; ++k4)
                            umma_ts(d_tmem, a_tmem + sub * kACols + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, 1u);
                    }
                    umma_commit(xaempty + s);      // frees the activation stage and the TMEM A slot
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(tfull + acc);
            __syncwarp();
        }
    } else if (warp >= 4 && warp < kEpiWarp0) {
        // ===================== dequant warps =====================
        const uint32_t grp = (uint32_t)(warp - 4) >> 2;
        const int quarter = warp & 3;             // TMEM lanes [32*quarter, +32)
        const int n_local = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a_col0;
        ItemIter iter(a);
        Item item;
        uint32_t it = 0;
        while (iter.next(item)) {
            for (uint32_t kb = item.kb0; kb < item.kb1; kb += KBS, ++it) {
                if (it % NDQ != grp) continue;    // group g owns the stages with it % NDQ == g
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                const uint32_t sw = it % SW, wph = (it / SW) & 1;
                const uint32_t sl = it % A, aph = (it / A) & 1;
                // 1. previous user of the TMEM slot, MMA(it - A), completed; 2. this stage's weights landed
                if (it >= (uint32_t)A) mbar_wait(xaempty + sl, aph ^ 1);
                mbar_wait(wfull + sw, wph);
                tc_fence_after();
                if (quarter == 0 && lane == 0) TRACE(4, it);
                const uint8_t *stage = smem_w + sw * C::kWStage;
                for (uint32_t sub = 0; sub < nk; ++sub) {
                    const uint4 *wpk = reinterpret_cast<const uint4 *>(stage + sub * C::kWBytes);
                    const uint2 prm = lds64(smem_u32(stage + KBS * C::kWBytes + sub * C::kPBytes) + (uint32_t)n_local * 8u);
                    uint32_t vals[32];
                    if (!(a.dbg & 2)) dequant_kblock<CB>(wpk, n_local, prm.x, prm.y, vals);
                    else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) vals[q] = prm.x + q;
                    }
                    if (!(a.dbg & 4)) tmem_st32(lane_addr + sl * C::kSlotCols + sub * kACols, vals);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(wempty + sw);   // all smem reads of this warp are done (values are in registers)
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(afull + sl);
                if (quarter == 0 && lane == 0) TRACE(5, it);
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue warps =====================
        asm volatile("griddepcontrol.wait;" ::: "memory");     // outputs / partial tiles may still be read by the previous kernels
        const int quarter = warp & 3;
        ItemIter iter(a);
        Item item;
        uint32_t n_item = 0;
        while (iter.next(item)) {
            const uint32_t nt = item.tile % a.n_tiles, mt = item.tile / a.n_tiles;
            const uint32_t acc = n_item % kAccStages, aph = (n_item / kAccStages) & 1;
            ++n_item;
            const uint32_t n = nt * 128 + quarter * 32 + lane;
            const bool n_ok = n < a.N;
            const bool direct = item.slot < 0;
            const float bias = (direct && a.bias != nullptr && n_ok) ? __ldg(a.bias + n) : 0.f;
            float *part = direct ? nullptr : a.partial + (size_t)item.slot * (NTOK * 128) + quarter * 32 + lane;
            mbar_wait(tfull + acc, aph);
            tc_fence_after();
            if (quarter == 0 && lane == 0) TRACE(6, n_item - 1);   // epilogue: accumulator ready (indexed by item)
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc_col0 + acc * NTOK;
            // wide TMEM loads (few round trips: tcgen05.ld competes with the MMA's accumulator traffic), and the
            // accumulator is handed back to the MMA warp as soon as its last column is in registers
            constexpr int CH = NTOK >= 64 ? 64 : NTOK;
#pragma unroll 1
            for (int c0 = 0; c0 < NTOK; c0 += CH) {
                uint32_t v[CH];
                tmem_ld_chunk<CH>(t_acc + c0, v);
                tmem_ld_wait();
                if (c0 + CH >= NTOK) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty + acc);
                }
                const uint32_t m_base = mt * NTOK + c0;
                if (direct) {
                    if (!n_ok || (a.dbg & 64)) continue;
                    if (a.y_f32) {
                        float *yp = a.y_f32 + (size_t)m_base * a.N + n;
#pragma unroll
                        for (int j = 0; j < CH; ++j)
                            if (m_base + j < a.M) yp[(size_t)j * a.N] = __uint_as_float(v[j]) + bias;
                    }
                    if (a.y_bf16) {
                        __nv_bfloat16 *yp = a.y_bf16 + (size_t)m_base * a.N + n;
#pragma unroll
                        for (int j = 0; j < CH; ++j)
                            if (m_base + j < a.M) yp[(size_t)j * a.N] = __float2bfloat16_rn(__uint_as_float(v[j]) + bias);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CH; ++j) part[(size_t)(c0 + j) * 128] = __uint_as_float(v[j]);
                }
            }
            if (quarter == 0 && lane == 0) TRACE(7, n_item - 1);   // epilogue: accumulator released
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}


template <int CB, int NTOK, int KBS, int NDQ, bool I8, bool PAIR>
__device__ __forceinline__ void umma_qlinear_body(const CUtensorMap &tmap_x, const UmmaArgs &a) {
    using C = Cfg<CB, NTOK, KBS, NDQ, I8, PAIR>;
    constexpr int kACols = C::kACols;
    static_assert(!PAIR || !I8, "the CTA-pair variant exists for the bf16 mode only");
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;                 // 0 = leader: issues the MMAs
    const uint32_t n_pairs = (a.n_tiles + 1) / 2;
    // column / token tile of an item.  Pair mode: this CTA's column tile is 2 * pair + rank; an odd tile count leaves the
    // last pair's second CTA without columns: it loads the last real tile again and its epilogue stores nothing
    // Tile order: column-tile-minor (CTAs running together share a token tile's activations).  DLLM_UMMA_DBG bit 32 selects a
    // rasterized order instead (groups of kRasterM token tiles, column-tile-major inside a group: the CTAs spread over
    // ~16 token tiles x ~10 column tiles) — measured: no difference, so L2 hot-spotting is not what makes an activation
    // stage take ~1.3 us from TMA issue to arrival.
    constexpr uint32_t kRasterM = 16;
    const uint32_t n_cols = PAIR ? n_pairs : a.n_tiles;                   // column tiles (or pairs of them) per token tile
    auto decode = [&](uint32_t tile, uint32_t &mt, uint32_t &nc) {
        if (a.stream_k || !(a.dbg & 32)) { mt = tile / n_cols; nc = tile % n_cols; return; }
        const uint32_t grp = tile / (kRasterM * n_cols), r = tile - grp * (kRasterM * n_cols);
        const uint32_t left = a.m_tiles - grp * kRasterM, gsz = left < kRasterM ? left : kRasterM;
        nc = r / gsz; mt = grp * kRasterM + (r - nc * gsz);
    };
    auto tile_mt = [&](const Item &it) -> uint32_t { uint32_t mt, nc; decode(it.tile, mt, nc); return mt; };
    auto tile_nt = [&](const Item &it) -> uint32_t { uint32_t mt, nc; decode(it.tile, mt, nc); return PAIR ? 2 * nc + rank : nc; };
    constexpr int SW = C::kWStages, A = C::kSlots, X = C::kXSlots;
    constexpr int kEpiWarp0 = 4 + 4 * NDQ;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_w = smem + C::kWOffset;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::kBarOffset);
    uint64_t *wfull = bars, *wempty = bars + SW;                     // W ring
    uint64_t *xfull = bars + 2 * SW, *xempty = xfull + X;             // activation ring (shared memory)
    // A ring (tensor memory).  While both rings have the same depth one barrier (and one tcgen05.commit per stage) frees the
    // activation stage and the A slot together: a second commit per stage costs ~3 % of the step rate (measured)
    uint64_t *afull = xempty + X, *aempty = (X == A) ? xempty : afull + A;
    uint64_t *tfull = afull + 2 * A, *tempty = tfull + kAccStages;   // accumulators (after the A ring's own `empty` barriers, used or not)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + kAccStages);

    // warp index via shuffle: provably warp-uniform, so the role branches below are uniform branches
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        // pair mode: afull / tempty of the LEADER collect the dequant / epilogue warps of both CTAs
        for (int s = 0; s < SW; ++s) { mbar_init(wfull + s, 1); mbar_init(wempty + s, 4); }
        for (int s = 0; s < X; ++s) { mbar_init(xfull + s, 1); mbar_init(xempty + s, 1); }
        for (int s = 0; s < A; ++s) { mbar_init(afull + s, PAIR ? 8 : 4); if (X != A) mbar_init(aempty + s, 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, PAIR ? 8 : 4); }
        fence_barrier_init();
    }
    if constexpr (PAIR) {
        cluster_sync_all();                                // barriers of both CTAs exist before anybody signals them
        if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
    } else {
        if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_col0 = 0;                          // kAccStages x NTOK accumulator columns
    const uint32_t a_col0 = kAccStages * NTOK;            // then A slots of KBS x 32 columns

    if (warp == 0) {
        // ===================== activation producer (warp-converged; one elected lane issues TMA) =====================
        ItemIter iter(a, PAIR);
        Item item;
        uint32_t it = 0;                               // stage counter of this CTA
        const uint32_t xfull_leader = PAIR ? leader_addr(xfull) : 0u;
        while (iter.next(item)) {
            const uint32_t mt = tile_mt(item);
            for (uint32_t kb = item.kb0; kb < item.kb1; kb += KBS, ++it) {
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                const uint32_t s = it % X, ph = (it / X) & 1;
                mbar_wait(xempty + s, ph ^ 1);
                if (elect_one()) {
                    TRACE(0, it);
                    uint8_t *stage = smem + s * C::kXStage;
                    if constexpr (PAIR) {
                        // this CTA's half of the tokens; the leader's barrier expects the bytes of both halves
                        if (rank == 0) mbar_arrive_expect_tx(xfull + s, 2 * KBS * C::kXBytes);
                        tma_load_3d_pair(stage, &tmap_x, xfull_leader + s * 8, 0, (int)(mt * NTOK + rank * (NTOK / 2)), (int)kb);
                    } else if constexpr (I8) {
                        // int8: one box {128 k, NTOK tokens} = both k-blocks of the stage; k past K is zero-filled
                        mbar_arrive_expect_tx(xfull + s, KBS * C::kXBytes);
                        tma_load_2d(stage, &tmap_x, xfull + s, (int)(kb * WL_TILE_K), (int)(mt * NTOK));
                    } else if (!(a.dbg & 8)) {
                        if (a.x3d) {
                            // one TMA for all KBS k-blocks: box {64 k, NTOK tokens, KBS k-blocks}; k-blocks past K are zero-filled
                            mbar_arrive_expect_tx(xfull + s, KBS * C::kXBytes);
                            tma_load_3d(stage, &tmap_x, xfull + s, 0, (int)(mt * NTOK), (int)kb);
                        } else {
                            mbar_arrive_expect_tx(xfull + s, nk * C::kXBytes);
                            for (uint32_t sub = 0; sub < nk; ++sub)
                                tma_load_2d(stage + sub * C::kXBytes, &tmap_x, xfull + s, (int)((kb + sub) * WL_TILE_K), (int)(mt * NTOK));
                        }
                    } else {
                        mbar_arrive(xfull + s);        // timing experiment: no activation traffic
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===================== weight producers (two warps alternate stages; they run ahead through the deep W ring) =====================
        const uint32_t me = (uint32_t)(warp - 2);
        ItemIter iter(a, PAIR);
        Item item;
        uint32_t it = 0;
        while (iter.next(item)) {
            const uint32_t nt = tile_nt(item) < a.n_tiles ? tile_nt(item) : a.n_tiles - 1;
            const uint8_t *wsrc = a.packed + ((size_t)nt * a.k_blocks) * C::kWBytes;
            const uint2 *psrc = a.dqparams + (size_t)nt * 128;
            for (uint32_t kb = item.kb0; kb < item.kb1; kb += KBS, ++it) {
                if ((it & 1) != me) continue;
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                const uint32_t s = it % SW, ph = (it / SW) & 1;
                mbar_wait(wempty + s, ph ^ 1);
                if (elect_one()) {
                    TRACE(2, it);                      // weight producer: W stage free, issuing loads
                    uint8_t *stage = smem_w + s * C::kWStage;
                    mbar_arrive_expect_tx(wfull + s, nk * (C::kWBytes + C::kPBytes));
                    bulk_load(stage, wsrc + (size_t)kb * C::kWBytes, nk * C::kWBytes, wfull + s);
                    for (uint32_t sub = 0; sub < nk; ++sub) {
                        const uint32_t g = (kb + sub) / a.group_kb;
                        bulk_load(stage + KBS * C::kWBytes + sub * C::kPBytes, psrc + (size_t)g * a.Npad, C::kPBytes, wfull + s);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (warp-converged; one elected lane issues; pair mode: the leader CTA only) =====================
        constexpr uint32_t idesc = I8 ? make_idesc_i8(NTOK) : (PAIR ? make_idesc_pair(NTOK) : make_idesc(NTOK));
        ItemIter iter(a, PAIR);
        Item item;
        uint32_t it = 0, n_item = 0;
        while (iter.next(item)) {
            const uint32_t acc = n_item % kAccStages, aph = (n_item / kAccStages) & 1;
            ++n_item;
            mbar_wait(tempty + acc, aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc_col0 + acc * NTOK;
            // Software-pipelined: the barriers of stage it+1 are polled after the first k-block of stage `it`
            // has been issued, so the tensor pipe never drains while this warp sits in a try_wait.
            bool ready = false;
            for (uint32_t kb = item.kb0; kb < item.kb1; kb += KBS, ++it) {
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                const uint32_t s = it % X, ph = (it / X) & 1;          // activation stage
                const uint32_t sa = it % A, pha = (it / A) & 1;        // A slot
                if (!ready) {
                    if (lane == 0) TRACE(1, it);
                    mbar_wait(xfull + s, ph);      // activation tiles landed
                    mbar_wait(afull + sa, pha);    // A slot written to TMEM
                    tc_fence_after();
                }
                if (lane == 0) TRACE(3, it);
                const uint32_t stage_addr = smem_u32(smem + s * C::kXStage);
                const uint32_t a_tmem = tmem_base + a_col0 + sa * C::kSlotCols;
                const bool first = kb == item.kb0;
                if (elect_one() && !(a.dbg & 1)) {             // k-block 0 of the stage
                    const uint64_t bdesc = make_b_desc(stage_addr);
                    if constexpr (I8) {
                        // 32 k (= 32 bytes of the 128-byte row, 8 TMEM columns of A) per MMA: two per k-block
#pragma unroll
                        for (int k2 = 0; k2 < 2; ++k2)
                            umma_ts_i8(d_tmem, a_tmem + k2 * 8, bdesc + (uint64_t)(k2 * 2), idesc, (!first || k2 > 0) ? 1u : 0u);
                    } else if constexpr (PAIR) {
#pragma unroll
                        for (int k4 = 0; k4 < WL_TILE_K / 16; ++k4)
                            umma_ts_pair(d_tmem, a_tmem + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, (!first || k4 > 0) ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int k4 = 0; k4 < WL_TILE_K / 16; ++k4)
                            umma_ts(d_tmem, a_tmem + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, (!first || k4 > 0) ? 1u : 0u);
                    }
                }
                __syncwarp();
                ready = false;
                if (kb + KBS < item.kb1 && !(a.dbg & 16)) {    // peek at the next stage while those MMAs run
                    const uint32_t s2 = (it + 1) % X, ph2 = ((it + 1) / X) & 1;
                    const uint32_t sa2 = (it + 1) % A, pha2 = ((it + 1) / A) & 1;
                    mbar_wait(xfull + s2, ph2);
                    mbar_wait(afull + sa2, pha2);
                    tc_fence_after();
                    ready = true;
                }
                if (elect_one()) {
                    for (uint32_t sub = 1; sub < nk && !(a.dbg & 1); ++sub) {
                        if constexpr (I8) {
                            const uint64_t bdesc = make_b_desc(stage_addr);      // second half of the same 128-byte rows
#pragma unroll
                            for (int k2 = 0; k2 < 2; ++k2)
                                umma_ts_i8(d_tmem, a_tmem + sub * kACols + k2 * 8, bdesc + (uint64_t)((sub * 2 + k2) * 2), idesc, 1u);
                        } else {
                            const uint64_t bdesc = make_b_desc(stage_addr + sub * C::kXBytes);
#pragma unroll
                            for (int k4 = 0; k4 < WL_TILE_K / 16; ++k4) {
                                if constexpr (PAIR) umma_ts_pair(d_tmem, a_tmem + sub * kACols + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, 1u);
                                else umma_ts(d_tmem, a_tmem + sub * kACols + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, 1u);
                            }
                        }
                    }
                    // frees the activation stage and the TMEM A slot (pair mode: in both CTAs)
                    if constexpr (PAIR) { umma_commit_pair(xempty + s); if constexpr (X != A) umma_commit_pair(aempty + sa); }
                    else { umma_commit(xempty + s); if constexpr (X != A) umma_commit(aempty + sa); }
                }
                __syncwarp();
            }
            if (elect_one()) { if constexpr (PAIR) umma_commit_pair(tfull + acc); else umma_commit(tfull + acc); }
            __syncwarp();
        }
    } else if (warp >= 4 && warp < kEpiWarp0) {
        // ===================== dequant warps =====================
        const uint32_t grp = (uint32_t)(warp - 4) >> 2;
        const int quarter = warp & 3;             // TMEM lanes [32*quarter, +32)
        const int n_local = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a_col0;
        ItemIter iter(a, PAIR);
        Item item;
        uint32_t it = 0;
        const uint32_t afull_leader = PAIR ? leader_addr(afull) : 0u;
        while (iter.next(item)) {
            for (uint32_t kb = item.kb0; kb < item.kb1; kb += KBS, ++it) {
                if (it % NDQ != grp) continue;    // group g owns the stages with it % NDQ == g
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                const uint32_t sw = it % SW, wph = (it / SW) & 1;
                const uint32_t sl = it % A, aph = (it / A) & 1;
                // 1. previous user of the TMEM slot, MMA(it - A), completed; 2. this stage's weights landed
                if (it >= (uint32_t)A) mbar_wait(aempty + sl, aph ^ 1);
                mbar_wait(wfull + sw, wph);
                tc_fence_after();
                if (quarter == 0 && lane == 0) TRACE(4, it);
                const uint8_t *stage = smem_w + sw * C::kWStage;
                for (uint32_t sub = 0; sub < nk; ++sub) {
                    const uint4 *wpk = reinterpret_cast<const uint4 *>(stage + sub * C::kWBytes);
                    const uint2 prm = lds64(smem_u32(stage + KBS * C::kWBytes + sub * C::kPBytes) + (uint32_t)n_local * 8u);
                    uint32_t vals[32];
                    if constexpr (I8) {
                        unpack_kblock_u8<CB>(wpk, n_local, vals);
                        tmem_st16(lane_addr + sl * C::kSlotCols + sub * kACols, vals);
                        continue;
                    }
                    if (!(a.dbg & 2)) dequant_kblock<CB>(wpk, n_local, prm.x, prm.y, vals);
                    else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) vals[q] = prm.x + q;
                    }
                    if (!(a.dbg & 4)) tmem_st32(lane_addr + sl * C::kSlotCols + sub * kACols, vals);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(wempty + sw);   // all smem reads of this warp are done (values are in registers)
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster(afull_leader + sl * 8); else mbar_arrive(afull + sl); }
                if (quarter == 0 && lane == 0) TRACE(5, it);
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;
        ItemIter iter(a, PAIR);
        Item item;
        uint32_t n_item = 0;
        const uint32_t tempty_leader = PAIR ? leader_addr(tempty) : 0u;
        while (iter.next(item)) {
            const uint32_t nt = tile_nt(item), mt = tile_mt(item);
            const uint32_t acc = n_item % kAccStages, aph = (n_item / kAccStages) & 1;
            ++n_item;
            const uint32_t n = nt * 128 + quarter * 32 + lane;
            const bool n_ok = n < a.N;
            const bool direct = item.slot < 0;
            const float bias = (direct && a.bias != nullptr && n_ok) ? __ldg(a.bias + n) : 0.f;
            float *part = direct ? nullptr : a.partial + (size_t)item.slot * (NTOK * 128) + quarter * 32 + lane;
            mbar_wait(tfull + acc, aph);
            tc_fence_after();
            if (quarter == 0 && lane == 0) TRACE(6, n_item - 1);   // epilogue: accumulator ready (indexed by item)
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc_col0 + acc * NTOK;
            // wide TMEM loads (few round trips: tcgen05.ld competes with the MMA's accumulator traffic), and the
            // accumulator is handed back to the MMA warp as soon as its last column is in registers
            constexpr int CH = NTOK >= 64 ? 64 : NTOK;
#pragma unroll 1
            for (int c0 = 0; c0 < NTOK; c0 += CH) {
                uint32_t v[CH];
                tmem_ld_chunk<CH>(t_acc + c0, v);
                tmem_ld_wait();
                if (c0 + CH >= NTOK) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster(tempty_leader + acc * 8); else mbar_arrive(tempty + acc); }
                }
                const uint32_t m_base = mt * NTOK + c0;
                if constexpr (I8) {
                    // exact: sum x q - zp sum x  (dequantize_tensor's `- zp`, quantization.rs:83, in the integer domain)
                    if (!n_ok) continue;
                    const int32_t *sx = reinterpret_cast<const int32_t *>(a.partial);
                    const int32_t zp = (int32_t)a.x3d;
                    if (a.dbg & 256) {
                        // int8 denoise mode: the exact integer sum leaves as  (sum - zp rowsum) * (weight scale * token scale) + bias
                        // (dequantize_tensor's `(q - zp) * scale`, quantization.rs:83, composed with the token's own step);
                        // a.trace = [M] f32 weight scale x token scale, y_f32 / y_bf16 = float outputs
                        const float *rs = reinterpret_cast<const float *>(a.trace);
                        const float bn = (a.bias != nullptr) ? __ldg(a.bias + n) : 0.f;
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            if (m_base + j < a.M) {
                                const int32_t e = (int32_t)v[j] - zp * __ldg(sx + m_base + j);
                                const float f = fmaf((float)e, __ldg(rs + m_base + j), bn);
                                if (a.y_f32) a.y_f32[(size_t)(m_base + j) * a.N + n] = f;
                                if (a.y_bf16) a.y_bf16[(size_t)(m_base + j) * a.N + n] = __float2bfloat16_rn(f);
                            }
                        }
                        continue;
                    }
                    int32_t *yp = reinterpret_cast<int32_t *>(a.y_f32) + (size_t)m_base * a.N + n;
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        if (m_base + j < a.M) yp[(size_t)j * a.N] = (int32_t)v[j] - zp * __ldg(sx + m_base + j);
                } else if (direct) {
                    if (!n_ok || (a.dbg & 64)) continue;
                    if (a.y_f32) {
                        float *yp = a.y_f32 + (size_t)m_base * a.N + n;
#pragma unroll
                        for (int j = 0; j < CH; ++j)
                            if (m_base + j < a.M) yp[(size_t)j * a.N] = __uint_as_float(v[j]) + bias;
                    }
                    if (a.y_bf16) {
                        __nv_bfloat16 *yp = a.y_bf16 + (size_t)m_base * a.N + n;
#pragma unroll
                        for (int j = 0; j < CH; ++j)
                            if (m_base + j < a.M) yp[(size_t)j * a.N] = __float2bfloat16_rn(__uint_as_float(v[j]) + bias);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CH; ++j) part[(size_t)(c0 + j) * 128] = __uint_as_float(v[j]);
                }
            }
            if (quarter == 0 && lane == 0) TRACE(7, n_item - 1);   // epilogue: accumulator released
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) {
        cluster_sync_all();                                // the partner may still signal barriers / read shared memory of this CTA
        if (warp == 1) {
            tc_fence_after();
            tmem_dealloc_pair(tmem_base, kTmemCols);
        }
    } else {
        if (warp == 1) {
            tc_fence_after();
            tmem_dealloc(tmem_base, kTmemCols);
        }
    }
}

// the generalized body as a 1-CTA kernel (used by the int8 mode)
template <int CB, int NTOK, int KBS, int NDQ, bool I8>
__global__ void __launch_bounds__((8 + 4 * NDQ) * 32, 1)
umma_qlinear_x_kernel(const __grid_constant__ CUtensorMap tmap_x, const UmmaArgs a) {
    umma_qlinear_body<CB, NTOK, KBS, NDQ, I8, false>(tmap_x, a);
}

// the CTA-pair variant (bf16 mode, whole tiles): a cluster of two CTAs per (token tile, pair of column tiles)
template <int CB, int NTOK, int KBS, int NDQ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((8 + 4 * NDQ) * 32, 1)
umma_qlinear_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const UmmaArgs a) {
    umma_qlinear_body<CB, NTOK, KBS, NDQ, false, true>(tmap_x, a);
}


// ------------------------------------------------------------------------------------------
// CTA-pair kernel with 256-token tiles ("pair2"): the dense kernel of the denoise step.
//
// What bounded the 1-CTA kernel: per 128 x 128 x 128 stage it moves 32 KB of activations and 10 KB of
// weights from L2 into the SM, dequantizes 16 K weights and pays one barrier hand-off chain, all for 512 cycles of MMA.
// Here a CLUSTER OF TWO CTAs owns 256 output columns x `ntok` (<= 256) tokens: each CTA dequantizes its own 128 columns
// into its own tensor memory ONCE per stage and the leader issues tcgen05.mma.cta_group::2 (M = 256) against TWO token
// halves, so a stage carries up to 1024 cycles of MMA per SM for the same dequant work, the same weight bytes and the same
// hand-offs — and each CTA loads only HALF of the activation tile (the B operand is read from both CTAs' shared memory).
// Per flop: activation bytes / 2, weight bytes / 2, dequant / 2, hand-offs / 2.
//
// Tensor memory (512 columns per CTA): two 128-column accumulators (token halves h = 0, 1; each CTA holds its 128 rows)
// + 4 A slots of 64 columns.  The accumulators are single-buffered, but the halves are committed / drained / re-armed
// separately: while the epilogue drains half 0 the tensor pipe still runs the last stage's half-1 MMAs, and the next
// tile's first half-0 MMAs run while half 1 is drained — the drain hides behind 512 cycles of MMAs on either side.
// Token <-> accumulator column: CTA r loads tokens [r*ntok/2, (r+1)*ntok/2) of the tile as ONE box; half h of the tile is
// rows [h*q, (h+1)*q) of BOTH CTAs' boxes (q = ntok/4), i.e. column c of half h is token r*ntok/2 + h*q + i with
// r = c / q, i = c % q.  ntok is a run-time multiple of 32 (MMA N = ntok/2 is a multiple of 16), chosen per problem so
// that the tile count fills the 74 CTA pairs evenly (e.g. 8192 tokens x 2048 columns: 224-token tiles = 4 full waves).
// ------------------------------------------------------------------------------------------
struct RsMaps { CUtensorMap m[8]; };       // receive buffers [M, N] bf16 of up to 8 ranks (peer-mapped)
struct Pair2Args {
    const uint8_t *packed;
    const uint2 *dqparams;
    const float *bias;
    float *y_f32;
    __nv_bfloat16 *y_bf16;
    uint32_t M, N, Npad, k_blocks, n_tiles, n_pairs, m_tiles, group_kb;
    uint32_t ntok;             // tokens per pair tile: multiple of 32, <= 256
    uint32_t tiles;            // n_pairs * m_tiles
    // fused reduce-scatter (tensor-parallel row layers, tp.cu): rs_world > 0 -> the bf16 tile rows are not stored into y but pushed
    // over NVLink into the RECEIVE buffer of the rank that owns those tokens (rank o owns tokens [o*rs_rows, (o+1)*rs_rows)), row
    // block `rs_rank` of it: one bulk tensor store per q-row block, straight from the staging buffer into peer memory
    uint32_t rs_world, rs_rank, rs_rows;
    // token-tile rotation: tile index t covers token tile (t / n_pairs + mt_rot) % m_tiles.  Under tensor parallelism every rank
    // starts at its OWN token slice: the fused reduce-scatter then pushes to a different owner from every rank at any time
    // (no incast on rank 0's links), and a gated consumer works on the rows it already holds while the others' arrive
    uint32_t mt_rot;
    // gated activation loads (the all-gather half of the exchange runs UNDER this GEMM): before the first load of a token tile of
    // slice s != gate_self the loader waits until gate[s] — a counter in this rank's arena that rank s's reduce / gather kernel
    // bumps once per block after its rows have landed here — has reached gate_target
    const uint32_t *gate;             // [ranks][8] counters: [s][g] covers tokens [s*gate_rows + g*gate_sub, + gate_sub)
    uint32_t gate_target, gate_rows, gate_sub, gate_self;
    unsigned int *gate_err;
    uint32_t dbg;              // timing experiments only (DLLM_UMMA_DBG): 1 skip MMAs, 2 skip dequant math, 8 skip activation loads, 64 skip stores
    long long *trace;          // dbg & 128: clock64 stamps of cluster 0's leader CTA: [role 0..7][256]
    // int8 variant (the int8 denoise mode): per token the int8 row sum and weight scale x activation step, and the weight's zero-point
    const int32_t *i8_rowsum;
    const float *i8_rowscale;
    int32_t i8_zp;
};
#define TRACE2(role, idx) do { if ((a.dbg & 128) && blockIdx.x == 0 && (idx) < 256) a.trace[(role) * 256 + (idx)] = clock64(); } while (0)

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}

template <int CB, bool I8 = false>
struct CfgP2 {
    static constexpr int KBS = 2;
    static constexpr int kXSlabMax = 128 * 128;                      // one k-block of this CTA's half of the tokens (<= 128 rows x 128 B)
    // int8: one 128-byte row holds 128 k = BOTH k-blocks of the stage; a k-block of A is 16 tensor-memory columns (64 bytes per row)
    static constexpr int kXStage = I8 ? kXSlabMax : KBS * kXSlabMax;
    static constexpr int kACols = I8 ? 16 : 32, kSlotCols = KBS * kACols;
    // accumulator buffers of 128 columns (one token half each): two — the halves of a tile.  int8: the A ring is half as wide, which
    // leaves room for a THIRD buffer (-DDLLM_I8_NBUF=3 -DDLLM_I8_PRE=2: tile t, half h lives in buffer (2 t + h) % 3, so the next tile's half 0 is
    // accumulated while this tile's halves are drained) — but only beside a 4-slot A ring, and then the dequant warps' latency
    // (1.5 K cycles per stage) bounds the stage period at 975 cycles instead of 750: measured slower (85.2 against 87.5 steps/s), so
    // two buffers and six slots stay.
#ifndef DLLM_I8_NBUF
#define DLLM_I8_NBUF 2
#endif
    // int8, A ring in SHARED memory (-DDLLM_I8_ASMEM=1): the unpacked codes of a stage are 16 KB per CTA — a SWIZZLE_128B tile like the
    // activations', written by the unpack warps with plain shared-memory stores and read by the MMA through a descriptor — which
    // leaves all 512 tensor-memory columns to FOUR accumulator buffers: the halves of tile t + 1 are accumulated while those of tile t
    // are drained.  Measured: the tile-boundary stall disappears, but shared memory then only holds 4 stages of X + A, and with the
    // unpack latency at 2 K cycles per stage (the epilogue warps now run all the time on the same sub-cores) the stage period is
    // 1000 cycles instead of 750: 82.2 against 88.7 steps/s — so the tensor-memory ring stays the default.  Both pass the parity tests.
#ifndef DLLM_I8_ASMEM
#define DLLM_I8_ASMEM 0
#endif
    static constexpr bool kASmem = I8 && DLLM_I8_ASMEM;
    static constexpr int kAccBufs = kASmem ? 4 : (I8 ? DLLM_I8_NBUF : 2);
    // X ring (shared memory) and A ring (tensor memory / shared memory): same depth, one commit frees both
    // (int8: SEVEN slots — eight would fill tensor memory, but shared memory then only holds 7 weight stages beside them.  The empty-pipeline timeline showed a slot's round
    //  trip — commit -> both CTAs' unpack warps -> tensor-memory store -> remote arrive -> MMA — at 3.2 K cycles before any work, so
    //  the stage period is (3.2 K + work) / slots: 774 cycles with six slots where the MMAs need 512)
#ifndef DLLM_I8_SLOTS
#define DLLM_I8_SLOTS 7
#endif
    static constexpr int kSlots = kASmem ? 4 : (I8 ? (kAccBufs == 3 ? 4 : DLLM_I8_SLOTS) : 4);
    static constexpr int kAStage = kASmem ? 128 * 128 : 0;             // 128 columns x 128 k bytes
    // how many stages half 0 is issued ahead of half 1 (1 = lock step)
#ifndef DLLM_I8_PRE
#define DLLM_I8_PRE 3
#endif
    static constexpr int kPre = (I8 && !kASmem) ? DLLM_I8_PRE : 1;
    // epilogue warps per CTA: 8 (two per tensor-memory lane quarter, q <= 64 columns each per half).  -DDLLM_I8_EPI16=1: 16 for int8,
    // four per lane quarter with q / 2 <= 32 columns each, i.e. one tensor-memory round trip and 16 output pairs per warp and half
    // instead of two and 32 — measured slower: 1024 threads leave 64 registers, half 0 is handed back after 1.9 K cycles instead
    // of 2.6 K but its conversion and staging then take 4.9 K (8.4 K per tile boundary instead of 6.1 K; 81.9 against 88.4 steps/s).
#ifndef DLLM_I8_EPI16
#define DLLM_I8_EPI16 0
#endif
    static constexpr int kEpiWarps = (I8 && DLLM_I8_EPI16) ? 16 : 8;
    static_assert(kPre >= 1 && kPre + 1 < kSlots, "half 0 cannot run further ahead than the rings are deep");
    static constexpr int kAccCols = kAccBufs * 128;
    static constexpr int kWBytes = WL_TILE_N * WL_TILE_K * CB / 8;
    static constexpr int kPBytes = I8 ? 0 : 128 * 8;                // (int8: no per-group parameters — scale and zero-point leave in the epilogue)
    static constexpr int kWStage = KBS * (kWBytes + kPBytes);
    // output staging: one accumulator half of this CTA as bf16, [2 parts][q <= 64 tokens][128 columns] = 32 KB, written
    // by the epilogue warps and stored with two bulk tensor copies (the direct 2-byte stores cost 29 % of the kernel)
    static constexpr int kOutBytes = 2 * 64 * 128 * 2;
    static constexpr int kSmemBudget = 220 * 1024;
    static constexpr int kWStagesRaw = (kSmemBudget - kSlots * (kXStage + kAStage) - kOutBytes) / kWStage;
    static constexpr int kWStages = kWStagesRaw > 12 ? 12 : kWStagesRaw;   // 1024 cycles of MMAs per stage: a few stages cover the L2 latency
    static constexpr int kAOffset = kSlots * kXStage;
    static constexpr int kOutOffset = kAOffset + kSlots * kAStage;
    static constexpr int kWOffset = kOutOffset + kOutBytes;
    static constexpr int kBarOffset = kWOffset + kWStages * kWStage;
    static constexpr int kNumBars = 2 * kWStages + 3 * kSlots + 2 * kAccBufs;
    // int8: {zp x row sum, weight scale x activation step} of the tile's <= 256 tokens, double-buffered by tile parity
    static constexpr int kTabOffset = kBarOffset + ((kNumBars * 8 + 16 + 15) & ~15);
    static constexpr int kTabBytes = I8 ? 2 * 256 * 8 : 0;
    static constexpr int kTotal = kTabOffset + kTabBytes + 1024;
    static_assert(kXStage % 1024 == 0 && kWStage % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte aligned stages");
    static_assert(kWStages >= kSlots, "W ring must be at least as deep as the A ring");
    static_assert(kTotal <= 227 * 1024, "shared memory over-subscribed");
    static_assert(kAccCols + (kASmem ? 0 : kSlots * kSlotCols) <= kTmemCols, "TMEM over-subscribed");
};

// non-blocking probe (try_wait may suspend the thread for a system-dependent time; test_wait never does)
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ uint32_t pack_bf16_bias(uint32_t lo, uint32_t hi, float bias) {
    __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(lo) + bias, __uint_as_float(hi) + bias);
    return *reinterpret_cast<uint32_t *>(&b2);
}
// q (32..64, multiple of 8) accumulator columns of this warp's lanes -> packed bf16 pairs.  Two tensor-memory loads are in
// flight per wait: a load + wait round trip costs ~500 cycles while the tensor pipe is busy, and the half is only handed
// back to the MMA warp when its last column is in registers.
// int8 variant: the accumulators are exact s32 sums; column j (a token) leaves as (sum - zp rowsum_j) * (scale step_j) + bias with
// {zp rowsum_j, scale step_j} read from the shared-memory table `tab` (address of this warp's first column's entry)
__device__ __forceinline__ uint32_t pack_bf16_deq(uint32_t lo, uint32_t hi, float bias, uint32_t tab) {
    uint32_t z0, s0, z1, s1;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(z0), "=r"(s0), "=r"(z1), "=r"(s1) : "r"(tab));
    const float f0 = fmaf((float)((int32_t)lo - (int32_t)z0), __uint_as_float(s0), bias);
    const float f1 = fmaf((float)((int32_t)hi - (int32_t)z1), __uint_as_float(s1), bias);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(f0, f1);
    return *reinterpret_cast<uint32_t *>(&b2);
}
// (Tried for int8: all 64 columns requested before ONE wait, buffer handed back, then the conversion of all of them — half 0 is handed
//  back after 1.9 K cycles instead of 3.0 K, but the conversion of its 64 columns (spilling) then sits in front of half 1's drain:
//  8.7 K cycles per tile boundary instead of 6.3 K, 75.8 against 87.5 steps/s.  With two unpack groups instead of three — 640 threads,
//  96 registers, hardly any spill — the kernel is slower as it stands (83.5 steps/s: the unpack latency of 1.5 K cycles per stage needs
//  three groups) and slower still with that drain (69.2).  The epilogue warps' 6 K cycles per tile — two tensor-memory round trips and
//  8 instructions per output pair, on sub-cores they share with the unpack warps — against 12 K cycles of MMAs at K = 2048 are what
//  is left between this kernel and the tensor pipe.)
struct NoRelease { __device__ __forceinline__ void operator()() const {} };
template <bool I8 = false, class Rel = NoRelease>
__device__ __forceinline__ void drain_columns(uint32_t t_acc, uint32_t q, float bias, uint32_t *pk, uint32_t tab = 0, Rel release = Rel()) {
    auto pack2 = [&](uint32_t lo, uint32_t hi, float b, uint32_t col) -> uint32_t {
        if constexpr (I8) return pack_bf16_deq(lo, hi, b, tab + col * 8u);
        else return pack_bf16_bias(lo, hi, b);
    };
#pragma unroll
    for (int p2 = 0; p2 < 2; ++p2) {
        const uint32_t c = (uint32_t)(p2 * 32);
        uint32_t v[32];
        const bool a16 = c + 16 <= q, a8 = !a16 && c + 8 <= q, b16 = c + 32 <= q, b8 = !b16 && c + 24 <= q;
        if (a16) tmem_ld16(t_acc + c, v); else if (a8) tmem_ld8(t_acc + c, v);
        if (b16) tmem_ld16(t_acc + c + 16, v + 16); else if (b8) tmem_ld8(t_acc + c + 16, v + 16);
        tmem_ld_wait();
        if (p2 == (q > 32 ? 1 : 0)) release();           // the last columns are in registers
        if (a16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[p2 * 16 + j] = pack2(v[2 * j], v[2 * j + 1], bias, c + 2 * j);
        } else if (a8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[p2 * 16 + j] = pack2(v[2 * j], v[2 * j + 1], bias, c + 2 * j);
        }
        if (b16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[p2 * 16 + 8 + j] = pack2(v[16 + 2 * j], v[16 + 2 * j + 1], bias, c + 16 + 2 * j);
        } else if (b8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[p2 * 16 + 8 + j] = pack2(v[16 + 2 * j], v[16 + 2 * j + 1], bias, c + 16 + 2 * j);
        }
    }
}
// packed pairs -> the staging buffer [token][128 columns] (this lane's column), tokens [0, q)
__device__ __forceinline__ void stage_columns(uint32_t sbase, uint32_t q, const uint32_t *pk) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        if ((uint32_t)(g * 8) < q) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const uint32_t w = pk[(g * 8 + j) >> 1];
                sts_u16(sbase + (uint32_t)(g * 8 + j) * 256u, w);
                sts_u16(sbase + (uint32_t)(g * 8 + j + 1) * 256u, w >> 16);
            }
        }
    }
}

template <int CB, int NDQ, bool I8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((4 + 4 * NDQ + CfgP2<CB, I8>::kEpiWarps) * 32, 1)
umma_qlinear_pair2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                          const __grid_constant__ RsMaps rs, const Pair2Args a) {
    // I8: the int8 variant (int8 denoise mode).  tmap_x = the int8 activations [M, K bytes] (box {128 k, ntok/2 tokens}: one
    // SWIZZLE_128B tile holds both k-blocks of a stage), A = the codes as unsigned bytes (16 tensor-memory columns per k-block),
    // tcgen05.mma.kind::i8 with K = 32 per instruction, s32 accumulators, and the dequantization in the epilogue.
    using C = CfgP2<CB, I8>;
    constexpr int KBS = C::KBS, kACols = C::kACols;
    constexpr int SW = C::kWStages, A = C::kSlots;
    constexpr int kEpiWarp0 = 4 + 4 * NDQ;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_w = smem + C::kWOffset;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::kBarOffset);
    uint64_t *wfull = bars, *wempty = bars + SW;                          // W ring
    uint64_t *xfull = bars + 2 * SW, *xempty = xfull + A, *afull = xempty + A;   // X ring + A ring (xempty frees both)
    constexpr int NB = C::kAccBufs;
    uint64_t *tfull = afull + A, *tempty = tfull + NB;                    // accumulator buffers
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + NB);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                              // 0 = leader: issues the MMAs
    const uint32_t ntok = a.ntok, half_rows = ntok >> 1, q = ntok >> 2, nh = ntok >> 1;   // rows per CTA, rows per CTA per half, MMA N
    const uint32_t slab = half_rows * 128u;                               // bytes of one k-block of activations in this CTA
    const uint32_t KB = a.k_blocks, n_pairs_grid = gridDim.x >> 1, tile0 = blockIdx.x >> 1;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        if (a.y_f32 == nullptr) prefetch_tmap(&tmap_y);
        for (int s = 0; s < SW; ++s) { mbar_init(wfull + s, 1); mbar_init(wempty + s, 4); }
        for (int s = 0; s < A; ++s) { mbar_init(xfull + s, 1); mbar_init(xempty + s, 1); mbar_init(afull + s, 8); }   // afull: dequant warps of both CTAs
        for (int i = 0; i < NB; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 2 * C::kEpiWarps); }            // tempty: the epilogue warps of both CTAs
        fence_barrier_init();
    }
    cluster_sync_all();                                                   // barriers of both CTAs exist before anybody signals them
    if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_col0 = C::kAccCols;
    // (setmaxnreg — 40 registers for the producer / MMA warps, 56 for the unpack warps, 136 for the epilogue warps — does not help:
    //  ptxas then allocates the WHOLE kernel for the smallest budget, 1560 bytes of spills.)
    if (warp == 0) {
        // ===================== activation producer: this CTA's half of the tile's tokens, bytes reported to the leader =====================
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const uint32_t xfull_leader = leader_addr(xfull);
        uint32_t it = 0;
        uint32_t gate_ok = 0xffffffffu;
        for (uint32_t tile = tile0; tile < a.tiles; tile += n_pairs_grid) {
            const uint32_t mt = (tile / a.n_pairs + a.mt_rot) % a.m_tiles;
            if (a.gate != nullptr) {
                const uint32_t tok = mt * ntok, slice = tok / a.gate_rows;
                const uint32_t gidx = slice * 8u + (tok - slice * a.gate_rows) / a.gate_sub;
                if (slice != a.gate_self && gidx != gate_ok) {
                    // these rows are being stored into this GPU's activation buffer by rank `slice` right now
                    const uint32_t *flag = a.gate + gidx;
                    unsigned long long t0 = 0;
                    uint32_t spins = 0;
                    for (;;) {
                        uint32_t v;
                        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if ((int32_t)(v - a.gate_target) >= 0) break;
                        if ((++spins & 0xff) == 0) {
                            unsigned long long now;
                            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                            if (t0 == 0) t0 = now;
                            else if (now - t0 > 4000000000ull) { if (lane == 0) atomicExch(a.gate_err, 100u + slice); break; }
                        }
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");     // the peer's generic-proxy stores -> this CTA's bulk tensor loads
                    gate_ok = gidx;
                }
            }
            for (uint32_t kb = 0; kb < KB; kb += KBS, ++it) {
                const uint32_t s = it % A, ph = (it / A) & 1;
                mbar_wait(xempty + s, ph ^ 1);
                if (elect_one()) {
                    if (I8 && !(a.dbg & 8)) {
                        // box {128 k, ntok/2 tokens}: both k-blocks of the stage; tokens past M and k past K are zero-filled
                        if (rank == 0) mbar_arrive_expect_tx(xfull + s, 2 * slab);
                        tma_load_2d_pair(smem + s * C::kXStage, &tmap_x, xfull_leader + s * 8, (int)(kb * WL_TILE_K), (int)(mt * ntok + rank * half_rows));
                    } else if (!(a.dbg & 8)) {
                        if (rank == 0) mbar_arrive_expect_tx(xfull + s, 2 * KBS * slab);
                        // box {64 k, ntok/2 tokens, KBS k-blocks}; tokens past M and k-blocks past K are zero-filled
                        tma_load_3d_pair(smem + s * C::kXStage, &tmap_x, xfull_leader + s * 8, 0, (int)(mt * ntok + rank * half_rows), (int)kb);
                    } else if (rank == 0) {
                        mbar_arrive(xfull + s);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===================== weight producers: this CTA's own 128 columns =====================
        const uint32_t me = (uint32_t)(warp - 2);
        uint32_t it = 0;
        for (uint32_t tile = tile0; tile < a.tiles; tile += n_pairs_grid) {
            const uint32_t nc = tile % a.n_pairs;
            const uint32_t nt = 2 * nc + rank < a.n_tiles ? 2 * nc + rank : a.n_tiles - 1;    // odd tile count: the last pair's second CTA re-reads the last tile (stores nothing)
            const uint8_t *wsrc = a.packed + ((size_t)nt * KB) * C::kWBytes;
            const uint2 *psrc = a.dqparams + (size_t)nt * 128;
            for (uint32_t kb = 0; kb < KB; kb += KBS, ++it) {
                if ((it & 1) != me) continue;
                const uint32_t nk = KB - kb < (uint32_t)KBS ? KB - kb : (uint32_t)KBS;
                const uint32_t s = it % SW, ph = (it / SW) & 1;
                mbar_wait(wempty + s, ph ^ 1);
                if (elect_one()) {
                    uint8_t *stage = smem_w + s * C::kWStage;
                    mbar_arrive_expect_tx(wfull + s, nk * (C::kWBytes + C::kPBytes));
                    bulk_load(stage, wsrc + (size_t)kb * C::kWBytes, nk * C::kWBytes, wfull + s);
                    if constexpr (C::kPBytes != 0) {
                        for (uint32_t sub = 0; sub < nk; ++sub) {
                            const uint32_t g = (kb + sub) / a.group_kb;
                            bulk_load(stage + KBS * C::kWBytes + sub * C::kPBytes, psrc + (size_t)g * a.Npad, C::kPBytes, wfull + s);
                        }
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (leader CTA; one elected lane): per stage, half 0 then half 1 =====================
        // c = f32, a = b = bf16 — int8: c = s32 (2 << 4), a = u8 (0 << 7), b = s8 (1 << 10) — K-major A and B, N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (I8 ? ((2u << 4) | (0u << 7) | (1u << 10)) : ((1u << 4) | (1u << 7) | (1u << 10))) |
                               ((nh >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        // one int8 MMA: 32 k (k-step sub * 2 + k2 of the stage in slot s) into accumulator d
        auto mma_i8 = [&](uint32_t d, uint32_t s, uint32_t sub, uint32_t k2, uint64_t bdesc, uint32_t acc) {
            const uint64_t step = (uint64_t)((sub * 2 + k2) * 2);
            if constexpr (C::kASmem)
                umma_ss_pair_i8(d, make_b_desc(smem_u32(smem + C::kAOffset + s * C::kAStage)) + step, bdesc + step, idesc, acc);
            else
                umma_ts_pair_i8(d, tmem_base + a_col0 + s * C::kSlotCols + sub * kACols + k2 * 8, bdesc + step, idesc, acc);
        };
        uint32_t it = 0, n_item = 0;
        if constexpr (C::kPre > 1) {
            // Half 0 runs kPre stages AHEAD of half 1: the issue order is h0(0) .. h0(kPre-1), then h1(j), h0(j + kPre) for every j.
            // At a tile boundary half 0's accumulator is then complete (and being drained) while half 1's last kPre stages still
            // run, and the next tile's first kPre stages of half 0 are issued as soon as that buffer is back — before half 1's buffer,
            // which is drained second, has been handed back.  (In lock step the tensor pipe idled for both drains: 6 K cycles per
            // tile, stage timeline.)  A stage's slot is released after its half-1 MMAs; kPre + 1 slots are held.
            for (uint32_t tile = tile0; tile < a.tiles; tile += n_pairs_grid) {
                const uint32_t g0 = 2 * n_item, g1 = 2 * n_item + 1;
                const uint32_t b0 = g0 % NB, b1 = g1 % NB, tp0 = (g0 / NB) & 1, tp1 = (g1 / NB) & 1;
                ++n_item;
                const uint32_t n_st = (KB + KBS - 1) / KBS;
                auto wait_stage = [&](uint32_t j) {
                    const uint32_t its = it + j, s = its % A, ph = (its / A) & 1;
                    mbar_wait(xfull + s, ph);          // both CTAs' activation halves landed
                    mbar_wait(afull + s, ph);          // both CTAs' A slots written to tensor memory
                    tc_fence_after();
                };
                auto issue = [&](uint32_t j, uint32_t h) {
                    const uint32_t its = it + j, s = its % A;
                    const uint32_t kb = j * KBS, nk = KB - kb < (uint32_t)KBS ? KB - kb : (uint32_t)KBS;
                    const uint32_t stage_addr = smem_u32(smem + s * C::kXStage) + h * q * 128u;
                    const uint32_t a_tmem = tmem_base + a_col0 + s * C::kSlotCols;
                    const uint32_t d_tmem = tmem_base + (h ? b1 : b0) * 128;
                    if (lane == 0) TRACE2(1 + h, its);
                    if (elect_one()) {
                        if (!(a.dbg & 1)) {
                            for (uint32_t sub = 0; sub < nk; ++sub) {
                                if constexpr (I8) {
                                    const uint64_t bdesc = make_b_desc(stage_addr);
#pragma unroll
                                    for (int k2 = 0; k2 < 2; ++k2)
                                        mma_i8(d_tmem, s, sub, (uint32_t)k2, bdesc, (j == 0 && sub == 0 && k2 == 0) ? 0u : 1u);
                                } else {
                                    const uint64_t bdesc = make_b_desc(stage_addr + sub * slab);
#pragma unroll
                                    for (int k4 = 0; k4 < WL_TILE_K / 16; ++k4)
                                        umma_ts_pair(d_tmem, a_tmem + sub * kACols + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, (j == 0 && sub == 0 && k4 == 0) ? 0u : 1u);
                                }
                            }
                        }
                        if (j + 1 == n_st) umma_commit_pair(tfull + (h ? b1 : b0));
                        if (h == 1) umma_commit_pair(xempty + s);      // frees the activation stage and the A slot in both CTAs
                    }
                    __syncwarp();
                };
                const uint32_t pre = n_st < (uint32_t)C::kPre ? n_st : (uint32_t)C::kPre;
                mbar_wait(tempty + b0, tp0 ^ 1);                       // the buffer's previous user drained (both CTAs)
                tc_fence_after();
                for (uint32_t j = 0; j < pre; ++j) { wait_stage(j); issue(j, 0); }
                mbar_wait(tempty + b1, tp1 ^ 1);
                tc_fence_after();
                // (interleaving the two halves' MMAs k-step by k-step, so that consecutive MMAs never accumulate into the same
                //  buffer, changes nothing: 86.7 against 87.6 steps/s)
                for (uint32_t j = 0; j < n_st; ++j) {
                    issue(j, 1);
                    if (j + pre < n_st) { wait_stage(j + pre); issue(j + pre, 0); }
                }
                it += n_st;
            }
        } else
        for (uint32_t tile = tile0; tile < a.tiles; tile += n_pairs_grid) {
            // tile n, half h: buffer (2 n + h) % NB, in its (2 n + h) / NB-th use
            const uint32_t g0 = 2 * n_item, g1 = 2 * n_item + 1;
            const uint32_t b0 = g0 % NB, b1 = g1 % NB, tp0 = (g0 / NB) & 1, tp1 = (g1 / NB) & 1;
            ++n_item;
            bool ready = false;
            for (uint32_t kb = 0; kb < KB; kb += KBS, ++it) {
                const uint32_t nk = KB - kb < (uint32_t)KBS ? KB - kb : (uint32_t)KBS;
                const uint32_t s = it % A, ph = (it / A) & 1;
                if (!ready) {
                    mbar_wait(xfull + s, ph);          // both CTAs' activation halves landed
                    mbar_wait(afull + s, ph);          // both CTAs' A slots written to tensor memory
                    tc_fence_after();
                }
                const uint32_t stage_addr = smem_u32(smem + s * C::kXStage);
                const uint32_t a_tmem = tmem_base + a_col0 + s * C::kSlotCols;
                const bool first = kb == 0, last = kb + KBS >= KB;
                if (lane == 0) TRACE2(1, it);
                // ---- token half 0 ----
                if (first) { mbar_wait(tempty + b0, tp0 ^ 1); tc_fence_after(); }     // the buffer's previous user drained (both CTAs)
                if (elect_one()) {
                    if (!(a.dbg & 1)) {
                        for (uint32_t sub = 0; sub < nk; ++sub) {
                            if constexpr (I8) {
                                // 32 k (32 bytes of the 128-byte row, 8 tensor-memory columns of A) per MMA: two per k-block
                                const uint64_t bdesc = make_b_desc(stage_addr);
#pragma unroll
                                for (int k2 = 0; k2 < 2; ++k2)
                                    mma_i8(tmem_base + b0 * 128, s, sub, (uint32_t)k2, bdesc, (first && sub == 0 && k2 == 0) ? 0u : 1u);
                            } else {
                                const uint64_t bdesc = make_b_desc(stage_addr + sub * slab);
#pragma unroll
                                for (int k4 = 0; k4 < WL_TILE_K / 16; ++k4)
                                    umma_ts_pair(tmem_base + b0 * 128, a_tmem + sub * kACols + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, (first && sub == 0 && k4 == 0) ? 0u : 1u);
                            }
                        }
                    }
                    if (last) umma_commit_pair(tfull + b0);
                }
                __syncwarp();
                // peek at the next stage while those MMAs run (never blocks: half 1 must be issued first)
                ready = false;
                if (!last) {
                    const uint32_t s2 = (it + 1) % A, ph2 = ((it + 1) / A) & 1;
                    if (mbar_try(xfull + s2, ph2) && mbar_try(afull + s2, ph2)) { tc_fence_after(); ready = true; }
                }
                // ---- token half 1 ----
                if (first) { mbar_wait(tempty + b1, tp1 ^ 1); tc_fence_after(); }
                if (lane == 0) TRACE2(2, it);
                if (elect_one()) {
                    if (!(a.dbg & 1)) {
                        for (uint32_t sub = 0; sub < nk; ++sub) {
                            if constexpr (I8) {
                                const uint64_t bdesc = make_b_desc(stage_addr + q * 128u);
#pragma unroll
                                for (int k2 = 0; k2 < 2; ++k2)
                                    mma_i8(tmem_base + b1 * 128, s, sub, (uint32_t)k2, bdesc, (first && sub == 0 && k2 == 0) ? 0u : 1u);
                            } else {
                                const uint64_t bdesc = make_b_desc(stage_addr + sub * slab + q * 128u);
#pragma unroll
                                for (int k4 = 0; k4 < WL_TILE_K / 16; ++k4)
                                    umma_ts_pair(tmem_base + b1 * 128, a_tmem + sub * kACols + k4 * 8, bdesc + (uint64_t)(k4 * 2), idesc, (first && sub == 0 && k4 == 0) ? 0u : 1u);
                            }
                        }
                    }
                    if (last) umma_commit_pair(tfull + b1);
                    umma_commit_pair(xempty + s);      // frees the activation stage and the A slot in both CTAs
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < kEpiWarp0) {
        // ===================== dequant warps: own 128 columns -> own tensor memory; arrive at the leader =====================
        const uint32_t grp = (uint32_t)(warp - 4) >> 2;
        const int quarter = warp & 3;
        const int n_local = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a_col0;
        const uint32_t afull_leader = leader_addr(afull);
        uint32_t it = 0;
        for (uint32_t tile = tile0; tile < a.tiles; tile += n_pairs_grid) {
            for (uint32_t kb = 0; kb < KB; kb += KBS, ++it) {
                if (it % NDQ != grp) continue;
                const uint32_t nk = KB - kb < (uint32_t)KBS ? KB - kb : (uint32_t)KBS;
                const uint32_t sw = it % SW, wph = (it / SW) & 1;
                const uint32_t sl = it % A, aph = (it / A) & 1;
                if (it >= (uint32_t)A) mbar_wait(xempty + sl, aph ^ 1);     // MMAs of the slot's previous user completed
                mbar_wait(wfull + sw, wph);
                tc_fence_after();
                if (quarter == 0 && lane == 0) TRACE2(0, it);
                const uint8_t *stage = smem_w + sw * C::kWStage;
                if constexpr (I8) {
                    // the codes as unsigned bytes in k order: no zero-point, no scale (both leave in the epilogue).  Both k-blocks of
                    // the stage in one straight-line block: their loads and unpack interleave
                    uint32_t vals[KBS][16];
#pragma unroll
                    for (uint32_t sub = 0; sub < (uint32_t)KBS; ++sub) {
                        if (sub >= nk) continue;
                        if (!(a.dbg & 2)) unpack_kblock_u8<CB>(reinterpret_cast<const uint4 *>(stage + sub * C::kWBytes), n_local, vals[sub]);
                        else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) vals[sub][e] = (uint32_t)(e + n_local);
                        }
                    }
#pragma unroll
                    for (uint32_t sub = 0; sub < (uint32_t)KBS; ++sub) {
                        if (sub >= nk) continue;
                        if constexpr (C::kASmem) {
                            // row n_local of the stage's A tile: 128 bytes = both k-blocks, 16-byte chunk c at c ^ (row & 7)
                            const uint32_t rowa = smem_u32(smem + C::kAOffset + sl * C::kAStage) + (uint32_t)n_local * 128u;
#pragma unroll
                            for (uint32_t c4 = 0; c4 < 4; ++c4) {
                                const uint32_t c = sub * 4 + c4;
                                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
                                             :: "r"(rowa + ((c ^ ((uint32_t)n_local & 7u)) << 4)), "r"(vals[sub][4 * c4]), "r"(vals[sub][4 * c4 + 1]),
                                                "r"(vals[sub][4 * c4 + 2]), "r"(vals[sub][4 * c4 + 3]) : "memory");
                            }
                        } else {
                            tmem_st16(lane_addr + sl * C::kSlotCols + sub * kACols, vals[sub]);
                        }
                    }
                } else
                for (uint32_t sub = 0; sub < nk; ++sub) {
                    const uint4 *wpk = reinterpret_cast<const uint4 *>(stage + sub * C::kWBytes);
                    {
                        const uint2 prm = lds64(smem_u32(stage + KBS * C::kWBytes + sub * C::kPBytes) + (uint32_t)n_local * 8u);
                        uint32_t vals[32];
                        if (!(a.dbg & 2)) dequant_kblock<CB>(wpk, n_local, prm.x, prm.y, vals);
                        else {
#pragma unroll
                            for (int e = 0; e < 32; ++e) vals[e] = prm.x + e;
                        }
                        tmem_st32(lane_addr + sl * C::kSlotCols + sub * kACols, vals);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(wempty + sw);
                if constexpr (C::kASmem) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // these stores -> the MMA's descriptor reads
                } else {
                    tmem_st_wait();
                    tc_fence_before();
                }
                __syncwarp();
                if (lane == 0) {
                    if constexpr (!C::kASmem) mbar_arrive_cluster_tmem(afull_leader + sl * 8);
                    else mbar_arrive_cluster(afull_leader + sl * 8);
                }
                if (quarter == 0 && lane == 0) TRACE2(3, it);
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue warps: drain half 0, re-arm it, drain half 1, re-arm it =====================
        // Eight warps: warp%4 = TMEM lane quarter, `part` = which CTA's token rows (accumulator columns [part*q, (part+1)*q) of a
        // half).  Two phases per half.  Phase 1 is on the tensor pipe's critical path (the accumulator half is single-
        // buffered): the warp's q <= 64 columns are pulled into registers as packed bf16 (<= 32 registers) and the half is
        // handed back to the MMA warp BEFORE anything is stored.  Phase 2, the stores, then overlaps the next MMAs.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const int quarter = warp & 3;
        constexpr uint32_t SPLIT = C::kEpiWarps / 8, kEpiThreads = C::kEpiWarps * 32;      // warps per (lane quarter, CTA part)
        const uint32_t part4 = (uint32_t)(warp - kEpiWarp0) >> 2;
        const uint32_t part = part4 / SPLIT, wsub = part4 % SPLIT, wq = q / SPLIT;          // this warp: columns [wsub * wq, + wq) of its part
        const uint32_t tempty_leader = leader_addr(tempty);
        const size_t ldy = a.N;
        uint32_t n_item = 0;
        for (uint32_t tile = tile0; tile < a.tiles; tile += n_pairs_grid) {
            const uint32_t mt = (tile / a.n_pairs + a.mt_rot) % a.m_tiles, nc = tile % a.n_pairs;
            const uint32_t nt = 2 * nc + rank;
            const uint32_t n = nt * 128 + quarter * 32 + lane;
            const bool n_ok = nt < a.n_tiles && n < a.N && !(a.dbg & 64);
            const float bias = (a.bias != nullptr && n_ok) ? __ldg(a.bias + n) : 0.f;
            uint32_t tab = 0;
            if constexpr (I8) {
                // {zp x row sum, weight scale x activation step} of the tile's tokens, in the order the accumulator columns hold
                // them (entry part * half_rows + h * q + column); double-buffered by tile parity, so one barrier per tile is enough
                tab = smem_u32(smem + C::kTabOffset) + (n_item & 1u) * 2048u;
                const uint32_t e = threadIdx.x - (uint32_t)kEpiWarp0 * 32u;
                if (e < ntok) {
                    const uint32_t tok = mt * ntok + e;
                    int32_t zs = 0;
                    float sc = 0.f;
                    if (tok < a.M) { zs = a.i8_zp * __ldg(a.i8_rowsum + tok); sc = __ldg(a.i8_rowscale + tok); }
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(tab + e * 8u), "r"((uint32_t)zs), "r"(__float_as_uint(sc)) : "memory");
                }
                named_bar_sync(1, kEpiThreads);
            }
            if (a.y_f32 == nullptr) {
                // bf16 output.  Both accumulator halves are on the tensor pipe's critical path (they are single-buffered), so they
                // are drained back to back: half 0 -> registers -> handed back -> written to the staging buffer (shared-memory
                // stores only); half 1 -> registers -> handed back; THEN the stores: half 0 leaves as bulk tensor copies, the
                // staging buffer is re-used for half 1.  (Storing half 0 before draining half 1 — the first version — kept half 1
                // for 1.9 K cycles longer at every tile boundary: stage timeline, DLLM_UMMA_DBG=128.)
                const uint32_t stg = smem_u32(smem + C::kOutOffset);
                const uint32_t stg_mine = stg + (part * q + wsub * wq) * 256u + (uint32_t)(quarter * 32 + lane) * 2u;
                auto store_half = [&](uint32_t h) {                 // one elected thread: the staged half -> global / peer memory
                    if (warp == kEpiWarp0 && lane == 0 && nt < a.n_tiles) {
                        const int tokA = (int)(mt * ntok + h * q);
                        if (a.rs_world == 0) {
                            tma_store_2d(&tmap_y, stg, (int)(nt * 128), tokA);
                            tma_store_2d(&tmap_y, stg + q * 256u, (int)(nt * 128), tokA + (int)half_rows);
                        } else {
                            // reduce-scatter fused into the epilogue: each q-row block goes to its owner's receive buffer
                            // (q divides rs_rows, so a block never straddles two owners)
#pragma unroll
                            for (int b2 = 0; b2 < 2; ++b2) {
                                const uint32_t tk = (uint32_t)tokA + (uint32_t)b2 * half_rows;
                                if (tk < a.M) {
                                    const uint32_t owner = tk / a.rs_rows;
                                    tma_store_2d(&rs.m[owner], stg + (uint32_t)b2 * q * 256u, (int)(nt * 128),
                                                 (int)(a.rs_rank * a.rs_rows + (tk - owner * a.rs_rows)));
                                }
                            }
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                };
                const bool st_on = !(a.dbg & 64);
                // staging buffer free?  Checked BEFORE the wait for the accumulator (off the critical path: the previous tile's
                // half-1 store has had a whole tile's time)
                if (st_on) {
                    if (warp == kEpiWarp0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    named_bar_sync(1, kEpiThreads);
                }
                uint32_t pk[32];
#pragma unroll 1
                for (uint32_t h = 0; h < 2; ++h) {
                    const uint32_t gb = 2 * n_item + h, buf = gb % NB, tph = (gb / NB) & 1;     // buffer of (tile, half) and its use parity
                    const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 128 + part * q + wsub * wq;
                    mbar_wait(tfull + buf, tph);
                    tc_fence_after();
                    if (warp == kEpiWarp0 && lane == 0) TRACE2(4 + h, n_item);
                    // this warp's columns are out: 16 such arrivals free the buffer
                    auto release = [&]() {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive_cluster_tmem(tempty_leader + buf * 8);
                        }
                        if (warp == kEpiWarp0 && lane == 0) TRACE2(6 + h, n_item);
                    };
                    if constexpr (I8) {
                        // (the buffer is handed back as soon as the last column is in registers, before the last batch is converted)
#if defined(DLLM_I8_LATE_RELEASE)
                        drain_columns<true>(t_acc, wq, bias, pk, tab + (part * half_rows + h * q + wsub * wq) * 8u);
                        release();
#else
                        drain_columns<true>(t_acc, wq, bias, pk, tab + (part * half_rows + h * q + wsub * wq) * 8u, release);
#endif
                    } else {
                        // (handing the buffer back before the last 32 columns are converted, as the int8 variant may, costs the bf16
                        //  kernel registers it does not have: 192 instead of 108 bytes spilled, 63.7 against 67.4 steps/s)
                        drain_columns<false>(t_acc, wq, bias, pk);
                        release();
                    }
                    if (h == 0 && st_on) stage_columns(stg_mine, wq, pk);         // registers -> staging [part][token][128 columns]
                }
                if (st_on) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    named_bar_sync(1, kEpiThreads);
                    store_half(0);                                               // rows past M / columns past N are clipped by the TMA unit
                    if (warp == kEpiWarp0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    named_bar_sync(1, kEpiThreads);
                    stage_columns(stg_mine, wq, pk);                             // pk still holds half 1
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    named_bar_sync(1, kEpiThreads);
                    store_half(1);
                }
            } else {
#pragma unroll 1
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t gb = 2 * n_item + h, buf = gb % NB, tph = (gb / NB) & 1;
                const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 128 + part * q + wsub * wq;
                const uint32_t tok0 = mt * ntok + part * half_rows + h * q + wsub * wq;     // token of this warp's first column
                const uint32_t n_tok = tok0 >= a.M ? 0u : (a.M - tok0 < wq ? a.M - tok0 : wq);   // valid rows among its wq
                {
                    // f32 (and optionally bf16) output: the stack's last layer only.  Columns are stored as they are read.
                    mbar_wait(tfull + buf, tph);
                    tc_fence_after();
#pragma unroll 1
                    for (uint32_t c0 = 0; c0 < wq; c0 += 8) {
                        uint32_t v[8];
                        tmem_ld8(t_acc + c0, v);
                        tmem_ld_wait();
                        if (c0 + 8 >= wq) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) {
                                mbar_arrive_cluster_tmem(tempty_leader + buf * 8);
                            }
                        }
                        if (!n_ok) continue;
                        const size_t o = (size_t)(tok0 + c0) * ldy + n;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (c0 + j < n_tok) {
                                float f;
                                if constexpr (I8) {
                                    uint32_t zs, sc;
                                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(zs), "=r"(sc) : "r"(tab + (part * half_rows + h * q + wsub * wq + c0 + j) * 8u));
                                    f = fmaf((float)((int32_t)v[j] - (int32_t)zs), __uint_as_float(sc), bias);
                                } else {
                                    f = __uint_as_float(v[j]) + bias;
                                }
                                a.y_f32[o + (size_t)j * ldy] = f;
                                if (a.y_bf16) a.y_bf16[o + (size_t)j * ldy] = __float2bfloat16_rn(f);
                            }
                        }
                    }
                }
            }
            }
            ++n_item;
        }
        if (warp == kEpiWarp0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // output writes complete before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // the partner may still signal barriers / read shared memory of this CTA
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

// stream-K fix-up: tiles that were cut by a CTA boundary get y = sum of their partial tiles (in CTA
// order: deterministic) + bias.  One block per output tile; complete tiles return immediately.
template <int NTOK>
__global__ void __launch_bounds__(256)
umma_streamk_fixup_kernel(const UmmaArgs a, uint32_t grid_main) {
    const uint32_t tile = blockIdx.x;
    const uint64_t KB = a.k_blocks, U = a.units, G = grid_main;
    const uint64_t t0 = (uint64_t)tile * KB, t1 = t0 + KB;
    const uint32_t c_lo = (uint32_t)(((t0 + 1) * G - 1) / U);     // CTA owning the tile's first unit
    const uint32_t c_hi = (uint32_t)((t1 * G - 1) / U);           // CTA owning its last unit
    if (c_lo == c_hi) return;                                     // one CTA saw the whole tile: written directly
    const uint32_t nt = tile % a.n_tiles, mt = tile / a.n_tiles;
    for (uint32_t e = threadIdx.x; e < NTOK * 128; e += blockDim.x) {
        const uint32_t ml = e >> 7, nl = e & 127;
        const uint32_t m = mt * NTOK + ml, n = nt * 128 + nl;
        if (m >= a.M || n >= a.N) continue;
        float v = 0.f;
        for (uint32_t c = c_lo; c <= c_hi; ++c) {
            const uint64_t u0 = U * c / G;
            const uint32_t slot = c * 2 + (u0 >= t0 ? 0u : 1u);   // CTA starts inside this tile -> its first item
            v += a.partial[(size_t)slot * (NTOK * 128) + e];
        }
        if (a.bias) v += __ldg(a.bias + n);
        if (a.y_f32) a.y_f32[(size_t)m * a.N + n] = v;
        if (a.y_bf16) a.y_bf16[(size_t)m * a.N + n] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

template <int CB, int NTOK, int KBS, int NDQ>
int32_t launch_umma(dllm_ctx *ctx, const dllm_qweight *qw, const void *x_bf16, size_t M, float *y_f32, void *y_bf16) {
    using C = Cfg<CB, NTOK, KBS, NDQ, false>;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    CUtensorMap tmap;
    const bool x3d = qw->K % WL_TILE_K == 0;
    CUresult r;
    if (x3d) {
        // {64 k (contiguous), M tokens (pitch K*2), K/64 k-blocks (pitch 128 B)}: one box = KBS k-block tiles, each
        // NTOK x 64 in the SWIZZLE_128B K-major layout the UMMA descriptor expects
        const cuuint64_t gdim[3] = {WL_TILE_K, (cuuint64_t)M, (cuuint64_t)(qw->K / WL_TILE_K)};
        const cuuint64_t gstride[2] = {(cuuint64_t)qw->K * 2, (cuuint64_t)WL_TILE_K * 2};
        const cuuint32_t box[3] = {WL_TILE_K, (cuuint32_t)NTOK, (cuuint32_t)KBS};
        const cuuint32_t estr[3] = {1, 1, 1};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(x_bf16), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gdim[2] = {(cuuint64_t)qw->K, (cuuint64_t)M};
        const cuuint64_t gstride[1] = {(cuuint64_t)qw->K * 2};
        const cuuint32_t box[2] = {WL_TILE_K, (cuuint32_t)NTOK};
        const cuuint32_t estr[2] = {1, 1};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(x_bf16), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);

    UmmaArgs a;
    a.packed = qw->d_packed; a.dqparams = qw->d_dqparams; a.bias = qw->d_bias;
    a.x3d = x3d ? 1u : 0u;
    a.M = (uint32_t)M; a.N = (uint32_t)qw->N; a.Npad = (uint32_t)(qw->n_tiles * 128);
    a.k_blocks = (uint32_t)qw->k_blocks; a.n_tiles = (uint32_t)qw->n_tiles;
    a.m_tiles = (uint32_t)((M + NTOK - 1) / NTOK);
    a.group_kb = (uint32_t)(qw->group / WL_TILE_K);
    const uint32_t tiles = a.n_tiles * a.m_tiles;
    a.units = (uint64_t)tiles * a.k_blocks;
    a.y_f32 = y_f32; a.y_bf16 = (__nv_bfloat16 *)y_bf16; a.partial = nullptr;
    static const uint32_t dbg_flags = getenv("DLLM_UMMA_DBG") ? (uint32_t)atoi(getenv("DLLM_UMMA_DBG")) : 0u;
    a.dbg = dbg_flags;
    a.trace = nullptr;
    if (a.dbg & 128) {
        DLLM_TRY(ensure_buf(ctx, ctx->lin_flags, 8 * 256 * sizeof(long long)));
        a.trace = (long long *)ctx->lin_flags.p;
        cudaMemsetAsync(a.trace, 0, 8 * 256 * sizeof(long long), ctx->stream);
    }
    // dense problems (>= 4 tiles per SM): whole tiles round-robin; otherwise stream-K over all SMs
    const uint32_t sms = (uint32_t)(ctx->sm_limit > 0 && ctx->sm_limit < ctx->sm_count ? ctx->sm_limit : ctx->sm_count);
    a.stream_k = 0;
    uint32_t grid = tiles < sms ? tiles : sms;
    if (tiles < 4 * sms) {
        const uint64_t g = a.units / (2 * KBS);     // at least two stages of work per CTA
        const uint32_t gk = (uint32_t)(g > sms ? sms : g);
        if (gk > grid || tiles > sms) { a.stream_k = 1; grid = gk; }
    }
    if (a.stream_k) {
        DLLM_TRY(ensure_buf(ctx, ctx->lin_ws, (size_t)2 * grid * NTOK * 128 * sizeof(float)));
        a.partial = (float *)ctx->lin_ws.p;
    }
    // CTA pairs (cta_group::2) for dense problems: two CTAs share every activation tile (half the L2 -> SM activation
    // traffic per SM, which is what bounds the 1-CTA kernel: ~10 TB/s of L2 reads at 8192 tokens)
    static const int pair_env = getenv("DLLM_UMMA_PAIR") ? atoi(getenv("DLLM_UMMA_PAIR")) : 0;
    const uint32_t n_pairs = (a.n_tiles + 1) / 2;
    const bool use_pair = pair_env == 1 && NTOK == 128 && !a.stream_k && x3d && !(a.dbg & 128) && n_pairs * a.m_tiles >= sms / 2;
    using CP = Cfg<CB, NTOK, KBS, NDQ, false, true>;
    CUtensorMap tmap_pair;
    if (use_pair) {
        const cuuint64_t gdim[3] = {WL_TILE_K, (cuuint64_t)M, (cuuint64_t)(qw->K / WL_TILE_K)};
        const cuuint64_t gstride[2] = {(cuuint64_t)qw->K * 2, (cuuint64_t)WL_TILE_K * 2};
        const cuuint32_t box[3] = {WL_TILE_K, (cuuint32_t)(NTOK / 2), (cuuint32_t)KBS};
        const cuuint32_t estr[3] = {1, 1, 1};
        r = enc(&tmap_pair, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(x_bf16), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    DLLM_TRY(ensure_smem_attr(ctx, umma_qlinear_kernel<CB, NTOK, KBS, NDQ>, C::kTotal));
    if (NTOK == 128) DLLM_TRY(ensure_smem_attr(ctx, umma_qlinear_pair_kernel<CB, NTOK, KBS, NDQ>, CP::kTotal));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->prof_on) {   // bracket this launch with events on the launching stream
        while (ctx->prof_ev.size() < 2 * (ctx->prof_n + 1)) {
            cudaEvent_t e;
            CUDA_TRY(ctx, cudaEventCreate(&e));
            ctx->prof_ev.push_back(e);
        }
        ev0 = ctx->prof_ev[2 * ctx->prof_n];
        ev1 = ctx->prof_ev[2 * ctx->prof_n + 1];
        CUDA_TRY(ctx, cudaEventRecord(ev0, ctx->stream));
    }
    if (use_pair) {
        const uint32_t pair_tiles = n_pairs * a.m_tiles;
        const uint32_t pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
        umma_qlinear_pair_kernel<CB, NTOK, KBS, NDQ><<<2 * pairs, (8 + 4 * NDQ) * 32, CP::kTotal, ctx->stream>>>(tmap_pair, a);
    } else {
        static const bool no_pdl = getenv("DLLM_UMMA_NO_PDL") != nullptr;      // experiments only
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3((8 + 4 * NDQ) * 32);
        cfg.dynamicSmemBytes = C::kTotal;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (no_pdl || ctx->no_pdl_once) ? 0 : 1;
        ctx->no_pdl_once = false;
        CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, umma_qlinear_kernel<CB, NTOK, KBS, NDQ>, tmap, a));
    }
    LAUNCH_CHECK(ctx);
    if (ev1) {
        CUDA_TRY(ctx, cudaEventRecord(ev1, ctx->stream));
        ctx->prof_n++;
        ctx->prof_flops += 2.0 * (double)M * (double)qw->K * (double)qw->N;
        // algorithmic bytes: packed codes + (scale, zp) per group and column + bf16 x + output
        ctx->prof_bytes += (double)qw->K * qw->N * qw->bits / 8.0 + (double)(qw->K / qw->group) * qw->N * 8.0 +
                           2.0 * M * qw->K + (y_f32 ? 4.0 : 0.0) * M * qw->N + (y_bf16 ? 2.0 : 0.0) * M * qw->N;
    }
    if (a.dbg & 128) {   // dump the stage timeline of CTA 0 (timing experiments only)
        std::vector<long long> h(8 * 256);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h.data(), a.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        FILE *f = fopen("gpurun_out/umma_trace.csv", "w");
        if (f) {
            fprintf(f, "it,prod_issue,mma_wait,mma_full,mma_aready,dq_start,dq_end,epi_start,epi_end\n");
            for (int i = 0; i < 256; ++i)
                fprintf(f, "%d,%lld,%lld,%lld,%lld,%lld,%lld,%lld,%lld\n", i, h[i], h[256 + i], h[512 + i], h[768 + i], h[1024 + i], h[1280 + i], h[1536 + i], h[1792 + i]);
            fclose(f);
        }
    }
    if (a.stream_k) {
        umma_streamk_fixup_kernel<NTOK><<<tiles, 256, 0, ctx->stream>>>(a, grid);
        LAUNCH_CHECK(ctx);
    }
    return DLLM_OK;
}


// tokens per pair tile: the candidate with the least estimated time on sms/2 CTA pairs — full waves count.  A tile's time is
// not proportional to its width: the weight tile is dequantized once per tile whatever the width, so a 128-token tile costs
// 80 % of a 256-token one.  Measured (scripts/ntok_sweep.py, profiles/r2_ntok_sweep.jsonl; [4096,14336] x 8192 tokens, us per
// wave): 128 -> 22.6, 160 -> 22.9, 192 -> 24.6, 224 -> 26.6, 256 -> 28.4, i.e. max(22.6, 12.6 + 0.0617 ntok): in units of
// token-columns, waves x max(365, ntok + 205).  (The first version charged ntok + 6 and picked 128- to 160-token tiles whenever
// they filled the last wave better: up to 1.64x slower than 256 at 4096 tokens.)  Ties go to the wider tile: fewer weight passes.
static uint32_t pair2_pick_ntok(size_t M, uint32_t n_pairs, uint32_t pairs_hw, uint32_t step = 32) {
    uint32_t best = 0;
    uint64_t best_cost = ~0ull;
    for (uint32_t ntok = 256; ntok >= 128; ntok -= step) {
        const uint64_t m_tiles = (M + ntok - 1) / ntok, tiles = m_tiles * n_pairs;
        const uint64_t waves = (tiles + pairs_hw - 1) / pairs_hw;
        const uint64_t cost = waves * (ntok + 205 > 365 ? ntok + 205 : 365);
        if (cost < best_cost) { best_cost = cost; best = ntok; }
    }
    return best;
}

static int pair2_mode() {
    // DLLM_UMMA_PAIR: unset / 2 = the 256-token CTA-pair kernel for dense problems (default), 0 = 1-CTA kernel only,
    // 1 = the 128-token pair variant of round 1 (experiments)
    static const int m = getenv("DLLM_UMMA_PAIR") ? atoi(getenv("DLLM_UMMA_PAIR")) : 2;
    return m;
}

// dense problems only (whole tiles, K % 64 == 0, enough tiles for the 74 pairs); everything else stays on the 1-CTA kernel
static bool pair2_applicable(const dllm_ctx *ctx, const dllm_qweight *qw, size_t M, const float *y_f32, const void *y_bf16) {
    if (pair2_mode() != 2 || qw->K % WL_TILE_K != 0 || M < 1024) return false;
    if (wl_container_bits(qw->bits) == 8) return false;        // 8-bit tiles leave too few W stages beside the output staging
    // bf16 output leaves through a bulk tensor store: 16-byte aligned rows
    if (!y_f32 && (qw->N % 8 != 0 || (reinterpret_cast<uintptr_t>(y_bf16) & 15u) != 0)) return false;
    const uint32_t n_pairs = (uint32_t)((qw->n_tiles + 1) / 2);
    const uint32_t pairs_hw = (uint32_t)(ctx->sm_limit > 0 && ctx->sm_limit < ctx->sm_count ? ctx->sm_limit : ctx->sm_count) / 2;
    const uint32_t ntok = pair2_pick_ntok(M, n_pairs, pairs_hw);
    return (uint64_t)((M + ntok - 1) / ntok) * n_pairs >= pairs_hw / 2;
}

template <int CB>
int32_t launch_umma_pair2(dllm_ctx *ctx, const dllm_qweight *qw, const void *x_bf16, size_t M, float *y_f32, void *y_bf16,
                          const UmmaRs *rsd = nullptr) {
    using C = CfgP2<CB>;
    constexpr int NDQ = kNDQ;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    Pair2Args a;
    a.packed = qw->d_packed; a.dqparams = qw->d_dqparams; a.bias = qw->d_bias;
    a.y_f32 = y_f32; a.y_bf16 = (__nv_bfloat16 *)y_bf16;
    a.M = (uint32_t)M; a.N = (uint32_t)qw->N; a.Npad = (uint32_t)(qw->n_tiles * 128);
    a.k_blocks = (uint32_t)qw->k_blocks; a.n_tiles = (uint32_t)qw->n_tiles; a.n_pairs = (a.n_tiles + 1) / 2;
    a.group_kb = (uint32_t)(qw->group / WL_TILE_K);
    const uint32_t pairs_hw = (uint32_t)(ctx->sm_limit > 0 && ctx->sm_limit < ctx->sm_count ? ctx->sm_limit : ctx->sm_count) / 2;
    const char *ntok_s = getenv("DLLM_UMMA_NTOK2");                                                 // experiments only (read per launch: scripts/ntok_sweep.py)
    const int ntok_env = ntok_s ? atoi(ntok_s) : 0;
    a.ntok = (ntok_env >= 32 && ntok_env <= 256 && ntok_env % 32 == 0) ? (uint32_t)ntok_env : pair2_pick_ntok(M, a.n_pairs, pairs_hw);
    a.rs_world = 0; a.rs_rank = 0; a.rs_rows = 1;
    a.mt_rot = 0;
    a.gate = nullptr; a.gate_target = 0; a.gate_rows = 1; a.gate_sub = 1; a.gate_self = 0; a.gate_err = nullptr;
    if (rsd) {
        // fused reduce-scatter: 64-row store blocks must not straddle two owners' token ranges
        a.ntok = 256;
        a.rs_world = (uint32_t)rsd->world; a.rs_rank = (uint32_t)rsd->rank; a.rs_rows = (uint32_t)rsd->rows;
        a.mt_rot = (uint32_t)(((size_t)rsd->rank * rsd->rows) / 256);
    }
    if (ctx->gate_armed) {
        // the producer of this GEMM's activations is still delivering the other ranks' token slices (tp.cu): whole 256-token
        // tiles per slice, own slice first
        ctx->gate_armed = false;
        if (ctx->gate_rows % 256 != 0 || M % ctx->gate_rows != 0) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "gated GEMM: slices must be multiples of 256 tokens");
        a.ntok = 256;
        a.gate = ctx->gate_counters; a.gate_target = ctx->gate_target; a.gate_rows = (uint32_t)ctx->gate_rows;
        a.gate_sub = (uint32_t)ctx->gate_sub;
        a.gate_self = ctx->gate_self; a.gate_err = ctx->p2p_err;
        a.mt_rot = (uint32_t)(((size_t)ctx->tp_rank * ctx->gate_rows) / 256);
    }
    a.m_tiles = (uint32_t)((M + a.ntok - 1) / a.ntok);
    a.tiles = a.n_pairs * a.m_tiles;
    static const uint32_t dbg_flags = getenv("DLLM_UMMA_DBG") ? (uint32_t)atoi(getenv("DLLM_UMMA_DBG")) : 0u;
    a.dbg = dbg_flags;
    a.trace = nullptr;
    if (a.dbg & 128) {
        DLLM_TRY(ensure_buf(ctx, ctx->lin_flags, 8 * 256 * sizeof(long long)));
        a.trace = (long long *)ctx->lin_flags.p;
        cudaMemsetAsync(a.trace, 0, 8 * 256 * sizeof(long long), ctx->stream);
    }

    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {WL_TILE_K, (cuuint64_t)M, (cuuint64_t)(qw->K / WL_TILE_K)};
    const cuuint64_t gstride[2] = {(cuuint64_t)qw->K * 2, (cuuint64_t)WL_TILE_K * 2};
    const cuuint32_t box[3] = {WL_TILE_K, a.ntok / 2, (cuuint32_t)C::KBS};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(x_bf16), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    // bf16 output [M, N] row-major: box {128 columns, ntok/4 tokens} = one part of an accumulator half in the staging buffer
    CUtensorMap tmap_y = tmap;
    if (!y_f32) {
        const cuuint64_t ydim[2] = {(cuuint64_t)qw->N, (cuuint64_t)M};
        const cuuint64_t ystride[1] = {(cuuint64_t)qw->N * 2};
        const cuuint32_t ybox[2] = {128, a.ntok / 4};
        const cuuint32_t yestr[2] = {1, 1};
        r = enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y_bf16, ydim, ystride, ybox, yestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (%d)", (int)r);
    }
    RsMaps rsm;
    memset(&rsm, 0, sizeof(rsm));
    if (rsd) {
        // every rank's receive buffer [M, N] bf16 (row block s = the partial sums rank s computed for that rank's tokens), as
        // mapped into this process; box = one staged part {128 columns, 64 tokens}
        const cuuint64_t ydim[2] = {(cuuint64_t)qw->N, (cuuint64_t)M};
        const cuuint64_t ystride[1] = {(cuuint64_t)qw->N * 2};
        const cuuint32_t ybox[2] = {128, a.ntok / 4};
        const cuuint32_t yestr[2] = {1, 1};
        for (int r2 = 0; r2 < rsd->world; ++r2) {
            r = enc(&rsm.m[r2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rsd->recv[r2], ydim, ystride, ybox, yestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled (receive buffer of rank %d) failed (%d)", r2, (int)r);
        }
    }
    DLLM_TRY(ensure_smem_attr(ctx, umma_qlinear_pair2_kernel<CB, NDQ>, C::kTotal));

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->prof_on) {
        while (ctx->prof_ev.size() < 2 * (ctx->prof_n + 1)) {
            cudaEvent_t e;
            CUDA_TRY(ctx, cudaEventCreate(&e));
            ctx->prof_ev.push_back(e);
        }
        ev0 = ctx->prof_ev[2 * ctx->prof_n];
        ev1 = ctx->prof_ev[2 * ctx->prof_n + 1];
        CUDA_TRY(ctx, cudaEventRecord(ev0, ctx->stream));
    }
    const uint32_t pairs = a.tiles < pairs_hw ? a.tiles : pairs_hw;
    static const bool no_pdl = getenv("DLLM_UMMA_NO_PDL") != nullptr;      // experiments only
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3((12 + 4 * NDQ) * 32);
    cfg.dynamicSmemBytes = C::kTotal;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // (no_pdl_once: the kernel before this one leaves SMs free for a reduce / gather kernel it is waiting for — a dependent
    //  launched early would occupy exactly those SMs and sit there at its dependency wait: deadlock until the gates time out)
    cfg.numAttrs = (no_pdl || ctx->no_pdl_once) ? 0 : 1;
    ctx->no_pdl_once = false;
    CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, umma_qlinear_pair2_kernel<CB, NDQ>, tmap, tmap_y, rsm, a));
    LAUNCH_CHECK(ctx);
    if (a.dbg & 128) {   // dump the timeline of cluster 0's leader (timing experiments only)
        std::vector<long long> hst(8 * 256);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(hst.data(), a.trace, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        FILE *f = fopen("gpurun_out/pair2_trace.csv", "w");
        if (f) {
            fprintf(f, "i,dq_start(stage),mma_h0(stage),mma_h1(stage),dq_end(stage),epi_h0_ready(tile),epi_h1_ready(tile),epi_h0_released(tile),epi_h1_released(tile)\n");
            for (int i = 0; i < 256; ++i) {
                fprintf(f, "%d", i);
                for (int r = 0; r < 8; ++r) fprintf(f, ",%lld", hst[r * 256 + i]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    if (ev1) {
        CUDA_TRY(ctx, cudaEventRecord(ev1, ctx->stream));
        ctx->prof_n++;
        ctx->prof_flops += 2.0 * (double)M * (double)qw->K * (double)qw->N;
        ctx->prof_bytes += (double)qw->K * qw->N * qw->bits / 8.0 + (double)(qw->K / qw->group) * qw->N * 8.0 +
                           2.0 * M * qw->K + (y_f32 ? 4.0 : 0.0) * M * qw->N + (y_bf16 ? 2.0 : 0.0) * M * qw->N;
    }
    return DLLM_OK;
}

// two k-blocks (128 k = one quantization group) per pipeline stage
template <int CB>
int32_t launch_umma_ntok(dllm_ctx *ctx, const dllm_qweight *qw, const void *x, size_t M, float *y_f32, void *y_bf16) {
    static const int ntok_env = getenv("DLLM_UMMA_NTOK") ? atoi(getenv("DLLM_UMMA_NTOK")) : 0;     // experiments only
    if constexpr (CB != 8) {
        if (pair2_applicable(ctx, qw, M, y_f32, y_bf16)) return launch_umma_pair2<CB>(ctx, qw, x, M, y_f32, y_bf16);
    }
    if (ctx->gate_armed) {
        ctx->gate_armed = false;
        DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "gated activations handed to a kernel that cannot wait for them (k_umma_gate_supported was not consulted)");
    }
    if (M <= 16) return launch_umma<CB, 16, 2, kNDQ>(ctx, qw, x, M, y_f32, y_bf16);
    if (M <= 32) return launch_umma<CB, 32, 2, kNDQ>(ctx, qw, x, M, y_f32, y_bf16);
    if (M <= 64 || ntok_env == 64) return launch_umma<CB, 64, 2, kNDQ>(ctx, qw, x, M, y_f32, y_bf16);
    static const int kbs_env = getenv("DLLM_UMMA_KBS") ? atoi(getenv("DLLM_UMMA_KBS")) : 0;         // experiments only
    if (kbs_env == 1) return launch_umma<CB, 128, 1, kNDQ>(ctx, qw, x, M, y_f32, y_bf16);
    return launch_umma<CB, 128, 2, kNDQ>(ctx, qw, x, M, y_f32, y_bf16);
}

// row sums of the int8 activations (one warp per token row; K % 4 == 0)
__global__ void __launch_bounds__(256)
rowsum_i8_kernel(const int8_t *__restrict__ x, uint32_t M, uint32_t K, int32_t *__restrict__ sx) {
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const int *p = reinterpret_cast<const int *>(x + (size_t)row * K);
    int acc = 0;
    for (uint32_t i = lane; i < K / 4; i += 32) acc = __dp4a(__ldg(p + i), 0x01010101, acc);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
    if (lane == 0) sx[row] = acc;
}

// fused dequantization of the int8 mode (the int8 denoise stack): row sums and (weight scale x token scale) per token come from the
// activation quantizer, the outputs are floats
struct I8Deq { const int32_t *rowsum; const float *rowscale; float *y_f32; void *y_bf16; };

template <int CB, int NTOK>
int32_t launch_umma_i8(dllm_ctx *ctx, const dllm_qweight *qw, const int8_t *xq, size_t M, int32_t *y, const I8Deq *dq = nullptr) {
    constexpr int KBS = 2, NDQ = kNDQ;
    using C = Cfg<CB, NTOK, KBS, NDQ, true>;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    CUtensorMap tmap;
    // [M tokens, K bytes] row-major; one box = 128 k x NTOK tokens in the SWIZZLE_128B K-major layout
    const cuuint64_t gdim[2] = {(cuuint64_t)qw->K, (cuuint64_t)M};
    const cuuint64_t gstride[1] = {(cuuint64_t)qw->K};
    const cuuint32_t box[2] = {2 * WL_TILE_K, (cuuint32_t)NTOK};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(xq), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);

    if (!dq) {
        DLLM_TRY(ensure_buf(ctx, ctx->lin_flags, M * sizeof(int32_t)));
        rowsum_i8_kernel<<<(unsigned)((M + 7) / 8), 256, 0, ctx->stream>>>(xq, (uint32_t)M, (uint32_t)qw->K, (int32_t *)ctx->lin_flags.p);
        LAUNCH_CHECK(ctx);
    }

    UmmaArgs a;
    a.packed = qw->d_packed; a.dqparams = qw->d_dqparams; a.bias = dq ? qw->d_bias : nullptr;
    a.x3d = (uint32_t)(int32_t)qw->tensor_zp;           // (int8 mode: the zero-point)
    a.M = (uint32_t)M; a.N = (uint32_t)qw->N; a.Npad = (uint32_t)(qw->n_tiles * 128);
    a.k_blocks = (uint32_t)qw->k_blocks; a.n_tiles = (uint32_t)qw->n_tiles;
    a.m_tiles = (uint32_t)((M + NTOK - 1) / NTOK);
    a.group_kb = (uint32_t)(qw->group / WL_TILE_K);
    const uint32_t tiles = a.n_tiles * a.m_tiles;
    a.units = (uint64_t)tiles * a.k_blocks;
    a.y_f32 = reinterpret_cast<float *>(y); a.y_bf16 = nullptr;
    a.partial = reinterpret_cast<float *>(ctx->lin_flags.p);           // (int8 mode: the row sums)
    a.dbg = 0; a.trace = nullptr;
    if (dq) {
        a.y_f32 = dq->y_f32; a.y_bf16 = (__nv_bfloat16 *)dq->y_bf16;
        a.partial = reinterpret_cast<float *>(const_cast<int32_t *>(dq->rowsum));
        a.trace = reinterpret_cast<long long *>(const_cast<float *>(dq->rowscale));
        a.dbg = 256;
    }
    a.stream_k = 0;                                      // whole tiles only: int32 accumulators never leave TMEM half-summed
    const uint32_t sms = (uint32_t)ctx->sm_count;
    const uint32_t grid = tiles < sms ? tiles : sms;
    DLLM_TRY(ensure_smem_attr(ctx, umma_qlinear_x_kernel<CB, NTOK, KBS, NDQ, true>, C::kTotal));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->prof_on) {
        while (ctx->prof_ev.size() < 2 * (ctx->prof_n + 1)) {
            cudaEvent_t e;
            CUDA_TRY(ctx, cudaEventCreate(&e));
            ctx->prof_ev.push_back(e);
        }
        ev0 = ctx->prof_ev[2 * ctx->prof_n];
        ev1 = ctx->prof_ev[2 * ctx->prof_n + 1];
        CUDA_TRY(ctx, cudaEventRecord(ev0, ctx->stream));
    }
    umma_qlinear_x_kernel<CB, NTOK, KBS, NDQ, true><<<grid, (8 + 4 * NDQ) * 32, C::kTotal, ctx->stream>>>(tmap, a);
    LAUNCH_CHECK(ctx);
    if (ev1) {
        CUDA_TRY(ctx, cudaEventRecord(ev1, ctx->stream));
        ctx->prof_n++;
        ctx->prof_flops += 2.0 * (double)M * (double)qw->K * (double)qw->N;
        ctx->prof_bytes += (double)qw->K * qw->N * qw->bits / 8.0 + 1.0 * M * qw->K + 4.0 * M * qw->N;
    }
    return DLLM_OK;
}

// ---- int8 variant of the 256-token CTA-pair kernel (the int8 denoise mode's dense linears) ----
static bool pair2_i8_applicable(const dllm_ctx *ctx, const dllm_qweight *qw, size_t M, const I8Deq *dq) {
    static const bool off = getenv("DLLM_I8_PAIR") && atoi(getenv("DLLM_I8_PAIR")) == 0;       // experiments: 0 = 1-CTA kernel only
    if (off || !dq || qw->K % WL_TILE_K != 0 || M < 1024 || wl_container_bits(qw->bits) == 8) return false;
    if (qw->K % 16 != 0) return false;                                         // row pitch of the int8 activations (TMA)
    if (!dq->y_f32 && (qw->N % 8 != 0 || (reinterpret_cast<uintptr_t>(dq->y_bf16) & 15u) != 0)) return false;
    const uint32_t n_pairs = (uint32_t)((qw->n_tiles + 1) / 2);
    const uint32_t pairs_hw = (uint32_t)ctx->sm_count / 2;
    const uint32_t ntok = pair2_pick_ntok(M, n_pairs, pairs_hw, DLLM_I8_EPI16 ? 64 : 32);
    return (uint64_t)((M + ntok - 1) / ntok) * n_pairs >= pairs_hw / 2;
}

template <int CB>
int32_t launch_umma_pair2_i8(dllm_ctx *ctx, const dllm_qweight *qw, const int8_t *xq, size_t M, const I8Deq *dq) {
    using C = CfgP2<CB, true>;
#ifndef DLLM_I8_NDQ
#define DLLM_I8_NDQ DLLM_NDQ
#endif
    constexpr int NDQ = DLLM_I8_NDQ;          // unpack groups of 4 warps
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    Pair2Args a;
    memset(&a, 0, sizeof(a));
    a.packed = qw->d_packed; a.dqparams = qw->d_dqparams; a.bias = qw->d_bias;
    a.y_f32 = dq->y_f32; a.y_bf16 = (__nv_bfloat16 *)dq->y_bf16;
    a.M = (uint32_t)M; a.N = (uint32_t)qw->N; a.Npad = (uint32_t)(qw->n_tiles * 128);
    a.k_blocks = (uint32_t)qw->k_blocks; a.n_tiles = (uint32_t)qw->n_tiles; a.n_pairs = (a.n_tiles + 1) / 2;
    a.group_kb = (uint32_t)(qw->group / WL_TILE_K);
    const uint32_t pairs_hw = (uint32_t)ctx->sm_count / 2;
    const char *ntok_s = getenv("DLLM_UMMA_NTOK2");                                                 // experiments only
    const int ntok_env = ntok_s ? atoi(ntok_s) : 0;
    // (16 epilogue warps: each drains q / 2 columns, a multiple of 8 -> token tiles in multiples of 64)
    constexpr uint32_t kStep = C::kEpiWarps == 16 ? 64 : 32;
    a.ntok = (ntok_env >= 64 && ntok_env <= 256 && ntok_env % kStep == 0) ? (uint32_t)ntok_env : pair2_pick_ntok(M, a.n_pairs, pairs_hw, kStep);
    a.rs_rows = 1; a.gate_rows = 1; a.gate_sub = 1;
    a.m_tiles = (uint32_t)((M + a.ntok - 1) / a.ntok);
    a.tiles = a.n_pairs * a.m_tiles;
    a.i8_rowsum = dq->rowsum; a.i8_rowscale = dq->rowscale; a.i8_zp = (int32_t)qw->tensor_zp;
    static const uint32_t dbg_flags = getenv("DLLM_UMMA_DBG") ? (uint32_t)atoi(getenv("DLLM_UMMA_DBG")) : 0u;
    a.dbg = dbg_flags & (1u | 2u | 8u | 64u | 128u);         // timing experiments only: 1 skip MMAs, 2 skip the unpack, 8 skip the activation loads, 64 skip stores, 128 stage timeline
    if (a.dbg & 128) {
        DLLM_TRY(ensure_buf(ctx, ctx->lin_ws, 8 * 256 * sizeof(long long)));
        a.trace = (long long *)ctx->lin_ws.p;
        cudaMemsetAsync(a.trace, 0, 8 * 256 * sizeof(long long), ctx->stream);
    }

    // int8 activations [M tokens, K bytes] row-major; one box = 128 k x ntok/2 tokens in the SWIZZLE_128B K-major layout
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)qw->K, (cuuint64_t)M};
    const cuuint64_t gstride[1] = {(cuuint64_t)qw->K};
    const cuuint32_t box[2] = {2 * WL_TILE_K, a.ntok / 2};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(xq), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    CUtensorMap tmap_y = tmap;
    if (!dq->y_f32) {
        const cuuint64_t ydim[2] = {(cuuint64_t)qw->N, (cuuint64_t)M};
        const cuuint64_t ystride[1] = {(cuuint64_t)qw->N * 2};
        const cuuint32_t ybox[2] = {128, a.ntok / 4};
        const cuuint32_t yestr[2] = {1, 1};
        r = enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dq->y_bf16, ydim, ystride, ybox, yestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (%d)", (int)r);
    }
    RsMaps rsm;
    memset(&rsm, 0, sizeof(rsm));
    DLLM_TRY(ensure_smem_attr(ctx, umma_qlinear_pair2_kernel<CB, NDQ, true>, C::kTotal));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->prof_on) {
        while (ctx->prof_ev.size() < 2 * (ctx->prof_n + 1)) {
            cudaEvent_t e;
            CUDA_TRY(ctx, cudaEventCreate(&e));
            ctx->prof_ev.push_back(e);
        }
        ev0 = ctx->prof_ev[2 * ctx->prof_n];
        ev1 = ctx->prof_ev[2 * ctx->prof_n + 1];
        CUDA_TRY(ctx, cudaEventRecord(ev0, ctx->stream));
    }
    const uint32_t pairs = a.tiles < pairs_hw ? a.tiles : pairs_hw;
    static const bool no_pdl = getenv("DLLM_UMMA_NO_PDL") != nullptr;      // experiments only
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3((4 + 4 * NDQ + C::kEpiWarps) * 32);
    cfg.dynamicSmemBytes = C::kTotal;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, umma_qlinear_pair2_kernel<CB, NDQ, true>, tmap, tmap_y, rsm, a));
    LAUNCH_CHECK(ctx);
    if (a.dbg & 128) {   // dump the timeline of cluster 0's leader (timing experiments only)
        std::vector<long long> hst(8 * 256);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(hst.data(), a.trace, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        static int n_dump = 0;
        char name[64];
        snprintf(name, sizeof(name), "gpurun_out/pair2_i8_trace_%d.csv", n_dump++ & 3);
        FILE *f = fopen(name, "w");
        if (f) {
            fprintf(f, "i,dq_start(stage),mma_h0(stage),mma_h1(stage),dq_end(stage),epi_h0_ready(tile),epi_h1_ready(tile),epi_h0_released(tile),epi_h1_released(tile)\n");
            for (int i = 0; i < 256; ++i) {
                fprintf(f, "%d", i);
                for (int r = 0; r < 8; ++r) fprintf(f, ",%lld", hst[r * 256 + i]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    if (ev1) {
        CUDA_TRY(ctx, cudaEventRecord(ev1, ctx->stream));
        ctx->prof_n++;
        ctx->prof_flops += 2.0 * (double)M * (double)qw->K * (double)qw->N;
        ctx->prof_bytes += (double)qw->K * qw->N * qw->bits / 8.0 + 1.0 * M * qw->K + (dq->y_f32 ? 4.0 : 2.0) * M * qw->N;
    }
    return DLLM_OK;
}

template <int CB>
int32_t launch_umma_i8_ntok(dllm_ctx *ctx, const dllm_qweight *qw, const int8_t *xq, size_t M, int32_t *y, const I8Deq *dq = nullptr) {
    if constexpr (CB != 8) {
        if (pair2_i8_applicable(ctx, qw, M, dq)) return launch_umma_pair2_i8<CB>(ctx, qw, xq, M, dq);
    }
    if (M <= 16) return launch_umma_i8<CB, 16>(ctx, qw, xq, M, y, dq);
    if (M <= 32) return launch_umma_i8<CB, 32>(ctx, qw, xq, M, y, dq);
    if (M <= 64) return launch_umma_i8<CB, 64>(ctx, qw, xq, M, y, dq);
    return launch_umma_i8<CB, 128>(ctx, qw, xq, M, y, dq);
}

}  // namespace

// exact integer linear on the tcgen05 int8 path: per-tensor quantized weights (one integer zero-point), K % 64 == 0 (whole
// k-blocks), K * 255 * 128 < 2^31 (no int32 overflow)
bool k_umma_i8_supported(const dllm_qweight *qw, size_t M) {
    return qw && M >= 1 && M < (1u << 31) && qw->per_tensor && qw->int_zps && qw->K % WL_TILE_K == 0 && qw->K <= 65536;
}

int32_t k_qlinear_umma_i8(dllm_ctx *ctx, const dllm_qweight *qw, const int8_t *xq_dev, size_t M, int32_t *y_i32_dev) {
    if (!k_umma_i8_supported(qw, M)) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "int8 path: per-tensor quantized weight with K %% 64 == 0 and K <= 65536 required");
    if ((reinterpret_cast<uintptr_t>(xq_dev) & 15u) != 0) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "xq must be 16-byte aligned");
    switch (wl_container_bits(qw->bits)) {
        case 2: return launch_umma_i8_ntok<2>(ctx, qw, xq_dev, M, y_i32_dev);
        case 4: return launch_umma_i8_ntok<4>(ctx, qw, xq_dev, M, y_i32_dev);
        default: return launch_umma_i8_ntok<8>(ctx, qw, xq_dev, M, y_i32_dev);
    }
}

// Row-parallel linear with the reduce-scatter fused into its epilogue (tp.cu): the dense CTA-pair kernel with 256-token tiles,
// bf16 output, every rank's token slice a multiple of the 64-row store block
bool k_umma_rs_supported(const dllm_ctx *ctx, const dllm_qweight *qw, size_t M, int world) {
    if (!qw || world < 2 || world > 8 || M % (size_t)world != 0 || (M / world) % 64 != 0 || qw->N % 8 != 0) return false;
    if (wl_container_bits(qw->bits) == 8 || !(qw->K % 8 == 0 && qw->group % WL_TILE_K == 0 && qw->int_zps)) return false;
    if (pair2_mode() != 2 || qw->K % WL_TILE_K != 0 || M < 1024) return false;
    const uint32_t n_pairs = (uint32_t)((qw->n_tiles + 1) / 2);
    const uint32_t pairs_hw = (uint32_t)(ctx->sm_limit > 0 && ctx->sm_limit < ctx->sm_count ? ctx->sm_limit : ctx->sm_count) / 2;
    return (uint64_t)((M + 255) / 256) * n_pairs >= pairs_hw / 2;
}

// a GEMM whose activation rows of the other ranks' token slices may still be arriving (flag-gated loads): the dense CTA-pair
// kernel with bf16 output and whole 256-token tiles per slice
bool k_umma_gate_supported(const dllm_ctx *ctx, const dllm_qweight *qw, size_t M, int world, const void *y_bf16) {
    if (!qw || world < 2 || M % (size_t)world != 0 || (M / world) % 256 != 0 || wl_container_bits(qw->bits) == 8) return false;
    if (!(qw->K % 8 == 0 && qw->group % WL_TILE_K == 0 && qw->int_zps)) return false;
    return pair2_applicable(ctx, qw, M, nullptr, y_bf16);
}

int32_t k_qlinear_umma_rs(dllm_ctx *ctx, const dllm_qweight *qw, const void *x_bf16_dev, size_t M, const UmmaRs *rs) {
    if (!rs || !k_umma_rs_supported(ctx, qw, M, rs->world) || rs->rows * (size_t)rs->world != M)
        DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "fused reduce-scatter: unsupported shape");
    if ((reinterpret_cast<uintptr_t>(x_bf16_dev) & 15u) != 0) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "x must be 16-byte aligned");
    for (int r = 0; r < rs->world; ++r)
        if (!rs->recv[r] || (reinterpret_cast<uintptr_t>(rs->recv[r]) & 15u) != 0) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "receive buffers must be 16-byte aligned");
    // (y_bf16 = any aligned non-null pointer: nothing is stored through it in this mode)
    if (wl_container_bits(qw->bits) == 2) return launch_umma_pair2<2>(ctx, qw, x_bf16_dev, M, nullptr, rs->recv[rs->rank], rs);
    return launch_umma_pair2<4>(ctx, qw, x_bf16_dev, M, nullptr, rs->recv[rs->rank], rs);
}

// the same exact integer linear with the dequantization fused into its epilogue: y = (sum - zp rowsum[m]) * rowscale[m] + bias[n]
// (rowscale = the weight's scale x the token's activation step; rowsum = the token's int8 row sum): the int8 denoise stack
int32_t k_qlinear_umma_i8_deq(dllm_ctx *ctx, const dllm_qweight *qw, const int8_t *xq_dev, const int32_t *rowsum_dev,
                              const float *rowscale_dev, size_t M, float *y_f32_dev, void *y_bf16_dev) {
    if (!k_umma_i8_supported(qw, M)) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "int8 path: per-tensor quantized weight with K %% 64 == 0 and K <= 65536 required");
    if ((reinterpret_cast<uintptr_t>(xq_dev) & 15u) != 0) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "xq must be 16-byte aligned");
    const I8Deq dq = {rowsum_dev, rowscale_dev, y_f32_dev, y_bf16_dev};
    switch (wl_container_bits(qw->bits)) {
        case 2: return launch_umma_i8_ntok<2>(ctx, qw, xq_dev, M, nullptr, &dq);
        case 4: return launch_umma_i8_ntok<4>(ctx, qw, xq_dev, M, nullptr, &dq);
        default: return launch_umma_i8_ntok<8>(ctx, qw, xq_dev, M, nullptr, &dq);
    }
}

bool k_umma_supported(const dllm_qweight *qw, size_t M) {
    // TMA needs a 16-byte row pitch for x (K % 8 == 0); everything else is padded / masked
    return qw && M >= 1 && M < (1u << 31) && qw->K % 8 == 0 && qw->group % WL_TILE_K == 0 && qw->int_zps;
}

int32_t k_qlinear_umma(dllm_ctx *ctx, const dllm_qweight *qw, const void *x_bf16_dev, size_t M, float *y_f32_dev,
                       void *y_bf16_dev) {
    if (!k_umma_supported(qw, M)) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "tcgen05 path: unsupported shape (K %% 8 != 0)");
    if ((reinterpret_cast<uintptr_t>(x_bf16_dev) & 15u) != 0) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "x must be 16-byte aligned");
    switch (wl_container_bits(qw->bits)) {
        case 2: return launch_umma_ntok<2>(ctx, qw, x_bf16_dev, M, y_f32_dev, y_bf16_dev);
        case 4: return launch_umma_ntok<4>(ctx, qw, x_bf16_dev, M, y_f32_dev, y_bf16_dev);
        default: return launch_umma_ntok<8>(ctx, qw, x_bf16_dev, M, y_f32_dev, y_bf16_dev);
    }
}
