// gemv_mma.cu — HBM-bound dequant-GEMV / skinny GEMM for 1..16 tokens (K4 fast path).
//
//   y[M,N] = x[M,K] · dequant(W) + b,  M <= 16     (diffuse-llm-rs/src/lib.rs:812 composed with
//                                                    dequantize_tensor, quantization.rs:81-85)
//
// At M <= 16 the op moves K·N·bits/8 bytes of codes for 2·M·K·N flops: it is bound by how fast the
// packed weights stream out of HBM, so the kernel is organised around the copy engine, not the math:
//
//   * one persistent CTA per SM; the k-blocks of all tiles are dealt out in equal contiguous ranges
//     (k-segment-major stream-K), so every SM streams the same number of bytes;
//   * a producer warp keeps a deep shared-memory ring (up to 32 stages, > 100 KB in flight per SM) full
//     with `cp.async.bulk` copies — one 2/4/8 KB packed tile (wlayout.cuh: contiguous in HBM) plus the
//     tile's 128 scales and zero-points per stage, completion on an mbarrier;
//   * the CTA's slice of the activations is staged ONCE into shared memory (bf16, fragment order) and
//     stays resident: no activation traffic in the steady state;
//   * NG groups of 8 consumer warps take the stages round-robin.  A warp owns 16 output columns of the
//     128-column tile: it reads its codes with conflict-free LDS.32 and turns them into fp16 operands in
//     registers with ONE LOP3 per pair — 0x6400 | q is the half 1024 + q, 0x6400 | (q << 4) is 1024 + 16 q (those
//     k positions meet activations prepared as x / 16) — so a 32-bit word of eight 4-bit codes costs one shift
//     and four LOP3, and feeds them to mma.sync.m16n8k16 (f16 operands, f32 accumulate) with the activations as
//     the B operand.  Neither the magic 1024 nor the zero-point is subtracted per weight: per (k-block, token)
//     the x preparation also stores T = sum of the fp16 operands and S = sum of x, split into fp16 parts, and ONE
//     more MMA per k-block with A = {64 zp, 32768, ...} and B = {-S/64, -T/64, ...} removes 1024 T + zp S from the
//     accumulator (sum x (1024 + q) - 1024 sum x - zp sum x = sum x (q - zp)).  The group's scale is applied in f32
//     to the k-block's partial sum.  Tensor cores are used only so that the FMA work costs one instruction per 256
//     weights: what bounds the kernel after the copy engine is the CUDA-core instructions of the unpack;
//   * tiles cut by a range boundary are reduced by the LAST CTA to arrive at the tile (atomic ticket,
//     no spinning), always in CTA order: results are deterministic and there is no fix-up launch.
//
// Numerics: codes and zero-points exact, x rounded to fp16 (saturated at +-65504), f32 accumulate, f32 scale; the
// 1024-offset cancels inside the f32 accumulator of one k-block (|1024 sum_64 x| * 2^-24 per MMA: two orders of
// magnitude below the fp16 rounding of x).  Zero-points must be the integers quantizer B produces (quantization.rs:55-56).
// Timeline instrumentation (globaltimer stamps per CTA / per stage) compiles in with -DDLLM_GEMV_TRACE.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "wlayout.cuh"

#ifndef DLLM_GEMV_EXP
#define DLLM_GEMV_EXP 0
#endif

namespace {

constexpr int kGroupWarps = 8;                 // 8 warps x 16 output columns = one 128-column tile
constexpr int kRedStride = 132;                // padded row of the cross-group reduction buffer (floats)
constexpr int kSmemBudget = 220 * 1024;      // of the 227 KB a CTA may use
constexpr int kXBudget = 112 * 1024;           // resident activation slice per CTA
constexpr int kMaxStages = 32;
constexpr int kMaxContrib = 160;              // a CTA contributes at most once to a tile, so grid <= kMaxContrib suffices

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// producer-side wait: the thread may stay suspended for up to `ns` before the try_wait returns false — a producer
// that polls a barrier in a tight loop steals issue slots from the consumer warps (measured: 27% of all issued
// instructions), and it is never latency-critical: it runs a whole ring ahead
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "GR_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra GR_DONE;\n\t"
        "bra GR_LOOP;\n\t"
        "GR_DONE:\n\t"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// same wait, returning a zero the compiler cannot see through: adding it to the shared-memory addresses of the
// loads that follow makes them data-dependent on the wait (they are plain asm loads, free to be scheduled otherwise)
__device__ __forceinline__ uint32_t mbar_wait_token(uint32_t bar, uint32_t parity) {
    uint32_t z;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "GT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "@p bra GT_DONE;\n\t"
        "bra GT_LOOP;\n\t"
        "GT_DONE:\n\t"
        "mov.u32 %0, 0;\n\t"
        "}\n" : "=r"(z) : "r"(bar), "r"(parity) : "memory");
    return z;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 16-byte asynchronous copy (LDGSTS) whose completion is reported to an mbarrier by cp_async_arrive
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// programmatic dependent launch: `launch_dependents` lets the next kernel of the stream start its prologue while this
// one runs; `wait` blocks until the previous kernel of the stream has completed and its memory is visible
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
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
// named barriers: kBarConsumers — among the consumer warps; kBarPartial — consumers arrive once a partial tile is
// in global memory, the epilogue warp waits for it; kBarFinal — the epilogue warp arrives with its verdict on the
// CTA's last tile, the consumers wait for it (they have nothing else left to do and help with that reduction)
// kBarEpiFree — the epilogue warp is ready for the next partial tile (a named barrier must not collect the arrivals of
// two tiles at once, so the consumers wait for it before they arrive on kBarPartial again)
constexpr int kBarConsumers = 1, kBarPartial = 2, kBarFinal = 3, kBarEpiFree = 4, kBarXReady = 5;
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(threads) : "memory"); }
// D(16 columns x 8 tokens, f32) += A(16 columns x 16 k, f16) · B(16 k x 8 tokens, f16)
__device__ __forceinline__ void mma_f16(float *d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// (a & mask) | magic in ONE LOP3 (with two immediates the compiler emits two)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t mask, uint32_t magic) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(magic));
    return d;
}
// k (inside a 64-k block) of element e (0..7) of the 16-byte activation unit that lane quad-index t
// consumes in round u (0..1); the code words of wlayout.cuh decide it (see consume_stage)
template <int CB>
__host__ __device__ __forceinline__ int gemv_kmap(int u, int t, int e) {
    if (CB == 4) return 32 * u + 8 * t + e;
    if (CB == 2) return 16 * t + 8 * u + e;
    return 32 * u + 16 * (e >> 2) + 4 * t + (e & 3);
}

struct GemvArgs {
    const uint8_t *packed;
    const uint2 *gparams;            // [G][Npad] {f32 scale, half2(64 zp, zp / 32)}
    const float *bias;
    const uint8_t *xb;               // !XR: prepared activations [k_blocks] tiles of gemv_x_tile_bytes(MT)
    const float *x;                  // XR: the f32 activations [M, K]; every CTA prepares its own slice in shared memory
    uint32_t K;
    float *y;                        // [M, N]
    float *partial;                  // [grid * max_items][MT][128]
    unsigned int *tickets;           // [n_tiles], zero between launches
    uint32_t M, N, Npad, k_blocks, n_tiles;
    uint32_t group_magic;            // ceil(2^32 / k-blocks per quantization group): kb / group_kb == umulhi(kb, magic) for kb < 2^16 (0: identity)
    uint32_t S, P;                   // k segments, CTAs per segment (grid = S * P)
    uint32_t max_items;              // partial slots per CTA
    uint32_t stages;                 // ring depth (multiple of NG)
    uint32_t x_off, red_off, bar_off;   // shared-memory carve-up (bytes)
    uint32_t bulk;                   // code tiles by one cp.async.bulk per stage (default) or by per-lane cp.async (DLLM_GEMV_BULK=0)
    unsigned long long *trace;       // -DDLLM_GEMV_TRACE: per CTA [32] globaltimer stamps, then [4][256] stage stamps of CTA 0
};

// the k-segment and unit range of one CTA; units of a segment are ordered (tile, k-block)
struct Range {
    uint32_t kb_s0, kbs;             // segment = k-blocks [kb_s0, kb_s0 + kbs)
    uint64_t units, u0, u1;          // units of the segment; this CTA's range
    __device__ Range(const GemvArgs &a, uint32_t cta) {
        const uint32_t seg = cta / a.P, j = cta - seg * a.P;
        kb_s0 = (uint32_t)((uint64_t)a.k_blocks * seg / a.S);
        kbs = (uint32_t)((uint64_t)a.k_blocks * (seg + 1) / a.S) - kb_s0;
        units = (uint64_t)a.n_tiles * kbs;
        u0 = units * j / a.P;
        u1 = units * (j + 1) / a.P;
    }
};

struct Item {
    uint32_t nt, kb0, kb1, ordinal;
};

struct ItemIter {
    uint64_t u, u1;
    uint32_t kb_s0, kbs, first_nt;
    __device__ ItemIter(const Range &r) : u(r.u0), u1(r.u1), kb_s0(r.kb_s0), kbs(r.kbs) { first_nt = kbs ? (uint32_t)(r.u0 / kbs) : 0; }
    __device__ bool next(Item &it) {
        if (u >= u1) return false;
        it.nt = (uint32_t)(u / kbs);
        const uint32_t off = (uint32_t)(u - (uint64_t)it.nt * kbs);
        const uint64_t left = u1 - u;
        const uint32_t len = (uint64_t)(kbs - off) <= left ? kbs - off : (uint32_t)left;
        it.kb0 = kb_s0 + off;
        it.kb1 = it.kb0 + len;
        it.ordinal = it.nt - first_nt;
        u += len;
        return true;
    }
};

// bytes of one k-block of prepared activations: [2 rounds][MT tokens][4 t] 16-byte units, then one 16-byte
// correction entry per token (x_corr_entry)
__host__ __device__ constexpr int gemv_x_tile_bytes(int MT) { return MT * 144; }

// one 64-k block of a 128-column tile, consumed by the 8 warps of one group.
// Shared-memory addresses, each already offset to this lane's element (lane = 4 g + t, output columns
// r0 = 16 w + g and r0 + 8):
//   cw : word t of chunk 0 of column r0 of the packed codes (wlayout.cuh): + 2048 per chunk, + 128 for column r0 + 8
//   pw : {f32 scale, half2(64 zp, zp / 32)} of column r0: + 64 for column r0 + 8
//   xw : the lane's 16-byte unit of token min(g, MT - 1) in round 0 of the fp16 activations: + 64 MT per round, + 512 for
//        token g + 8.  (Lanes with g >= MT feed MMA columns of tokens that do not exist and are never stored; they
//        re-read the last token instead of zeroing registers.)
//   xc : word t of that token's correction entry: + 128 for token g + 8
template <int CB, int MT>
__device__ __forceinline__ void consume_kblock(uint32_t cw, uint32_t pw, uint32_t xw, uint32_t xc, uint32_t zmask, uint32_t cfill, float (*ya)[4]) {
    constexpr int NB = MT > 8 ? 2 : 1;           // 8-token MMA column blocks
    constexpr uint32_t kMagic = 0x64006400u;     // half2(1024, 1024)
    const uint2 p0 = lds64(pw), p1 = lds64(pw + 64);
    float d[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) d[nb][0] = d[nb][1] = d[nb][2] = d[nb][3] = 0.f;
#if DLLM_GEMV_EXP == 2      // timing experiment: two independent accumulation chains
    float d2[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) d2[nb][0] = d2[nb][1] = d2[nb][2] = d2[nb][3] = 0.f;
#define DLLM_EXP_ACC(u, nb) ((u) ? d2[nb] : d[nb])
#else
#define DLLM_EXP_ACC(u, nb) d[nb]
#endif

    // activations of both rounds first: independent of everything else
    uint4 b[2][NB];
    uint32_t bc[NB];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            b[u][nb] = lds128(xw + u * (MT * 64) + nb * 512);
        }
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) bc[nb] = lds32(xc + nb * 128);

    if (CB == 4) {
        // word = codes 0..7 at nibbles {0,4,1,5,2,6,3,7}: (w & 0x000f000f) | magic = 1024 + codes (0,1),
        // (w & 0x00f000f0) | magic = 1024 + 16 x codes (2,3) (their activations are stored as x / 16), the same of
        // w >> 8 = codes (4,5), (6,7)
        uint32_t q[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) { q[u][0] = lds32(cw + u * 2048); q[u][1] = lds32(cw + u * 2048 + 128); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t v0 = h ? q[u][0] >> 8 : q[u][0], v1 = h ? q[u][1] >> 8 : q[u][1];
#if DLLM_GEMV_EXP == 4      // timing experiment: no unpack
                const uint32_t a0 = v0, a1 = v1, a2 = v0 + 1, a3 = v1 + 1;
#else
                const uint32_t a0 = and_or(v0, 0x000f000fu, kMagic);
                const uint32_t a1 = and_or(v1, 0x000f000fu, kMagic);
                const uint32_t a2 = and_or(v0, 0x00f000f0u, kMagic);
                const uint32_t a3 = and_or(v1, 0x00f000f0u, kMagic);
#endif
#if DLLM_GEMV_EXP == 3      // timing experiment: half of the data MMAs (operands still computed)
                if (h) { asm volatile("" :: "r"(a0), "r"(a1), "r"(a2), "r"(a3)); continue; }
#endif
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) mma_f16(DLLM_EXP_ACC(u, nb), a0, a1, a2, a3, h ? b[u][nb].z : b[u][nb].x, h ? b[u][nb].w : b[u][nb].y);
            }
        }
    } else if (CB == 2) {
        // one word = 16 codes, code i at field (i >> 1) + 8 (i & 1): (w >> 2p) & 0x00030003 = codes (2p, 2p + 1)
        const uint32_t q0 = lds32(cw), q1 = lds32(cw + 128);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = 4 * u + 2 * h;
                const uint32_t a0 = and_or(q0 >> (2 * p), 0x00030003u, kMagic);
                const uint32_t a1 = and_or(q1 >> (2 * p), 0x00030003u, kMagic);
                const uint32_t a2 = and_or(q0 >> (2 * p + 2), 0x00030003u, kMagic);
                const uint32_t a3 = and_or(q1 >> (2 * p + 2), 0x00030003u, kMagic);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) mma_f16(d[nb], a0, a1, a2, a3, h ? b[u][nb].z : b[u][nb].x, h ? b[u][nb].w : b[u][nb].y);
            }
        }
    } else {
        // 16-k chunks, word t = 4 codes in byte order; PRMT with 0x64 bytes builds half2(1024 + q_i, 1024 + q_j)
        uint32_t q[4][2];
#pragma unroll
        for (int c = 0; c < 4; ++c) { q[c][0] = lds32(cw + c * 2048); q[c][1] = lds32(cw + c * 2048 + 128); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = 2 * u + h;
                const uint32_t a0 = __byte_perm(q[c][0], kMagic, 0x5150);
                const uint32_t a1 = __byte_perm(q[c][1], kMagic, 0x5150);
                const uint32_t a2 = __byte_perm(q[c][0], kMagic, 0x5352);
                const uint32_t a3 = __byte_perm(q[c][1], kMagic, 0x5352);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) mma_f16(d[nb], a0, a1, a2, a3, h ? b[u][nb].z : b[u][nb].x, h ? b[u][nb].w : b[u][nb].y);
            }
        }
    }
    // the correction MMA: k slots (0,1) = (64 zp, zp / 32) x the fp16 parts of -S/64, slots 2..5 = (32768, 32768, 32, 2^-6) x
    // the parts of -T/64 (x_corr_entry), the other slots zero: d -= zp S + 1024 T, i.e. d = sum x (q - zp)
    // (quantization.rs:83's `- zp`)
#if DLLM_GEMV_EXP != 1   // (1: timing experiment without the correction MMA — wrong results)
    {
        const uint32_t ca0 = and_or(p0.y, zmask, cfill), ca1 = and_or(p1.y, zmask, cfill);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) mma_f16(d[nb], ca0, ca1, 0u, 0u, bc[nb], 0u);
    }
#endif
#if DLLM_GEMV_EXP == 2
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) { d[nb][0] += d2[nb][0]; d[nb][1] += d2[nb][1]; d[nb][2] += d2[nb][2]; d[nb][3] += d2[nb][3]; }
#endif
    // dequantize_tensor's `* scale` (quantization.rs:83), applied to the k-block's partial sum in f32
    const float s0 = __uint_as_float(p0.x), s1 = __uint_as_float(p1.x);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        ya[nb][0] = fmaf(s0, d[nb][0], ya[nb][0]);
        ya[nb][1] = fmaf(s0, d[nb][1], ya[nb][1]);
        ya[nb][2] = fmaf(s1, d[nb][2], ya[nb][2]);
        ya[nb][3] = fmaf(s1, d[nb][3], ya[nb][3]);
    }
}

// The correction entry of one (k-block, token): the B operand of consume_kblock's correction MMA, word t for lane
// quad-index t.  s_all = sum of the 64 activations as the weights see them, s_fed = sum of the 64 fp16 operands as the
// MMAs see them (equal unless CB == 4, where half of them are fed as x / 16).  |s / 64| <= 65504 always.  An f32 is
// carried as fp16 parts that are rescaled to the magnitude of the leading part (hi, 2^11 (v - hi), 2^22 (v - hi - mid)),
// so none of them is an fp16 subnormal and power-of-two scalings of x commute with the whole path; the A operand holds
// the matching factors:   word 0: -S/64 as (hi, mid')   x (64 zp, zp / 32)
//                         word 1: -T/64 as (hi, hi)     x (32768, 32768)
//                         word 2: -T/64 as (mid', lo')  x (32, 2^-6)             word 3: zero
__device__ __forceinline__ uint4 x_corr_entry(float s_all, float s_fed) {
    const float vs = -s_all * 0.015625f, vt = -s_fed * 0.015625f;
    const __half shi = __float2half_rn(vs);
    const __half smid = __float2half_rn((vs - __half2float(shi)) * 2048.f);
    const __half thi = __float2half_rn(vt);
    const float r = (vt - __half2float(thi)) * 2048.f;
    const __half tmid = __float2half_rn(r);
    const __half tlo = __float2half_rn((r - __half2float(tmid)) * 2048.f);
    return make_uint4((uint32_t)__half_as_ushort(shi) | ((uint32_t)__half_as_ushort(smid) << 16),
                      (uint32_t)__half_as_ushort(thi) * 0x10001u,
                      (uint32_t)__half_as_ushort(tmid) | ((uint32_t)__half_as_ushort(tlo) << 16), 0u);
}

// One 16-byte unit of the prepared activations: the 8 fp16 values of token `tok` that lane quad-index t consumes in
// round u of k-block kb (gemv_kmap).  They are two runs of 4 consecutive k; `vec` = K % 4 == 0 and x is 16-byte aligned.
// The load and the conversion are separate so that a thread can have the loads of several units in flight.
template <int CB>
__device__ __forceinline__ void gemv_x_load(const float *__restrict__ x, uint32_t M, uint32_t K, bool vec, uint32_t kb, uint32_t u, uint32_t tok, uint32_t t, float *v) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t k = kb * WL_TILE_K + (uint32_t)gemv_kmap<CB>((int)u, (int)t, 4 * h);
        if (tok < M && vec && k + 3 < K) {
            const float4 f = __ldg(reinterpret_cast<const float4 *>(x + (size_t)tok * K + k));
            v[4 * h] = f.x; v[4 * h + 1] = f.y; v[4 * h + 2] = f.z; v[4 * h + 3] = f.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[4 * h + e] = (tok < M && k + e < K) ? __ldg(x + (size_t)tok * K + k + e) : 0.f;
        }
    }
}
// CB == 4: elements 2, 3, 6, 7 meet weights unpacked as 1024 + 16 q and are stored as x / 16 (exact but for fp16
// subnormals).  s_all / s_fed are the unit's share of x_corr_entry's sums, from the ROUNDED operands.
template <int CB>
__device__ __forceinline__ uint4 gemv_x_convert(const float *v, float &s_all, float &s_fed) {
    uint32_t o[4];
    float sa = 0.f, sf = 0.f;
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
        const bool sixteenth = CB == 4 && (e2 & 1);
        const float pre = sixteenth ? 0.0625f : 1.0f;
        const __half2 h2 = __floats2half2_rn(fminf(fmaxf(v[2 * e2], -65504.f), 65504.f) * pre, fminf(fmaxf(v[2 * e2 + 1], -65504.f), 65504.f) * pre);
        const float2 f2 = __half22float2(h2);
        const float pair = f2.x + f2.y;
        sf += pair;
        sa += sixteenth ? 16.f * pair : pair;
        o[e2] = *reinterpret_cast<const uint32_t *>(&h2);
    }
    s_all = sa; s_fed = sf;
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// the units of the k-blocks [kb0, kb0 + n_kb) into `dst` (tile of k-block kb0 first), all threads of whole warps:
// 8 consecutive lanes hold the 8 units (u, t) of one (k-block, token), so its sums are three shuffles away.
// UNR units per thread and pass, all loads issued before the first conversion: the pass costs one L2 round trip.
template <int CB, int MT, int UNR>
__device__ __forceinline__ void prepare_x_tiles(const float *__restrict__ x, uint32_t M, uint32_t K, uint32_t kb0, uint32_t n_kb,
                                                uint8_t *dst, uint32_t tid, uint32_t nthreads) {
    const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const uint32_t total = n_kb * MT * 8;
    for (uint32_t base = tid & ~31u; base < total; base += UNR * nthreads) {
        float v[UNR][8];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const uint32_t idx = base + j * nthreads + (tid & 31);
            if (idx < total) {
                gemv_x_load<CB>(x, M, K, vec, kb0 + (idx >> 3) / MT, (idx >> 2) & 1, (idx >> 3) % MT, idx & 3, v[j]);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const uint32_t idx = base + j * nthreads + (tid & 31);
            const bool valid = idx < total;                               // total % 8 == 0: a group of 8 lanes is all valid or all not
            const uint32_t t = idx & 3, u = (idx >> 2) & 1, tok = (idx >> 3) % MT, kb = (idx >> 3) / MT;
            float sa, sf;
            const uint4 unit = gemv_x_convert<CB>(v[j], sa, sf);
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, m);
                sf += __shfl_xor_sync(0xffffffffu, sf, m);
            }
            if (valid) {
                uint8_t *tile = dst + (size_t)kb * gemv_x_tile_bytes(MT);
                *reinterpret_cast<uint4 *>(tile + ((u * MT + tok) * 4 + t) * 16) = unit;
                if ((idx & 7) == 0) *reinterpret_cast<uint4 *>(tile + MT * 128 + tok * 16) = x_corr_entry(sa, sf);
            }
        }
    }
}

// y tile = sum of the partial tiles of all contributors (fixed order) + bias; `nthreads` threads, this one is `tid`
template <int MT>
__device__ __forceinline__ void reduce_tile(const GemvArgs &a, uint32_t cnt, const uint32_t *slots, uint32_t nt, int tid, int nthreads) {
    for (int e4 = tid; e4 < MT * 32; e4 += nthreads) {                 // 4 consecutive columns of one token
        const int tok = e4 >> 5, nl = (e4 & 31) * 4;
        const uint32_t n = nt * 128 + nl;
        if ((uint32_t)tok >= a.M || n >= a.N) continue;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (uint32_t c = 0; c < cnt; ++c) {
            const float4 pv = __ldcg(reinterpret_cast<const float4 *>(a.partial + (size_t)slots[c] * (MT * 128)) + e4);
            v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
        }
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (n + i < a.N) a.y[(size_t)tok * a.N + n + i] = vv[i] + (a.bias ? __ldg(a.bias + n + i) : 0.f);
    }
}

// XR: consumers + epilogue warp convert the CTA's k-segment of x to fp16 in shared memory (the producers are
// already streaming weights meanwhile), then meet at kBarXReady
template <int CB, int MT>
__device__ __forceinline__ void prepare_x_slice(const GemvArgs &a, const Range &rg, uint8_t *xs, int tid, int nthreads) {
    if (rg.u0 < rg.u1) prepare_x_tiles<CB, MT, 3>(a.x, a.M, a.K, rg.kb_s0, rg.kbs, xs, (uint32_t)tid, (uint32_t)nthreads);
    named_bar_sync(kBarXReady, nthreads);
}

// XR: the CTA's slice of the activations is resident in shared memory (small M·K); otherwise every stage carries the
// activations of its k-blocks (they come from L2: the prepared x is at most a few hundred KB)
template <int CB, int MT, int NG, int NP, int KBS, bool XR>
__global__ void __launch_bounds__((NG * kGroupWarps + NP + 1) * 32, 1)
gemv_mma_kernel(const GemvArgs a) {
    constexpr int NB = MT > 8 ? 2 : 1;
    constexpr int kWBytes = WL_TILE_N * WL_TILE_K * CB / 8;
    constexpr int kXTile = gemv_x_tile_bytes(MT);   // activations of one k-block
    // a stage: KBS code tiles, KBS x 128 {scale, zero-point} pairs, and (unless XR) KBS activation tiles
    constexpr int kStage = KBS * (kWBytes + 1024 + (XR ? 0 : kXTile));
    constexpr int kConsumers = NG * kGroupWarps * 32;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *ring = smem;
    uint8_t *xs = smem + a.x_off;
    float *red = reinterpret_cast<float *>(smem + a.red_off);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + a.bar_off);
    uint64_t *empty = full + a.stages;
    uint64_t *xfull = empty + a.stages;
    uint32_t *verdict = reinterpret_cast<uint32_t *>(xfull + 1);     // [0] this CTA arrived last at its final tile  [1] contributors
    uint32_t *slots = verdict + 2;                                   // partial slots of the contributors of one tile

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const Range rg(a, blockIdx.x);
#ifdef DLLM_GEMV_TRACE
#define STRACE(role, idx) do { if (blockIdx.x == 0 && (idx) < 256) { unsigned long long _t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(_t)); a.trace[gridDim.x * 32 + (role) * 256 + (idx)] = _t; } } while (0)
#define GTRACE(slot) do { if ((slot) < 32) { unsigned long long _t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(_t)); a.trace[blockIdx.x * 32 + (slot)] = _t; } } while (0)
#else
#define STRACE(role, idx) do { } while (0)
#define GTRACE(slot) do { } while (0)
#endif
    if (threadIdx.x == 0) GTRACE(0);
    // The next kernel of the stream (typically the next layer's GEMV) may start now: its producers prefetch weights —
    // which depend on nothing — into their ring while this kernel still runs; everything else of it waits (pdl_wait).
    pdl_launch_dependents();

    if (threadIdx.x == 0) {
        // full: one arrival per lane of the owning producer warp, triggered when that lane's cp.async copies landed
        for (uint32_t s = 0; s < a.stages; ++s) { mbar_init(full + s, 33); mbar_init(empty + s, kGroupWarps); }
        mbar_init(xfull, 1);
        fence_barrier_init();
    }
    __syncthreads();

    if (warp >= NG * kGroupWarps && warp < NG * kGroupWarps + NP) {
        // ===================== producers: NP warps take the stages round-robin =====================
        // (the mbarrier / bulk-copy instructions of ONE warp cost ~100 cycles each and do not overlap, so a
        //  single producer warp caps the SM at a fraction of its HBM share; NP warps and KBS tiles per copy
        //  lift that cap)
        const uint32_t me = (uint32_t)(warp - NG * kGroupWarps);
        if (!XR) pdl_wait();                    // the prepared activations come from the kernel before this one
        ItemIter iter(rg);
        Item item;
        uint32_t it0 = 0, my_it = me, my_s = me, my_ph = 0;
        while (iter.next(item)) {
            const uint32_t it1 = it0 + (item.kb1 - item.kb0 + KBS - 1) / KBS;
            const uint8_t *wsrc = a.packed + (size_t)item.nt * a.k_blocks * kWBytes;
            const uint2 *psrc = a.gparams + (size_t)item.nt * 128;
            for (; my_it < it1; my_it += NP) {
                const uint32_t kb = item.kb0 + (my_it - it0) * KBS;
                const uint32_t nk = item.kb1 - kb < (uint32_t)KBS ? item.kb1 - kb : (uint32_t)KBS;
                mbar_wait_relaxed(empty + my_s, my_ph ^ 1, 2000);
                if (lane == 0) STRACE(0, my_it);
                uint8_t *st = ring + (size_t)my_s * kStage;
                // The KBS code tiles are contiguous in HBM: one cp.async.bulk (complete_tx on the stage's mbarrier).
                // Scales / zero-points (and the activations, when they are not resident) ride on 16-byte cp.async of
                // the warp's 32 lanes; each lane's arrival on the same mbarrier fires when its copies have landed.
                // (Measured on 14336^2 4-bit: bulk 33.2 us, all-cp.async 34.6 us — the latter costs the producer
                // warps 16 more issue slots per lane and stage.)
                const uint8_t *src = wsrc + (size_t)kb * kWBytes;
                if (a.bulk) {
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full + my_s, nk * kWBytes);
                        bulk_load(st, src, nk * kWBytes, full + my_s);
                    }
                    __syncwarp();
                } else {
#pragma unroll 8
                    for (uint32_t i = lane; i < nk * (kWBytes / 16); i += 32) cp_async16(st + i * 16, src + (size_t)i * 16);
                    if (lane == 0) mbar_arrive_addr(smem_u32(full + my_s));
                }
#pragma unroll
                for (int sub = 0; sub < KBS; ++sub) {
                    if ((uint32_t)sub < nk) {
                        const uint2 *pg = psrc + (size_t)(a.group_magic ? __umulhi(kb + sub, a.group_magic) : kb + sub) * a.Npad;
                        cp_async16(st + KBS * kWBytes + sub * 1024 + lane * 16, pg + lane * 2);
                        cp_async16(st + KBS * kWBytes + sub * 1024 + 512 + lane * 16, pg + 64 + lane * 2);
                    }
                }
                if (!XR) {
                    const uint8_t *xsrc = a.xb + (size_t)kb * kXTile;
#pragma unroll 4
                    for (uint32_t i = lane; i < nk * (kXTile / 16); i += 32) cp_async16(st + KBS * (kWBytes + 1024) + i * 16, xsrc + (size_t)i * 16);
                }
                cp_async_arrive(full + my_s);
                if (lane == 0) STRACE(1, my_it);
                __syncwarp();
                my_s += NP;
                if (my_s >= a.stages) { my_s -= a.stages; my_ph ^= 1; }
            }
            it0 = it1;
        }
    } else if (warp < NG * kGroupWarps) {
        // ===================== consumers =====================
        const int grp = warp / kGroupWarps, w = warp % kGroupWarps;
        const int g = lane >> 2, t = lane & 3;
        const int ctid = threadIdx.x;            // consumer warps come first: 0 .. kConsumers-1
        float ya[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) ya[nb][0] = ya[nb][1] = ya[nb][2] = ya[nb][3] = 0.f;
        // per-lane shared-memory addresses (see consume_kblock), kept in registers (the empty asm stops the
        // compiler from re-deriving them from the shared window base inside the loop)
        const uint32_t r0 = 16 * w + g;
        uint32_t cwa = smem_u32(ring) + (r0 * 4 + t) * 4 + grp * kStage;      // codes of this group's next stage
        uint32_t dpw = KBS * kWBytes + r0 * 8 - (r0 * 4 + t) * 4;             // its parameters, relative to cwa
        uint32_t xs_lane = (XR ? smem_u32(xs) : KBS * (kWBytes + 1024) - (r0 * 4 + t) * 4) + (g < MT ? g : MT - 1) * 64 + 16 * t;   // !XR: relative to cwa
        uint32_t fa = smem_u32(full) + grp * 8;                               // its full barrier; empty = + 8 stages
        const uint32_t ring_bytes = a.stages * kStage, bar_bytes = a.stages * 8;
        // correction entry of the lane's token, relative to its activation unit; A-operand pattern of the correction MMA
        uint32_t dxc = MT * 128 - (g < MT ? g : MT - 1) * 48 - 12 * t;
        uint32_t zmask = t == 0 ? 0xffffffffu : 0u;
        uint32_t cfill = t == 1 ? 0x78007800u : t == 2 ? 0x24005000u : 0u;                    // half2(32768, 32768), half2(32, 2^-6)
        asm volatile("" : "+r"(cwa), "+r"(dpw), "+r"(xs_lane), "+r"(fa), "+r"(dxc), "+r"(zmask), "+r"(cfill));
        if (ctid == 0) GTRACE(1);
        // activations, partial-tile workspace, tickets and y belong to the stream's previous kernels until they are done
        // (the iterator's 64-bit divisions are done before the wait: nothing of it depends on the previous kernel)
        uint32_t n_item = 0, last_nt = 0;
        bool last_whole = true;
        ItemIter iter(rg);
        Item item;
        bool have_item = iter.next(item);
        pdl_wait();
        if (XR) prepare_x_slice<CB, MT>(a, rg, xs, ctid, kConsumers + 32);
        if (ctid == 0) GTRACE(2);
        // stage counter `it` of the CTA; this group owns the stages with it % NG == grp and walks only those
        uint32_t it0 = 0, my_it = (uint32_t)grp, my_s = (uint32_t)grp, my_ph = 0;
        for (; have_item; have_item = iter.next(item)) {
            const uint32_t len = item.kb1 - item.kb0;
            const uint32_t it1 = it0 + (len + KBS - 1) / KBS;
            uint32_t xa = XR ? xs_lane + (item.kb0 - rg.kb_s0 + (my_it - it0) * KBS) * kXTile : 0u;
            for (; my_it < it1; my_it += NG) {
                const uint32_t c = cwa + mbar_wait_token(fa, my_ph);
                if (ctid == 0 && my_it == 0) GTRACE(3);
                if (w == 0 && lane == 0) STRACE(2, my_it);
                const uint32_t xk = XR ? xa : c + xs_lane;
                consume_kblock<CB, MT>(c, c + dpw, xk, xk + dxc, zmask, cfill, ya);
                if (KBS == 2 && !(my_it + 1 == it1 && (len & 1)))               // the item's last stage may hold one k-block
                    consume_kblock<CB, MT>(c + kWBytes, c + dpw + 1024, xk + kXTile, xk + kXTile + dxc, zmask, cfill, ya);
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(fa + bar_bytes);
                if (w == 0 && lane == 0) STRACE(3, my_it);
                xa += NG * KBS * kXTile; cwa += NG * kStage; fa += NG * 8; my_s += NG;
                if (my_s >= a.stages) { my_s -= a.stages; cwa -= ring_bytes; fa -= bar_bytes; my_ph ^= 1; }
            }
            it0 = it1;
            if (ctid == 0) GTRACE(4 + 2 * n_item);
            // ---------- tile (or tile part) done: sum the consumer groups, store, hand over to the epilogue warp ----------
            float *mine = red + (size_t)grp * (MT * kRedStride);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const int tok = 8 * nb + 2 * t;
                if (tok < MT) { mine[tok * kRedStride + 16 * w + g] = ya[nb][0]; mine[tok * kRedStride + 16 * w + g + 8] = ya[nb][2]; }
                if (tok + 1 < MT) { mine[(tok + 1) * kRedStride + 16 * w + g] = ya[nb][1]; mine[(tok + 1) * kRedStride + 16 * w + g + 8] = ya[nb][3]; }
                ya[nb][0] = ya[nb][1] = ya[nb][2] = ya[nb][3] = 0.f;
            }
            named_bar_sync(kBarConsumers, kConsumers);
            const bool whole = a.S == 1 && item.kb0 == rg.kb_s0 && item.kb1 == rg.kb_s0 + rg.kbs;
            float *dst = whole ? nullptr : a.partial + ((size_t)blockIdx.x * a.max_items + item.ordinal) * (MT * 128);
            for (int e = ctid; e < MT * 128; e += kConsumers) {
                const int tok = e >> 7, nl = e & 127;
                float v = 0.f;
#pragma unroll
                for (int q = 0; q < NG; ++q) v += red[(size_t)q * (MT * kRedStride) + tok * kRedStride + nl];
                if (whole) {
                    const uint32_t n = item.nt * 128 + nl;
                    if ((uint32_t)tok < a.M && n < a.N) a.y[(size_t)tok * a.N + n] = v + (a.bias ? __ldg(a.bias + n) : 0.f);
                } else {
                    dst[e] = v;
                }
            }
            named_bar_sync(kBarConsumers, kConsumers);                  // `red` may be overwritten by the next tile
            // the partial tile is in global memory: the epilogue warp publishes it (fence + ticket) and, if this CTA
            // was the last contributor, reduces the tile — while the consumers are already on the next tile
            if (!whole) {
                named_bar_sync(kBarEpiFree, kConsumers + 32);
                named_bar_arrive(kBarPartial, kConsumers + 32);
            }
            if (ctid == 0) GTRACE(5 + 2 * n_item);
            ++n_item;
            last_whole = whole;
            last_nt = item.nt;
        }
        if (n_item != 0 && !last_whole) {
            // the CTA's last tile: wait for the verdict; if this CTA arrived last, all consumer threads reduce it
            named_bar_sync(kBarFinal, kConsumers + 32);
            if (verdict[0]) reduce_tile<MT>(a, verdict[1], slots, last_nt, ctid, kConsumers);
        }
    } else {
        // ===================== epilogue warp =====================
        pdl_wait();
        if (XR) prepare_x_slice<CB, MT>(a, rg, xs, kConsumers + lane, kConsumers + 32);
        ItemIter iter(rg);
        Item item;
        bool more = iter.next(item);
        while (more) {
            const Item cur = item;
            more = iter.next(item);
            const bool whole = a.S == 1 && cur.kb0 == rg.kb_s0 && cur.kb1 == rg.kb_s0 + rg.kbs;
            if (whole) continue;
            named_bar_arrive(kBarEpiFree, kConsumers + 32);
            named_bar_sync(kBarPartial, kConsumers + 32);
            uint32_t cnt = 0, last = 0;
            if (lane == 0) {
                __threadfence();                       // cumulative: publishes the consumers' stores ordered before the barrier
                // contributors of this tile, in the fixed order (segment, CTA): their partial slots
                for (uint32_t sg = 0; sg < a.S; ++sg) {
                    const uint32_t k0 = (uint32_t)((uint64_t)a.k_blocks * sg / a.S);
                    const uint32_t kn = (uint32_t)((uint64_t)a.k_blocks * (sg + 1) / a.S) - k0;
                    const uint64_t U = (uint64_t)a.n_tiles * kn, a0 = (uint64_t)cur.nt * kn, a1 = a0 + kn;
                    const uint32_t j_lo = (uint32_t)(((a0 + 1) * a.P - 1) / U), j_hi = (uint32_t)((a1 * a.P - 1) / U);
                    for (uint32_t j = j_lo; j <= j_hi; ++j) {
                        const uint32_t first_nt = (uint32_t)((U * j / a.P) / kn);
                        if (cnt < kMaxContrib) slots[cnt] = (sg * a.P + j) * a.max_items + (cur.nt - first_nt);
                        ++cnt;
                    }
                }
                const uint32_t old = atomicAdd(a.tickets + cur.nt, 1u);
                last = old + 1 == cnt ? 1u : 0u;
                if (last) a.tickets[cur.nt] = 0;       // every contributor has arrived: re-arm for the next launch
                __threadfence();
                verdict[0] = last;
                verdict[1] = cnt;
            }
            cnt = __shfl_sync(0xffffffffu, cnt, 0);
            last = __shfl_sync(0xffffffffu, last, 0);  // (the shuffle also orders lane 0's shared-memory writes before the reads)
            if (!more) {
                named_bar_arrive(kBarFinal, kConsumers + 32);           // the consumers take it from here
            } else if (last) {
                reduce_tile<MT>(a, cnt, slots, cur.nt, lane, 32);
            }
        }
    }
}

// x[M,K] f32 -> the prepared activation tiles of all k-blocks (activations too large to stay resident)
template <int CB, int MT>
__global__ void __launch_bounds__(256)
gemv_xprep_kernel(const float *__restrict__ x, uint32_t M, uint32_t K, uint32_t k_blocks, uint8_t *__restrict__ xb) {
    prepare_x_tiles<CB, MT, 1>(x, M, K, 0, k_blocks, xb, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

template <int CB, int MT, bool XR>
int32_t launch_gemv_mma(dllm_ctx *ctx, const dllm_qweight *qw, const float *x, size_t M, float *y) {
    // NP == NG: a ring slot must always be filled by the same producer warp and drained by the same consumer
    // group (the stage count is a multiple of both) — parity waits of different warps on one slot could alias
    constexpr int NG = 3, NP = 3, KBS = 2;
    static_assert(NG % NP == 0 || NP % NG == 0, "ring depth is a multiple of max(NG, NP) only");
    constexpr int kWBytes = WL_TILE_N * WL_TILE_K * CB / 8;
    constexpr int kXTile = gemv_x_tile_bytes(MT);
    constexpr int kStage = KBS * (kWBytes + 1024 + (XR ? 0 : kXTile));
    const uint32_t k_blocks = (uint32_t)qw->k_blocks, n_tiles = (uint32_t)qw->n_tiles;
    const uint32_t sms = (uint32_t)ctx->sm_count;

    // XR: the whole activation block stays resident (one k segment).  The segment machinery (S > 1: a CTA keeps only
    // its k segment's slice) remains for activations that are too large for that but were asked to stay resident.
    const uint32_t red_bytes = NG * MT * kRedStride * 4;
    uint32_t S = 1;
    if (XR) {
        for (;; ++S) {
            const uint64_t xb = (uint64_t)((k_blocks + S - 1) / S) * kXTile;
            if (S >= k_blocks || (xb <= (uint64_t)kXBudget && xb + red_bytes + 1024 + 8ull * kStage <= (uint64_t)kSmemBudget)) break;
        }
    }
    uint32_t P = sms / S;
    if (P == 0) { P = 1; }
    // small problems: at least 4 k-blocks per CTA
    const uint64_t total_units = (uint64_t)n_tiles * k_blocks;
    while (P > 1 && total_units / ((uint64_t)S * P) < 4) --P;
    while (S * P > (uint32_t)kMaxContrib) --P;
    static const int grid_cap = getenv("DLLM_GEMV_GRID") ? atoi(getenv("DLLM_GEMV_GRID")) : 0;   // experiments only
    if (grid_cap > 0) while (P > 1 && S * P > (uint32_t)grid_cap) --P;
    const uint32_t grid = S * P;

    GemvArgs a;
    a.packed = qw->d_packed; a.gparams = qw->d_gparams; a.bias = qw->d_bias;
    a.y = y;
    a.M = (uint32_t)M; a.N = (uint32_t)qw->N; a.Npad = n_tiles * 128; a.k_blocks = k_blocks; a.n_tiles = n_tiles;
    const uint32_t group_kb = qw->per_tensor ? k_blocks : (uint32_t)(qw->group / WL_TILE_K);
    if (k_blocks >= 65536) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "GEMV path: K too large");
    a.group_magic = group_kb == 1 ? 0u : (uint32_t)(((1ull << 32) + group_kb - 1) / group_kb);   // 0: one k-block per group
    a.S = S; a.P = P;
    static const uint32_t bulk_mode = getenv("DLLM_GEMV_BULK") ? (uint32_t)atoi(getenv("DLLM_GEMV_BULK")) : 1u;
    a.bulk = bulk_mode;
    a.trace = nullptr;
#ifdef DLLM_GEMV_TRACE
    DLLM_TRY(ensure_buf(ctx, ctx->lin_flags, ((size_t)grid * 32 + 1024) * 8));
    a.trace = (unsigned long long *)ctx->lin_flags.p;
    cudaMemsetAsync(a.trace, 0, ((size_t)grid * 32 + 1024) * 8, ctx->stream);
#endif
    const uint32_t kbs_max = (k_blocks + S - 1) / S, kbs_min = k_blocks / S;
    const uint64_t len_max = ((uint64_t)n_tiles * kbs_max + P - 1) / P;
    a.max_items = (uint32_t)(len_max / (kbs_min ? kbs_min : 1)) + 2;

    const uint32_t xbytes = XR ? kbs_max * kXTile : 0u;
    uint32_t stages = (uint32_t)((kSmemBudget - xbytes - red_bytes - 1024) / kStage);
    if (stages > (uint32_t)kMaxStages) stages = kMaxStages;
    stages -= stages % (NG > NP ? NG : NP);
    if (stages < 2 * NG || stages <= (uint32_t)NP) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "GEMV path: shared-memory ring too small");
    a.stages = stages;
    a.x_off = stages * kStage;
    a.red_off = a.x_off + ((xbytes + 127) & ~127u);
    a.bar_off = a.red_off + ((red_bytes + 127) & ~127u);
    const size_t smem_bytes = a.bar_off + (2 * stages + 1) * 8 + (2 + kMaxContrib) * 4;

    if (!XR) DLLM_TRY(ensure_buf(ctx, ctx->act[2], (size_t)k_blocks * kXTile));
    DLLM_TRY(ensure_buf(ctx, ctx->lin_ws, (size_t)grid * a.max_items * MT * 128 * sizeof(float)));
    if (ctx->gemv_tickets.cap < n_tiles * sizeof(unsigned int)) {
        DLLM_TRY(ensure_buf(ctx, ctx->gemv_tickets, n_tiles * sizeof(unsigned int)));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->gemv_tickets.p, 0, ctx->gemv_tickets.cap, ctx->stream));
    }
    a.xb = (const uint8_t *)ctx->act[2].p;
    a.x = x; a.K = (uint32_t)qw->K;
    a.partial = (float *)ctx->lin_ws.p;
    a.tickets = (unsigned int *)ctx->gemv_tickets.p;

    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(gemv_mma_kernel<CB, MT, NG, NP, KBS, XR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        // the activation-prep kernel runs right before: same shared-memory carve-out, so the SMs are not
        // re-partitioned (L1 vs shared) between the two launches
        CUDA_TRY(ctx, cudaFuncSetAttribute(gemv_xprep_kernel<CB, MT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    if (!XR) {
        const uint32_t units16 = k_blocks * 2 * MT * 4;
        gemv_xprep_kernel<CB, MT><<<(units16 + 255) / 256, 256, 0, ctx->stream>>>(x, (uint32_t)M, (uint32_t)qw->K, k_blocks, (uint8_t *)ctx->act[2].p);
        LAUNCH_CHECK(ctx);
    }
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
    {
        // programmatic stream serialization: this kernel may be scheduled before the stream's previous kernel has
        // finished (it waits in-kernel where it depends on it, see pdl_wait)
        static const bool no_pdl = getenv("DLLM_GEMV_NO_PDL") != nullptr;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3((NG * kGroupWarps + NP + 1) * 32);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = no_pdl ? 0 : 1;
        CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, gemv_mma_kernel<CB, MT, NG, NP, KBS, XR>, a));
    }
    LAUNCH_CHECK(ctx);
    if (ev1) {
        CUDA_TRY(ctx, cudaEventRecord(ev1, ctx->stream));
        ctx->prof_n++;
        ctx->prof_flops += 2.0 * (double)M * (double)qw->K * (double)qw->N;
        ctx->prof_bytes += (double)qw->K * qw->N * qw->bits / 8.0 + (double)(qw->K / qw->group) * qw->N * 8.0 +
                           4.0 * M * qw->K + 4.0 * M * qw->N;
    }
#ifdef DLLM_GEMV_TRACE
    {   // dump the per-CTA timeline (timing experiments only)
        std::vector<unsigned long long> h((size_t)grid * 32 + 1024);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h.data(), a.trace, h.size() * 8, cudaMemcpyDeviceToHost);
        FILE *f = fopen("gpurun_out/gemv_trace.csv", "w");
        if (f) {
            unsigned long long t0 = ~0ull;
            for (uint32_t c = 0; c < grid; ++c) if (h[c * 32] && h[c * 32] < t0) t0 = h[c * 32];
            fprintf(f, "cta,start,cons_start,x_ready,first_full,then (item_done, flush_done)...  [ns since the first CTA started]\n");
            for (uint32_t c = 0; c < grid; ++c) {
                fprintf(f, "%u", c);
                for (int i = 0; i < 32; ++i) fprintf(f, ",%lld", h[c * 32 + i] ? (long long)(h[c * 32 + i] - t0) : -1ll);
                fprintf(f, "\n");
            }
            fclose(f);
        }
        f = fopen("gpurun_out/gemv_stages.csv", "w");
        if (f) {
            unsigned long long t0 = h[0];
            fprintf(f, "it,prod_empty_ok,prod_issued,cons_full_ok,cons_done  [ns since CTA 0 started]\n");
            for (int i = 0; i < 256; ++i) {
                fprintf(f, "%d", i);
                for (int r = 0; r < 4; ++r) { unsigned long long v = h[(size_t)grid * 32 + r * 256 + i]; fprintf(f, ",%lld", v ? (long long)(v - t0) : -1ll); }
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
#endif
    return DLLM_OK;
}

// the prepared activations of up to kXResident bytes stay resident in shared memory; larger ones ride the ring
constexpr size_t kXResident = 64 * 1024;

template <int CB, int MT>
int32_t launch_gemv_x(dllm_ctx *ctx, const dllm_qweight *qw, const float *x, size_t M, float *y) {
    if (qw->k_blocks * (size_t)gemv_x_tile_bytes(MT) <= kXResident) return launch_gemv_mma<CB, MT, true>(ctx, qw, x, M, y);
    return launch_gemv_mma<CB, MT, false>(ctx, qw, x, M, y);
}

template <int CB>
int32_t launch_gemv_mt(dllm_ctx *ctx, const dllm_qweight *qw, const float *x, size_t M, float *y) {
    if (M <= 1) return launch_gemv_x<CB, 1>(ctx, qw, x, M, y);
    if (M <= 2) return launch_gemv_x<CB, 2>(ctx, qw, x, M, y);
    if (M <= 4) return launch_gemv_x<CB, 4>(ctx, qw, x, M, y);
    if (M <= 8) return launch_gemv_x<CB, 8>(ctx, qw, x, M, y);
    return launch_gemv_x<CB, 16>(ctx, qw, x, M, y);
}

}  // namespace

bool k_gemv_supported(const dllm_qweight *qw, size_t M) {
    // fp16 operands: integer zero-points (subtracted exactly) and scales in fp16's full-precision range
    return qw && M >= 1 && M <= 16 && (qw->per_tensor || qw->group % WL_TILE_K == 0) && qw->int_zps;
}

int32_t k_qlinear_gemv(dllm_ctx *ctx, const dllm_qweight *qw, const float *x_dev, size_t M, float *y_dev) {
    if (!k_gemv_supported(qw, M)) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "GEMV path: 1 <= M <= 16 required");
    switch (wl_container_bits(qw->bits)) {
        case 2: return launch_gemv_mt<2>(ctx, qw, x_dev, M, y_dev);
        case 4: return launch_gemv_mt<4>(ctx, qw, x_dev, M, y_dev);
        default: return launch_gemv_mt<8>(ctx, qw, x_dev, M, y_dev);
    }
}
