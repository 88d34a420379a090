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
//   * the CTA's slice of the activations is staged ONCE into shared memory (signed-digit int8 columns in MMA fragment
//     order, prepared inside the kernel by the consumer warps while the producers already stream weights) and stays
//     resident: no activation traffic in the steady state;
//   * NG groups of 8 consumer warps take the stages round-robin.  A warp owns 16 output columns of the
//     128-column tile: it reads its codes with conflict-free LDS.32, spreads the nibbles to bytes (two ANDs and a shift
//     per eight 4-bit codes; 8-bit codes are used as they are) and feeds them as the unsigned A operand of
//     mma.sync.m16n8k32 (u8 x s8 -> s32) with the activations' digit columns as B — two MMAs per 64-k block, exact
//     integer sums; zero-point, scale and the activation block's power-of-two step are applied to the k-block's partial
//     sum in f32 (see "the int8 tensor path" below).  The legacy tensor path costs 8 cycles of a sub-core's pipe per
//     instruction whatever the type, so the f16 m16n8k16 this replaces (5 per k-block with the zero-point correction)
//     was what bounded the 4-bit kernel (measured with instruction-removal builds, scripts/build_variants.sh);
//   * tiles cut by a range boundary are reduced by the LAST CTA to arrive at the tile (atomic ticket,
//     no spinning), always in CTA order: results are deterministic and there is no fix-up launch.
//
// Numerics: codes, zero-points and all integer sums exact; x is rounded to 23-bit (1-2 tokens) or 15-bit (4-16 tokens)
// block fixed point per 64 activations — |error| <= 2^-22 resp. 2^-14 of the block's max |x| per element — f32
// scale and f32 accumulation across k-blocks.  Zero-points must be the integers quantizer B produces (quantization.rs:55-56).
// Timeline instrumentation (globaltimer stamps per CTA / per stage) compiles in with -DDLLM_GEMV_TRACE.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <stdio.h>
#include <vector>

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
constexpr int kPreTiles = 2;                   // partial tiles of other CTAs a CTA may prefetch for its final tile
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
// The packed weights are read once per call: evict-first keeps them from washing everything else out of the L2 — the
// activations, the partial tiles and, above all, this kernel's own INSTRUCTIONS (measured: the first instructions after
// the dependency wait took 3-4 us to arrive, the time of an instruction fetch from HBM under full streaming load).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
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
// D(16 weight columns x 8 activation columns, s32) = A(16 x 32 k, u8) · B(32 k x 8, s8) + C
__device__ __forceinline__ void mma_u8s8(int *d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1, const int *c) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}
// ---- the int8 tensor path ---------------------------------------------------------------------------------------
// mma.sync.m16n8k32 (u8 x s8 -> s32, SASS IMMA.16832) costs the same 8 cycles of a sub-core's tensor pipe as the f16
// m16n8k16 (measured: scripts/ubench/mma_rate.cu) and covers twice the k.  A = the codes themselves as unsigned bytes (no
// magic numbers, no zero-point subtraction per weight), B = the activations as signed-digit fixed point: per (k-block,
// token) x ~ delta * (d0 * 256^(P-1) + ... + d_{P-1}), digits in [-128, 127], delta a power of two chosen from the block's
// max |x|.  Every digit is its own B COLUMN of the same MMA (the weights are shared), so P digits cost nothing extra while
// tokens * P <= 8.  The int32 sums are exact; the zero-point leaves as zp * sum(digits) per (k-block, column), the scale and
// delta * 256^j are applied in f32 to the k-block's partial sum.
//
// Activation columns: P = 3 digits (|error| <= 2^-22 max|x| of the 64-element block per element: f32-grade) for 1-2
// tokens, 2 digits (<= 2^-14 max|x|) for 4..16 tokens; column c = token * P + digit.
__host__ __device__ constexpr int gemv_parts(int MT) { return MT <= 2 ? 3 : 2; }
__host__ __device__ constexpr int gemv_cols(int MT) { return MT * gemv_parts(MT); }
__host__ __device__ constexpr int gemv_nb(int MT) { return (gemv_cols(MT) + 7) / 8; }          // 8-column MMA blocks
// bytes of one k-block of prepared activations: [columns][4 t] 16-byte units {u0 b0, u0 b1, u1 b0, u1 b1} (the B registers
// of the k-block's two MMAs for lane quad-index t), then per column block [4 t] entries {factor(2t), factor(2t+1),
// digit sum(2t), digit sum(2t+1)} (f32) for the accumulator columns lane t holds
__host__ __device__ constexpr int gemv_x_tile_bytes(int MT) { return gemv_cols(MT) * 64 + gemv_nb(MT) * 64; }

// k (inside a 64-k block) of byte j (0..3) of B register r (0..1) of MMA u (0..1) for lane quad-index t: whatever code
// the unpack of wlayout.cuh's words puts at that position of the A fragment (see consume_kblock)
template <int CB>
__host__ __device__ __forceinline__ constexpr int gemv_kmap(int u, int t, int r, int j) {
    // 4-bit: word t of chunk u = codes 32 u + 8 t + e at nibble (e >> 1) + 4 (e & 1); w & 0x0f0f0f0f = e {0,4,1,5}, (w >> 4) & .. = e {2,6,3,7}
    if (CB == 4) return 32 * u + 8 * t + (r ? 2 : 0) + (j >> 1) + 4 * (j & 1);
    // 2-bit: word t = codes 16 t + i at field (i >> 1) + 8 (i & 1); (w >> 2 s) & 0x03030303 = i {2s, 8+2s, 2s+1, 9+2s}, s = 2 u + r
    if (CB == 2) return 16 * t + 2 * (2 * u + r) + (j >> 1) + 8 * (j & 1);
    // 8-bit: word t of chunk c = codes 16 c + 4 t + i in byte order; c = 2 u + r
    return 16 * (2 * u + r) + 4 * t + j;
}

#ifdef DLLM_GEMV_TRACE
constexpr int kTraceSlots = 64;
unsigned long long *g_trace_buf = nullptr;
size_t g_trace_stride = 0, g_trace_launch = 0;
uint32_t g_trace_grid = 0;
#endif

struct GemvArgs {
    const uint8_t *packed;
    const uint2 *gparams;            // [G][Npad] {f32 scale, f32 zp}
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
    uint32_t prefill;                // stages requested before the activations are resident (XR)
    uint32_t x_off, red_off, bar_off, pre_off;   // shared-memory carve-up (bytes)
    uint32_t fast;                   // 1: every aligned stage lies in one quantization group, 2: every aligned pair does, 0: neither
    uint32_t bulk;                   // code tiles by one cp.async.bulk per stage (default) or by per-lane cp.async (DLLM_GEMV_BULK=0)
    unsigned long long *trace;       // -DDLLM_GEMV_TRACE: per CTA [32] globaltimer stamps, then [4][256] stage stamps of CTA 0
};

// the k-segment and unit range of one CTA; units of a segment are ordered (tile, k-block).
// 32-bit arithmetic throughout (the launch checks units * (P + 1) < 2^32): an inlined 64-bit division is ~150 instructions,
// and the prologue of this kernel runs instruction-cache-cold on every launch.
struct Range {
    uint32_t kb_s0, kbs;             // segment = k-blocks [kb_s0, kb_s0 + kbs)
    uint32_t units, u0, u1;          // units of the segment; this CTA's range
    __device__ Range(const GemvArgs &a, uint32_t cta) {
        const uint32_t seg = cta / a.P, j = cta - seg * a.P;
        kb_s0 = a.k_blocks * seg / a.S;
        kbs = a.k_blocks * (seg + 1) / a.S - kb_s0;
        units = a.n_tiles * kbs;
        u0 = units * j / a.P;
        u1 = units * (j + 1) / a.P;
    }
};

struct Item {
    uint32_t nt, kb0, kb1, ordinal;
};

struct ItemIter {
    uint32_t u, u1;
    uint32_t kb_s0, kbs, first_nt;
    __device__ ItemIter(const Range &r) : u(r.u0), u1(r.u1), kb_s0(r.kb_s0), kbs(r.kbs) { first_nt = kbs ? r.u0 / kbs : 0; }
    __device__ bool next(Item &it) {
        if (u >= u1) return false;
        it.nt = u / kbs;
        const uint32_t off = u - it.nt * kbs;
        const uint32_t left = u1 - u;
        const uint32_t len = kbs - off <= left ? kbs - off : left;
        it.kb0 = kb_s0 + off;
        it.kb1 = it.kb0 + len;
        it.ordinal = it.nt - first_nt;
        u += len;
        return true;
    }
};

// The stages of an item are ALIGNED groups of KBS k-blocks: k-block kb always sits in slot kb % KBS of its stage (the
// item's first and last stage may be partly empty), so the aligned pairs (2j, 2j + 1) — one quantization group when the
// group is >= 128 — always meet in one stage whatever k-block the stream-K range starts at.
template <int KBS>
__device__ __forceinline__ uint32_t item_stages(const Item &it) {
    return (it.kb1 - (it.kb0 & ~(uint32_t)(KBS - 1)) + KBS - 1) / KBS;
}
// quantization group of a k-block (magic == 0: one k-block per group)
__device__ __forceinline__ uint32_t kb_group(uint32_t kb, uint32_t magic) { return magic ? __umulhi(kb, magic) : kb; }

// one 64-k block of a 128-column tile, consumed by the 8 warps of one group.
// Shared-memory addresses, each already offset to this lane's element (lane = 4 g + t, output columns
// r0 = 16 w + g and r0 + 8):
//   cw : word t of chunk 0 of column r0 of the packed codes (wlayout.cuh): + 2048 per chunk, + 128 for column r0 + 8
//   pw : {f32 scale, f32 zp} of column r0: + 64 for column r0 + 8
//   xw : the lane's 16-byte unit of activation column min(g, columns - 1): + 512 per column block.  (Lanes beyond the
//        last column feed MMA columns that are never stored; they re-read the last one instead of zeroing registers.)
//   xf : the lane's factor entry of column block 0: + 64 per column block
// ya[nb] = the lane's f32 accumulators: columns 8 nb + 2 t, + 1 of weight column r0, then the same of r0 + 8.
// A fragments of one k-block's two MMAs: rows r0 / r0 + 8, k positions as gemv_kmap says
template <int CB>
__device__ __forceinline__ void load_a_frags(uint32_t cw, uint32_t (*a)[4]) {
    if (CB == 4) {
        uint32_t q[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) { q[u][0] = lds32(cw + u * 2048); q[u][1] = lds32(cw + u * 2048 + 128); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            a[u][0] = q[u][0] & 0x0f0f0f0fu; a[u][1] = q[u][1] & 0x0f0f0f0fu;
            a[u][2] = (q[u][0] >> 4) & 0x0f0f0f0fu; a[u][3] = (q[u][1] >> 4) & 0x0f0f0f0fu;
        }
    } else if (CB == 2) {
        const uint32_t q0 = lds32(cw), q1 = lds32(cw + 128);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            a[u][0] = (q0 >> (4 * u)) & 0x03030303u; a[u][1] = (q1 >> (4 * u)) & 0x03030303u;
            a[u][2] = (q0 >> (4 * u + 2)) & 0x03030303u; a[u][3] = (q1 >> (4 * u + 2)) & 0x03030303u;
        }
    } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            a[u][0] = lds32(cw + (2 * u) * 2048); a[u][1] = lds32(cw + (2 * u) * 2048 + 128);
            a[u][2] = lds32(cw + (2 * u + 1) * 2048); a[u][3] = lds32(cw + (2 * u + 1) * 2048 + 128);
        }
    }
}

template <int CB, int MT>
__device__ __forceinline__ void consume_kblock(uint32_t cw, uint32_t pw, uint32_t xw, uint32_t xf, float (*ya)[4]) {
    constexpr int NB = gemv_nb(MT);
    // s32 accumulators start at the bits of the float 1.5 * 2^23: while |sum| < 2^22 the integer sum IS the float
    // 12582912 + sum, so no I2F (a quarter-rate instruction) is needed.  |sum| <= 64 * 128 * 255 < 2^21.
    constexpr int kMagicBits = 0x4B400000;
    constexpr float kMagic = 12582912.f;
    const int cinit[4] = {kMagicBits, kMagicBits, kMagicBits, kMagicBits};
    const uint2 p0 = lds64(pw), p1 = lds64(pw + 64);
    uint32_t a[2][4];
    load_a_frags<CB>(cw, a);
    // dequantize_tensor's `(q - zp) * scale` (quantization.rs:83): sum x q - zp sum x, then * scale, per k-block in f32
    const float s0 = __uint_as_float(p0.x), s1 = __uint_as_float(p1.x);
    const float nz0 = -__uint_as_float(p0.y), nz1 = -__uint_as_float(p1.y);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        const uint4 b = lds128(xw + nb * 512);
        const uint4 f = lds128(xf + nb * 64);
        int d[4];
        mma_u8s8(d, a[0][0], a[0][1], a[0][2], a[0][3], b.x, b.y, cinit);
        mma_u8s8(d, a[1][0], a[1][1], a[1][2], a[1][3], b.z, b.w, d);
        const float fac0 = __uint_as_float(f.x), fac1 = __uint_as_float(f.y), sx0 = __uint_as_float(f.z), sx1 = __uint_as_float(f.w);
        // every intermediate is an integer below 2^24: exact
        const float e0 = fmaf(nz0, sx0, __int_as_float(d[0]) - kMagic), e1 = fmaf(nz0, sx1, __int_as_float(d[1]) - kMagic);
        const float e2 = fmaf(nz1, sx0, __int_as_float(d[2]) - kMagic), e3 = fmaf(nz1, sx1, __int_as_float(d[3]) - kMagic);
        ya[nb][0] = fmaf(s0 * fac0, e0, ya[nb][0]);
        ya[nb][1] = fmaf(s0 * fac1, e1, ya[nb][1]);
        ya[nb][2] = fmaf(s1 * fac0, e2, ya[nb][2]);
        ya[nb][3] = fmaf(s1 * fac1, e3, ya[nb][3]);
    }
}

// The two k-blocks of an aligned pair (2j, 2j + 1) that share their quantization parameters (group >= 128): their four MMAs
// chain into ONE int32 sum per accumulator, and the float work — zero-point, scale, the activations' step — is done once for
// 128 k instead of once per 64.  The activations of an aligned pair share their power-of-two step (prepare_x_tiles), so one
// factor covers both k-blocks; the digit sums of the two add exactly.  |sum| <= 128 * 128 * 255 < 2^22 still holds, and the
// magic offset leaves BEFORE the zero-point term joins so that every intermediate stays below 2^24.
// The second k-block's codes are `cstep` bytes after the first one's, its activations kXTile bytes after.
template <int CB, int MT>
__device__ __forceinline__ void consume_pair(uint32_t cw, uint32_t cstep, uint32_t pw, uint32_t xw, uint32_t xf, float (*ya)[4]) {
    constexpr int NB = gemv_nb(MT);
    constexpr int kXTile = gemv_x_tile_bytes(MT);
    constexpr int kMagicBits = 0x4B400000;
    constexpr float kMagic = 12582912.f;
    const int cinit[4] = {kMagicBits, kMagicBits, kMagicBits, kMagicBits};
    const uint2 p0 = lds64(pw), p1 = lds64(pw + 64);
    uint32_t a[4][4];
    load_a_frags<CB>(cw, a);
    load_a_frags<CB>(cw + cstep, a + 2);
    const float s0 = __uint_as_float(p0.x), s1 = __uint_as_float(p1.x);
    const float nz0 = -__uint_as_float(p0.y), nz1 = -__uint_as_float(p1.y);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        const uint4 b0 = lds128(xw + nb * 512), b1 = lds128(xw + kXTile + nb * 512);
        const uint4 f0 = lds128(xf + nb * 64);
        const uint2 f1 = lds64(xf + kXTile + nb * 64 + 8);               // the second k-block's digit sums
        int d[4];
        mma_u8s8(d, a[0][0], a[0][1], a[0][2], a[0][3], b0.x, b0.y, cinit);
        mma_u8s8(d, a[1][0], a[1][1], a[1][2], a[1][3], b0.z, b0.w, d);
        mma_u8s8(d, a[2][0], a[2][1], a[2][2], a[2][3], b1.x, b1.y, d);
        mma_u8s8(d, a[3][0], a[3][1], a[3][2], a[3][3], b1.z, b1.w, d);
        const float fac0 = __uint_as_float(f0.x), fac1 = __uint_as_float(f0.y);
        const float sx0 = __uint_as_float(f0.z) + __uint_as_float(f1.x), sx1 = __uint_as_float(f0.w) + __uint_as_float(f1.y);
        const float e0 = fmaf(nz0, sx0, __int_as_float(d[0]) - kMagic), e1 = fmaf(nz0, sx1, __int_as_float(d[1]) - kMagic);
        const float e2 = fmaf(nz1, sx0, __int_as_float(d[2]) - kMagic), e3 = fmaf(nz1, sx1, __int_as_float(d[3]) - kMagic);
        ya[nb][0] = fmaf(s0 * fac0, e0, ya[nb][0]);
        ya[nb][1] = fmaf(s0 * fac1, e1, ya[nb][1]);
        ya[nb][2] = fmaf(s1 * fac0, e2, ya[nb][2]);
        ya[nb][3] = fmaf(s1 * fac1, e3, ya[nb][3]);
    }
}

// The 16 activations of lane quad-index t in one k-block are 4 runs of 4 consecutive k: k of run q's first element
template <int CB>
__host__ __device__ __forceinline__ constexpr int gemv_run_k0(int q, int t) {
    return CB == 4 ? 32 * (q >> 1) + 8 * t + 4 * (q & 1) : CB == 2 ? 16 * t + 4 * q : 16 * q + 4 * t;
}
// position (4 * run + index) in those runs of the activation that byte j of B register r of MMA u needs
template <int CB>
__host__ __device__ __forceinline__ constexpr int gemv_run_pos(int u, int r, int j) {
    const int k = gemv_kmap<CB>(u, 0, r, j);
    for (int q = 0; q < 4; ++q)
        if (k >= gemv_run_k0<CB>(q, 0) && k < gemv_run_k0<CB>(q, 0) + 4) return 4 * q + (k - gemv_run_k0<CB>(q, 0));
    return -1;
}

// four consecutive activations of a token with every guard (ragged K, tokens beyond M, unaligned x): out of line, so the
// common path below is four straight-line LDG.128 — the instructions between the dependency wait and the first load are
// executed cold (~100 ns per 128-byte line of code), and the inline guards were most of them
__device__ __noinline__ float4 gemv_x_load4_guarded(const float *__restrict__ x, uint32_t M, uint32_t K, uint32_t tok, uint32_t k) {
    float4 f;
    f.x = (tok < M && k < K) ? __ldcg(x + (size_t)tok * K + k) : 0.f;
    f.y = (tok < M && k + 1 < K) ? __ldcg(x + (size_t)tok * K + k + 1) : 0.f;
    f.z = (tok < M && k + 2 < K) ? __ldcg(x + (size_t)tok * K + k + 2) : 0.f;
    f.w = (tok < M && k + 3 < K) ? __ldcg(x + (size_t)tok * K + k + 3) : 0.f;
    return f;
}
// loads of one (k-block, token, t) work item: 16 floats in run order.  `vec` = K % 4 == 0 and x is 16-byte aligned.
// The load and the conversion are separate so that a thread can have the loads of several items in flight.
template <int CB>
__device__ __forceinline__ void gemv_x_load(const float *__restrict__ x, uint32_t M, uint32_t K, bool vec, uint32_t kb, uint32_t tok, uint32_t t, float *v) {
    const bool fast = vec && tok < M && (kb + 1) * WL_TILE_K <= K;          // the whole k-block exists
    const float *row = x + (size_t)tok * K + (size_t)kb * WL_TILE_K;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t kk = (uint32_t)gemv_run_k0<CB>(q, (int)t);
        const float4 f = fast ? __ldcg(reinterpret_cast<const float4 *>(row + kk)) : gemv_x_load4_guarded(x, M, K, tok, kb * WL_TILE_K + kk);
        v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
    }
}

// the work items of the k-blocks [kb0, kb0 + n_kb) into `dst` (tile of k-block kb0 first), all threads of whole warps:
// 4 consecutive lanes (t = 0..3) hold the 64 activations of one (k-block, token), so its max |x| and digit sums are two
// shuffles away.  UNR items per thread and pass, all loads issued before the first conversion.
#ifdef DLLM_GEMV_TRACE
#define XTRACE(slot) do { if (xtrace && tid == 0) { unsigned long long _t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(_t)); xtrace[slot] = _t; } } while (0)
#else
#define XTRACE(slot) do { } while (0)
#endif
template <int CB, int MT, int UNR>
__device__ __forceinline__ void prepare_x_tiles(const float *__restrict__ x, uint32_t M, uint32_t K, uint32_t kb0, uint32_t n_kb,
                                                uint8_t *dst, uint32_t tid, uint32_t nthreads, unsigned long long *xtrace = nullptr) {
    constexpr int P = gemv_parts(MT), NC = gemv_cols(MT), NB = gemv_nb(MT);
    // |x / delta| <= 2^kTop: leaves the top digit within [-64, 64] and lets the float -> int conversion be one FFMA
    // (x * (1/delta) + 1.5 * 2^23: the sum's low mantissa bits are the rounded integer; F2I is a quarter-rate instruction)
    constexpr int kTop = P == 3 ? 22 : 14;
    constexpr int kMagicBits = 0x4B400000;
    const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    // work item = (aligned k-block pair, token, half, t): the 8 lanes of a (pair, token) share the step delta, whatever
    // part of the pair [kb0, kb0 + n_kb) covers — a CTA whose range cuts a pair computes the same delta as its neighbour
    const uint32_t kbA0 = kb0 & ~1u;
    const uint32_t total = n_kb ? (((kb0 + n_kb + 1) >> 1) - (kb0 >> 1)) * MT * 8 : 0u;
    XTRACE(20);
#pragma unroll 1
    for (uint32_t base = tid & ~31u; base < total; base += UNR * nthreads) {
        float v[UNR][16];
        XTRACE(21);
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const uint32_t idx = base + j * nthreads + (tid & 31);
            if (idx < total) {
                gemv_x_load<CB>(x, M, K, vec, kbA0 + 2 * ((idx >> 3) / MT) + ((idx >> 2) & 1), (idx >> 3) % MT, idx & 3, v[j]);
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[j][e] = 0.f;
            }
        }
        XTRACE(24);
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            if (base + j * nthreads >= total) break;                      // (warp-uniform) nothing left for this warp
            const uint32_t idx = base + j * nthreads + (tid & 31);
            const uint32_t t = idx & 3, tok = (idx >> 3) % MT, kb = kbA0 + 2 * ((idx >> 3) / MT) + ((idx >> 2) & 1);
            // total % 8 == 0: the 8 lanes of a (pair, token) are all inside or all outside; only k-blocks of the range are stored
            const bool valid = idx < total && kb >= kb0 && kb < kb0 + n_kb;
            // block scale over the pair's 128 activations: delta = 2^(E - kTop) with max|x| < 2^E
            float m = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) m = fmaxf(m, fabsf(v[j][e]));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fminf(m, 3.0e38f);
            if (j == 0 && m >= 0.f) XTRACE(25);
            int E = (int)(__float_as_uint(m) >> 23) - 126;
            E = E < -100 ? -100 : E;
            const float inv = __int_as_float((127 + kTop - E) << 23), delta = __int_as_float((127 + E - kTop) << 23);
            // signed digits (balanced, low to high: d = (int8)(r & 255), r = (r + 128) >> 8), packed straight into the B
            // registers: word ur = register ur & 1 of MMA ur >> 1.  (Non-finite activations give unspecified finite digits.)
            uint32_t w[P][4];
            int sum[P];
#pragma unroll
            for (int pp = 0; pp < P; ++pp) sum[pp] = 0;
#pragma unroll
            for (int ur = 0; ur < 4; ++ur) {
                int r0[4], r1[4], r2[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int e = gemv_run_pos<CB>(ur >> 1, ur & 1, jj);
                    r0[jj] = __float_as_int(fmaf(v[j][e], inv, 12582912.f));          // low byte = digit 0 (the magic's is 0)
                    r1[jj] = (r0[jj] + (128 - kMagicBits)) >> 8;
                    r2[jj] = (r1[jj] + 128) >> 8;
                }
                // digit `lvl` (0 = lowest) of the four elements -> one word; part 0 is the TOP digit
#pragma unroll
                for (int lvl = 0; lvl < P; ++lvl) {
                    const int *r = lvl == 0 ? r0 : lvl == 1 ? r1 : r2;
                    const uint32_t lo = __byte_perm((uint32_t)r[0], (uint32_t)r[1], 0x0040), hi = __byte_perm((uint32_t)r[2], (uint32_t)r[3], 0x0040);
                    const uint32_t word = __byte_perm(lo, hi, 0x5410);
                    w[P - 1 - lvl][ur] = word;
                    sum[P - 1 - lvl] = __dp4a((int)word, 0x01010101, sum[P - 1 - lvl]);
                }
            }
#pragma unroll
            for (int pp = 0; pp < P; ++pp) {
                sum[pp] += __shfl_xor_sync(0xffffffffu, sum[pp], 1);
                sum[pp] += __shfl_xor_sync(0xffffffffu, sum[pp], 2);
            }
            if (valid) {
                uint8_t *tile = dst + (size_t)(kb - kb0) * gemv_x_tile_bytes(MT);
#pragma unroll
                for (int pp = 0; pp < P; ++pp) {
                    const uint32_t c = tok * P + pp;
                    *reinterpret_cast<uint4 *>(tile + c * 64 + t * 16) = make_uint4(w[pp][0], w[pp][1], w[pp][2], w[pp][3]);
                    if (t == 0) {
                        // entry of column c: block c >> 3, lane quad-index (c & 7) >> 1, slot c & 1
                        float *ent = reinterpret_cast<float *>(tile + NC * 64 + (c >> 3) * 64 + ((c & 7) >> 1) * 16) + (c & 1);
                        float fac = delta;
#pragma unroll
                        for (int k2 = pp; k2 < P - 1; ++k2) fac *= 256.f;
                        ent[0] = fac;
                        ent[2] = (float)sum[pp];
                    }
                }
                if (NC < 8 * NB && t == 0 && tok == MT - 1) {             // columns that do not exist: factor 0
#pragma unroll
                    for (int c = NC; c < 8 * NB; ++c) {
                        float *ent = reinterpret_cast<float *>(tile + NC * 64 + (c >> 3) * 64 + ((c & 7) >> 1) * 16) + (c & 1);
                        ent[0] = 0.f; ent[2] = 0.f;
                    }
                }
            }
        }
    }
}

// y tile = sum of the partial tiles of all contributors (fixed order) + bias; `nthreads` threads, this one is `tid`
template <int MT>
__device__ __noinline__ void reduce_tile(const float *__restrict__ partial, float *__restrict__ y, const float *__restrict__ bias, uint32_t M, uint32_t N,
                                         uint32_t cnt, const uint32_t *slots, uint32_t nt, int tid, int nthreads) {
    for (int e4 = tid; e4 < MT * 32; e4 += nthreads) {                 // 4 consecutive columns of one token
        const int tok = e4 >> 5, nl = (e4 & 31) * 4;
        const uint32_t n = nt * 128 + nl;
        if ((uint32_t)tok >= M || n >= N) continue;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (uint32_t c = 0; c < cnt; ++c) {
            const float4 pv = __ldcg(reinterpret_cast<const float4 *>(partial + (size_t)slots[c] * (MT * 128)) + e4);
            v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
        }
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (n + i < N) y[(size_t)tok * N + n + i] = vv[i] + (bias ? __ldg(bias + n + i) : 0.f);
    }
}

// XR: consumers + epilogue warp convert the CTA's k-segment of x to digit columns in shared memory (the producers are
// already streaming weights meanwhile), then meet at kBarXReady.  Inlined straight after the dependency wait: what this
// code costs there is mostly the non-sequential control transfers of its first, instruction-cache-cold execution
// (a call + a pass loop around it cost 1.3 us per launch).
template <int CB, int MT>
__device__ __forceinline__ void prepare_x_slice(const float *__restrict__ x, uint32_t M, uint32_t K, uint32_t kb0, uint32_t n_kb, uint8_t *xs,
                                             int tid, int nthreads, unsigned long long *xtrace) {
    prepare_x_tiles<CB, MT, 1>(x, M, K, kb0, n_kb, xs, (uint32_t)tid, (uint32_t)nthreads, xtrace);
#ifdef DLLM_GEMV_TRACE
    if (tid == 0 && xtrace) { unsigned long long _t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(_t)); xtrace[26] = _t; }
#endif
    named_bar_sync(kBarXReady, nthreads);
}

// XR: the CTA's slice of the activations is resident in shared memory (small M·K); otherwise every stage carries the
// activations of its k-blocks (they come from L2: the prepared x is at most a few hundred KB)
template <int CB, int MT, int NG, int NP, int KBS, bool XR>
__global__ void __launch_bounds__((NG * kGroupWarps + NP + 1) * 32, 1)
gemv_mma_kernel(const GemvArgs a) {
    constexpr int NB = gemv_nb(MT), NC = gemv_cols(MT), NPART = gemv_parts(MT);
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
    // verdict: [0] this CTA arrived last at its final tile  [1] contributors  [2] the other contributors' partial tiles of the
    // final tile were prefetched into `pre`  [3] the consumers have reached the final flush  [4] this CTA's place among the contributors
    uint32_t *verdict = reinterpret_cast<uint32_t *>(xfull + 1);
    uint32_t *slots = verdict + 8;                                   // partial slots of the contributors of one tile
    float *pre = reinterpret_cast<float *>(smem + a.pre_off);        // [kPreTiles][MT * 128]

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
        verdict[2] = 0; verdict[3] = 0;
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
        bool x_ready = false;
        const uint64_t w_policy = l2_evict_first_policy();
        while (iter.next(item)) {
            const uint32_t it1 = it0 + item_stages<KBS>(item);
            const uint32_t kbA = item.kb0 & ~(uint32_t)(KBS - 1);
            const uint8_t *wsrc = a.packed + (size_t)item.nt * a.k_blocks * kWBytes;
            const uint2 *psrc = a.gparams + (size_t)item.nt * 128;
            for (; my_it < it1; my_it += NP) {
                // the stage's k-blocks [kb, kb + nk) and the slot of the first one
                const uint32_t sbase = kbA + (my_it - it0) * KBS;
                const uint32_t kb = sbase > item.kb0 ? sbase : item.kb0;
                const uint32_t nk = (sbase + KBS < item.kb1 ? sbase + KBS : item.kb1) - kb;
                const uint32_t sl0 = kb - sbase;
                // only `prefill` stages are requested before the activations are in shared memory: the loads of the
                // activation preparation would otherwise queue behind a whole ring of bulk copies on this SM's memory path
                if (XR && !x_ready && my_it >= a.prefill) { mbar_wait_relaxed(xfull, 0, 1000); x_ready = true; }
                mbar_wait_relaxed(empty + my_s, my_ph ^ 1, 2000);
                if (lane == 0) STRACE(0, my_it);
                uint8_t *st = ring + (size_t)my_s * kStage;
                // The stage's code tiles are contiguous in HBM: one cp.async.bulk (complete_tx on the stage's mbarrier).
                // Scales / zero-points (and the activations, when they are not resident) ride on 16-byte cp.async of
                // the warp's 32 lanes; each lane's arrival on the same mbarrier fires when its copies have landed.
                // (Measured on 14336^2 4-bit: bulk 33.2 us, all-cp.async 34.6 us — the latter costs the producer
                // warps 16 more issue slots per lane and stage.)
                const uint8_t *src = wsrc + (size_t)kb * kWBytes;
                if (a.bulk) {
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full + my_s, nk * kWBytes);
                        bulk_load(st + sl0 * kWBytes, src, nk * kWBytes, full + my_s, w_policy);
                    }
                    __syncwarp();
                } else {
#pragma unroll 8
                    for (uint32_t i = lane; i < nk * (kWBytes / 16); i += 32) cp_async16(st + sl0 * kWBytes + i * 16, src + (size_t)i * 16);
                    if (lane == 0) mbar_arrive_addr(smem_u32(full + my_s));
                }
                // parameters: fetched into the slot of the first k-block of every quantization group the stage touches
                // (the consumers walk the same way and read that slot for the group's other k-blocks)
                uint32_t prev_g = 0xffffffffu;
#pragma unroll
                for (int sub = 0; sub < KBS; ++sub) {
                    if ((uint32_t)sub >= sl0 && (uint32_t)sub < sl0 + nk) {
                        const uint32_t g = kb_group(sbase + sub, a.group_magic);
                        if (g != prev_g) {
                            prev_g = g;
                            const uint2 *pg = psrc + (size_t)g * a.Npad;
                            cp_async16(st + KBS * kWBytes + sub * 1024 + lane * 16, pg + lane * 2);
                            cp_async16(st + KBS * kWBytes + sub * 1024 + 512 + lane * 16, pg + 64 + lane * 2);
                        }
                    }
                }
                if (!XR) {
                    const uint8_t *xsrc = a.xb + (size_t)kb * kXTile;
                    uint8_t *xdst = st + KBS * (kWBytes + 1024) + sl0 * kXTile;
#pragma unroll 4
                    for (uint32_t i = lane; i < nk * (kXTile / 16); i += 32) cp_async16(xdst + i * 16, xsrc + (size_t)i * 16);
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
        const uint32_t col = g < NC ? g : NC - 1;                             // (only 1-2 tokens have fewer than 8 columns)
        uint32_t xs_lane = (XR ? smem_u32(xs) : KBS * (kWBytes + 1024) - (r0 * 4 + t) * 4) + col * 64 + 16 * t;   // !XR: relative to cwa
        uint32_t fa = smem_u32(full) + grp * 8;                               // its full barrier; empty = + 8 stages
        const uint32_t ring_bytes = a.stages * kStage, bar_bytes = a.stages * 8;
        uint32_t dxf = NC * 64 - col * 64;                                    // the lane's factor entry, relative to its activation unit
        asm volatile("" : "+r"(cwa), "+r"(dpw), "+r"(xs_lane), "+r"(fa), "+r"(dxf));
        if (ctid == 0) GTRACE(1);
        // activations, partial-tile workspace, tickets and y belong to the stream's previous kernels until they are done
        // (the iterator's 64-bit divisions are done before the wait: nothing of it depends on the previous kernel)
        uint32_t n_item = 0, last_nt = 0;
        bool last_whole = true;
        ItemIter iter(rg);
        Item item;
        bool have_item = iter.next(item);
        pdl_wait();
        if (ctid == 0) GTRACE(29);
        if (XR) prepare_x_slice<CB, MT>(a.x, a.M, a.K, rg.kb_s0, rg.u0 < rg.u1 ? rg.kbs : 0u, xs, ctid, kConsumers + 32, a.trace ? a.trace + blockIdx.x * 32 : nullptr);
        if (XR && ctid == 0) mbar_arrive_addr(smem_u32(xfull));
        if (ctid == 0) GTRACE(2);
        // stage counter `it` of the CTA; this group owns the stages with it % NG == grp and walks only those
        uint32_t it0 = 0, my_it = (uint32_t)grp, my_s = (uint32_t)grp, my_ph = 0;
        for (; have_item; have_item = iter.next(item)) {
            const uint32_t it1 = it0 + item_stages<KBS>(item);
            const uint32_t kbA = item.kb0 & ~(uint32_t)(KBS - 1);
            for (; my_it < it1; my_it += NG) {
                const uint32_t c = cwa + mbar_wait_token(fa, my_ph);
                if (ctid == 0 && my_it == 0) GTRACE(3);
                if (w == 0 && lane == 0) STRACE(2, my_it);
                // the stage holds the k-blocks [lo, hi) of the aligned group starting at sbase, each in its own slot
                const uint32_t sbase = kbA + (my_it - it0) * KBS;
                const uint32_t lo = sbase > item.kb0 ? sbase : item.kb0, hi = sbase + KBS < item.kb1 ? sbase + KBS : item.kb1;
                const uint32_t xk = XR ? xs_lane + (sbase - rg.kb_s0) * kXTile : c + xs_lane;      // slot 0's activations (may lie before the slice: never read then)
                if (a.fast && sbase >= item.kb0 && sbase + KBS <= item.kb1) {
                    // a full stage of a weight whose aligned pairs share their parameters (every stage but an item's first and
                    // last): straight-line code.  fast == 1: the whole stage is one quantization group (parameters in slot 0),
                    // fast == 2: one group per pair (slot of the pair's first k-block).
#pragma unroll
                    for (int j = 0; j < KBS; j += 2)
                        consume_pair<CB, MT>(c + j * kWBytes, kWBytes, c + dpw + (a.fast == 1 ? 0u : (uint32_t)j * 1024u), xk + j * kXTile, xk + j * kXTile + dxf, ya);
                } else {
                uint32_t prev_g = 0xffffffffu, pslot = 0;
#pragma unroll 1
                for (uint32_t j = 0; j < (uint32_t)KBS; j += 2) {
                    const uint32_t ka = sbase + j;
                    const bool pa = ka >= lo && ka < hi, pb = KBS > 1 && ka + 1 >= lo && ka + 1 < hi;
                    const uint32_t ga = kb_group(ka, a.group_magic), gb = kb_group(ka + 1, a.group_magic);
                    if (pa && pb && ga == gb) {
                        if (ga != prev_g) { prev_g = ga; pslot = j; }
                        consume_pair<CB, MT>(c + j * kWBytes, kWBytes, c + dpw + pslot * 1024, xk + j * kXTile, xk + j * kXTile + dxf, ya);
                    } else {
#pragma unroll 1
                        for (uint32_t h = 0; h < 2; ++h) {
                            if (!(h ? pb : pa)) continue;
                            const uint32_t g = h ? gb : ga;
                            if (g != prev_g) { prev_g = g; pslot = j + h; }
                            consume_kblock<CB, MT>(c + (j + h) * kWBytes, c + dpw + pslot * 1024, xk + (j + h) * kXTile, xk + (j + h) * kXTile + dxf, ya);
                        }
                    }
                }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(fa + bar_bytes);
                if (w == 0 && lane == 0) STRACE(3, my_it);
                cwa += NG * kStage; fa += NG * 8; my_s += NG;
                if (my_s >= a.stages) { my_s -= a.stages; cwa -= ring_bytes; fa -= bar_bytes; my_ph ^= 1; }
            }
            it0 = it1;
            if (ctid == 0) GTRACE(4 + 2 * n_item);
            // ---------- tile (or tile part) done: sum the consumer groups, store, hand over to the epilogue warp ----------
            // the digit columns of a token are summed here (fixed order), then the groups through shared memory
            float *mine = red + (size_t)grp * (MT * kRedStride);
            if (NPART == 2) {
                // columns 2 t, 2 t + 1 of block nb = the two digits of token 4 nb + t
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const int tok = 4 * nb + t;
                    mine[tok * kRedStride + 16 * w + g] = ya[nb][0] + ya[nb][1];
                    mine[tok * kRedStride + 16 * w + g + 8] = ya[nb][2] + ya[nb][3];
                }
            } else {
                // columns 0..2 = token 0, 3..5 = token 1: lane t = 0 holds (0, 1), t = 1 (2, 3), t = 2 (4, 5)
                const int src = (lane & ~3) + 1;
                const float m00 = __shfl_sync(0xffffffffu, ya[0][0], src), m01 = __shfl_sync(0xffffffffu, ya[0][1], src);
                const float m10 = __shfl_sync(0xffffffffu, ya[0][2], src), m11 = __shfl_sync(0xffffffffu, ya[0][3], src);
                if (t == 0) {
                    mine[16 * w + g] = (ya[0][0] + ya[0][1]) + m00;
                    mine[16 * w + g + 8] = (ya[0][2] + ya[0][3]) + m10;
                } else if (t == 2 && MT > 1) {
                    mine[kRedStride + 16 * w + g] = m01 + (ya[0][0] + ya[0][1]);
                    mine[kRedStride + 16 * w + g + 8] = m11 + (ya[0][2] + ya[0][3]);
                }
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) ya[nb][0] = ya[nb][1] = ya[nb][2] = ya[nb][3] = 0.f;
            named_bar_sync(kBarConsumers, kConsumers);
            const bool whole = a.S == 1 && item.kb0 == rg.kb_s0 && item.kb1 == rg.kb_s0 + rg.kbs;
            // The CTA's final tile part: the other contributors have usually published theirs long ago (they were at the
            // start of their ranges), and the epilogue warp has then already copied their partial tiles into shared memory:
            // the tile is finished right here, in the same fixed order, without a round trip through global memory.
            const bool final_part = !whole && iter.u >= iter.u1;
            bool prefetched = false;
            if (final_part) {
                if (ctid == 0) *reinterpret_cast<volatile uint32_t *>(verdict + 3) = 1;
                named_bar_sync(kBarEpiFree, kConsumers + 32);            // the epilogue warp's prefetch attempt is over
                prefetched = *reinterpret_cast<volatile uint32_t *>(verdict + 2) != 0;
            }
            float *dst = whole ? nullptr : a.partial + ((size_t)blockIdx.x * a.max_items + item.ordinal) * (MT * 128);
            for (int e = ctid; e < MT * 128; e += kConsumers) {
                const int tok = e >> 7, nl = e & 127;
                float v = 0.f;
#pragma unroll
                for (int q = 0; q < NG; ++q) v += red[(size_t)q * (MT * kRedStride) + tok * kRedStride + nl];
                const uint32_t n = item.nt * 128 + nl;
                if (whole) {
                    if ((uint32_t)tok < a.M && n < a.N) a.y[(size_t)tok * a.N + n] = v + (a.bias ? __ldg(a.bias + n) : 0.f);
                } else if (prefetched) {
                    const uint32_t cnt = verdict[1], me = verdict[4];
                    float acc = 0.f;                                    // the same order as reduce_tile: bit-identical results
#pragma unroll 1
                    for (uint32_t c = 0; c < cnt; ++c) acc += c == me ? v : pre[(c < me ? c : c - 1) * (MT * 128) + e];
                    if ((uint32_t)tok < a.M && n < a.N) a.y[(size_t)tok * a.N + n] = acc + (a.bias ? __ldg(a.bias + n) : 0.f);
                } else {
                    dst[e] = v;
                }
            }
            named_bar_sync(kBarConsumers, kConsumers);                  // `red` may be overwritten by the next tile
            // the partial tile is in global memory: the epilogue warp publishes it (fence + ticket) and, if this CTA
            // was the last contributor, reduces the tile — while the consumers are already on the next tile
            if (prefetched) {
                if (ctid == 0) a.tickets[item.nt] = 0;                  // the others' arrivals: re-arm for the next launch
            } else if (!whole) {
                if (!final_part) named_bar_sync(kBarEpiFree, kConsumers + 32);
                named_bar_arrive(kBarPartial, kConsumers + 32);
            }
            if (ctid == 0) GTRACE(5 + 2 * n_item);
            ++n_item;
            last_whole = whole || prefetched;
            last_nt = item.nt;
        }
        if (n_item != 0 && !last_whole) {
            // the CTA's last tile: wait for the verdict; if this CTA arrived last, all consumer threads reduce it
            named_bar_sync(kBarFinal, kConsumers + 32);
            if (ctid == 0) GTRACE(30);
            if (verdict[0]) reduce_tile<MT>(a.partial, a.y, a.bias, a.M, a.N, verdict[1], slots, last_nt, ctid, kConsumers);
        }
        if (ctid == 0) GTRACE(31);
    } else {
        // ===================== epilogue warp =====================
        pdl_wait();
        if (XR) prepare_x_slice<CB, MT>(a.x, a.M, a.K, rg.kb_s0, rg.u0 < rg.u1 ? rg.kbs : 0u, xs, kConsumers + lane, kConsumers + 32, nullptr);
        ItemIter iter(rg);
        Item item;
        bool more = iter.next(item);
        while (more) {
            const Item cur = item;
            more = iter.next(item);
            const bool whole = a.S == 1 && cur.kb0 == rg.kb_s0 && cur.kb1 == rg.kb_s0 + rg.kbs;
            if (whole) continue;
            if (more) named_bar_arrive(kBarEpiFree, kConsumers + 32);
            uint32_t cnt = 0, last = 0, me = 0;
            if (lane == 0) {
                // contributors of this tile, in the fixed order (segment, CTA): their partial slots.  (Computed while the
                // consumers are still working on the tile: the 64-bit divisions are off the critical path.)
                for (uint32_t sg = 0; sg < a.S; ++sg) {
                    const uint32_t k0 = a.k_blocks * sg / a.S;
                    const uint32_t kn = a.k_blocks * (sg + 1) / a.S - k0;
                    const uint32_t U = a.n_tiles * kn, a0 = cur.nt * kn, a1 = a0 + kn;
                    const uint32_t j_lo = ((a0 + 1) * a.P - 1) / U, j_hi = (a1 * a.P - 1) / U;
                    for (uint32_t j = j_lo; j <= j_hi; ++j) {
                        const uint32_t first_nt = (U * j / a.P) / kn;
                        if (cnt < kMaxContrib) slots[cnt] = (sg * a.P + j) * a.max_items + (cur.nt - first_nt);
                        if (sg * a.P + j == blockIdx.x) me = cnt;
                        ++cnt;
                    }
                }
            }
            if (!more) {
                // the CTA's final tile part: once every other contributor has published (ticket == contributors - 1), copy
                // their partial tiles into shared memory — while the consumers are still streaming.  Gives up as soon as
                // the consumers arrive at the flush; they then take the ordinary path below.
                cnt = __shfl_sync(0xffffffffu, cnt, 0);
                me = __shfl_sync(0xffffffffu, me, 0);
                bool ok = false;
                if (cnt >= 2 && cnt - 1 <= (uint32_t)kPreTiles) {
                    for (;;) {
                        uint32_t tk = 0;
                        if (lane == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(tk) : "l"(a.tickets + cur.nt) : "memory");
                        tk = __shfl_sync(0xffffffffu, tk, 0);
                        if (tk == cnt - 1) {
                            uint32_t pi = 0;
                            for (uint32_t c = 0; c < cnt; ++c) {
                                if (c == me) continue;
                                const float4 *src = reinterpret_cast<const float4 *>(a.partial + (size_t)slots[c] * (MT * 128));
                                for (int e4 = lane; e4 < MT * 32; e4 += 32) reinterpret_cast<float4 *>(pre)[pi * (MT * 32) + e4] = __ldcg(src + e4);
                                ++pi;
                            }
                            ok = true;
                            break;
                        }
                        if (*reinterpret_cast<volatile uint32_t *>(verdict + 3)) break;
                        __nanosleep(200);
                    }
                }
                if (lane == 0) { verdict[1] = cnt; verdict[4] = me; *reinterpret_cast<volatile uint32_t *>(verdict + 2) = ok ? 1u : 0u; }
                __syncwarp();
                __threadfence_block();
                named_bar_arrive(kBarEpiFree, kConsumers + 32);
                if (ok) break;                                          // the consumers finish the tile
            }
            named_bar_sync(kBarPartial, kConsumers + 32);
            if (lane == 0) {
                __threadfence();                       // cumulative: publishes the consumers' stores ordered before the barrier
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
                reduce_tile<MT>(a.partial, a.y, a.bias, a.M, a.N, cnt, slots, cur.nt, lane, 32);
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

template <int CB, int MT, bool XR, int KBS>
int32_t launch_gemv_mma(dllm_ctx *ctx, const dllm_qweight *qw, const float *x, size_t M, float *y) {
    // NP == NG: a ring slot must always be filled by the same producer warp and drained by the same consumer
    // group (the stage count is a multiple of both) — parity waits of different warps on one slot could alias
    constexpr int NG = 3, NP = 3;
    static_assert(NG % NP == 0 || NP % NG == 0, "ring depth is a multiple of max(NG, NP) only");
    constexpr int kWBytes = WL_TILE_N * WL_TILE_K * CB / 8;
    constexpr int kXTile = gemv_x_tile_bytes(MT);
    constexpr int kStage = KBS * (kWBytes + 1024 + (XR ? 0 : kXTile));
    const uint32_t k_blocks = (uint32_t)qw->k_blocks, n_tiles = (uint32_t)qw->n_tiles;
    const uint32_t sms = (uint32_t)ctx->sm_count;

    // XR: the whole activation block stays resident (one k segment).  The segment machinery (S > 1: a CTA keeps only
    // its k segment's slice) remains for activations that are too large for that but were asked to stay resident.
    const uint32_t red_bytes = NG * MT * kRedStride * 4, pre_bytes = kPreTiles * MT * 512;
    uint32_t S = 1;
    if (XR) {
        for (;; ++S) {
            const uint64_t xb = (uint64_t)((k_blocks + S - 1) / S) * kXTile;
            if (S >= k_blocks || (xb <= (uint64_t)kXBudget && xb + red_bytes + pre_bytes + 1280 + 8ull * kStage <= (uint64_t)kSmemBudget)) break;
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
    if (k_blocks >= 65536 || (uint64_t)n_tiles * k_blocks * (kMaxContrib + 1) >= (1ull << 32)) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "GEMV path: weight too large");
    a.group_magic = group_kb == 1 ? 0u : (uint32_t)(((1ull << 32) + group_kb - 1) / group_kb);   // 0: one k-block per group
    a.S = S; a.P = P;
    a.fast = (qw->per_tensor || group_kb % KBS == 0) ? 1u : (group_kb % 2 == 0 ? 2u : 0u);
    static const uint32_t bulk_mode = getenv("DLLM_GEMV_BULK") ? (uint32_t)atoi(getenv("DLLM_GEMV_BULK")) : 1u;
    a.bulk = bulk_mode;
    a.trace = nullptr;
#ifdef DLLM_GEMV_TRACE
    static const bool trace_chain = getenv("DLLM_GEMV_TRACE_CHAIN") != nullptr;     // one slot per launch, dumped by dllm_debug_gemv_trace_dump
    const size_t trace_stride = (size_t)grid * 32 + 1024;
    if (trace_chain) {
        if (!g_trace_buf) { cudaMalloc(&g_trace_buf, kTraceSlots * trace_stride * 8); cudaMemset(g_trace_buf, 0, kTraceSlots * trace_stride * 8); }
        g_trace_stride = trace_stride; g_trace_grid = grid;
        a.trace = g_trace_buf + (size_t)(g_trace_launch++ % kTraceSlots) * trace_stride;
    } else {
        DLLM_TRY(ensure_buf(ctx, ctx->lin_flags, trace_stride * 8));
        a.trace = (unsigned long long *)ctx->lin_flags.p;
        cudaMemsetAsync(a.trace, 0, trace_stride * 8, ctx->stream);
    }
#endif
    const uint32_t kbs_max = (k_blocks + S - 1) / S, kbs_min = k_blocks / S;
    const uint64_t len_max = ((uint64_t)n_tiles * kbs_max + P - 1) / P;
    a.max_items = (uint32_t)(len_max / (kbs_min ? kbs_min : 1)) + 2;

    const uint32_t xbytes = XR ? kbs_max * kXTile : 0u;
    uint32_t stages = (uint32_t)((kSmemBudget - xbytes - red_bytes - pre_bytes - 1280) / kStage);
    if (stages > (uint32_t)kMaxStages) stages = kMaxStages;
    stages -= stages % (NG > NP ? NG : NP);
    if (stages < 2 * NG || stages <= (uint32_t)NP) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "GEMV path: shared-memory ring too small");
    a.stages = stages;
    static const int prefill_env = getenv("DLLM_GEMV_PREFILL") ? atoi(getenv("DLLM_GEMV_PREFILL")) : -1;   // experiments only
    a.prefill = prefill_env >= 0 ? (uint32_t)prefill_env : stages;
    a.x_off = stages * kStage;
    a.red_off = a.x_off + ((xbytes + 127) & ~127u);
    a.bar_off = a.red_off + ((red_bytes + 127) & ~127u);
    a.pre_off = (a.bar_off + (2 * stages + 1) * 8 + (8 + kMaxContrib) * 4 + 15) & ~15u;
    const size_t smem_bytes = a.pre_off + pre_bytes;

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

    {   // per (context's device, kernel instance): function attributes are per device
        const void *key = reinterpret_cast<const void *>(gemv_xprep_kernel<CB, MT>);
        bool done = false;
        for (const void *k : ctx->smem_attr_done) done = done || k == key;
        DLLM_TRY(ensure_smem_attr(ctx, gemv_mma_kernel<CB, MT, NG, NP, KBS, XR>, 227 * 1024));
        if (!done) {
            // the activation-prep kernel runs right before: same shared-memory carve-out, so the SMs are not
            // re-partitioned (L1 vs shared) between the two launches
            CUDA_TRY(ctx, cudaFuncSetAttribute(gemv_xprep_kernel<CB, MT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            ctx->smem_attr_done.push_back(key);
        }
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
    if (!trace_chain) {   // dump the per-CTA timeline (timing experiments only)
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

// k-blocks per ring stage.  A stage costs the same hand-offs (two mbarrier round trips, a bulk copy, the consumers' wait)
// whatever it carries, so narrow codes take four k-blocks per stage where the ring stays >= 9 stages deep
// (DLLM_GEMV_KBS = 2 / 4 overrides: experiments only)
template <int CB, int MT, bool XR>
int32_t launch_gemv_kbs(dllm_ctx *ctx, const dllm_qweight *qw, const float *x, size_t M, float *y) {
    static const int kbs_env = getenv("DLLM_GEMV_KBS") ? atoi(getenv("DLLM_GEMV_KBS")) : 0;
    constexpr int kWBytes = WL_TILE_N * WL_TILE_K * CB / 8, kXTile = gemv_x_tile_bytes(MT);
    constexpr int kStage4 = 4 * (kWBytes + 1024 + (XR ? 0 : kXTile));
    const size_t fixed = (XR ? qw->k_blocks * (size_t)kXTile : 0) + 3 * MT * kRedStride * 4 + kPreTiles * MT * 512 + 1280;
    const size_t stages4 = fixed < (size_t)kSmemBudget ? (kSmemBudget - fixed) / kStage4 : 0;
    const bool want4 = kbs_env ? kbs_env == 4 : CB == 2;
    if (want4 && stages4 >= 9 && qw->k_blocks >= 8) return launch_gemv_mma<CB, MT, XR, 4>(ctx, qw, x, M, y);
    return launch_gemv_mma<CB, MT, XR, 2>(ctx, qw, x, M, y);
}

template <int CB, int MT>
int32_t launch_gemv_x(dllm_ctx *ctx, const dllm_qweight *qw, const float *x, size_t M, float *y) {
    static const bool no_xr = getenv("DLLM_GEMV_XR") && atoi(getenv("DLLM_GEMV_XR")) == 0;      // experiments only
    if (!no_xr && qw->k_blocks * (size_t)gemv_x_tile_bytes(MT) <= kXResident) return launch_gemv_kbs<CB, MT, true>(ctx, qw, x, M, y);
    return launch_gemv_kbs<CB, MT, false>(ctx, qw, x, M, y);
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

#ifdef DLLM_GEMV_TRACE
// timing experiments only: the per-CTA stamps of the last launches (one slot per launch; a replayed CUDA graph
// overwrites the slots its launches were captured with), absolute globaltimer ns
extern "C" __attribute__((visibility("default"))) int dllm_debug_gemv_trace_dump(const char *path) {
    if (!g_trace_buf) return -1;
    cudaDeviceSynchronize();
    std::vector<unsigned long long> h(kTraceSlots * g_trace_stride);
    cudaMemcpy(h.data(), g_trace_buf, h.size() * 8, cudaMemcpyDeviceToHost);
    FILE *f = fopen(path, "w");
    if (!f) return -2;
    fprintf(f, "slot,cta,start,cons_start,x_ready,first_full,s4,s5,s6,s7,s8,s9,xl_issued,xl_landed,x_done,pdl,verdict,end,last\n");
    const size_t n = g_trace_launch < (size_t)kTraceSlots ? g_trace_launch : (size_t)kTraceSlots;
    for (size_t sl = 0; sl < n; ++sl)
        for (uint32_t c = 0; c < g_trace_grid; ++c) {
            const unsigned long long *r = h.data() + sl * g_trace_stride + (size_t)c * 32;
            unsigned long long last = 0;
            for (int i = 0; i < 32; ++i) if (r[i] > last) last = r[i];
            fprintf(f, "%zu,%u", sl, c);
            for (int i = 0; i < 10; ++i) fprintf(f, ",%llu", r[i]);
            unsigned long long lastc = 0;
            for (int i = 0; i < 24; ++i) if (r[i] > lastc) lastc = r[i];
            fprintf(f, ",%llu,%llu,%llu,%llu,%llu,%llu,%llu\n", r[24], r[25], r[26], r[29], r[30], r[31], lastc);
        }
    fclose(f);
    return 0;
}
#endif

bool k_gemv_supported(const dllm_qweight *qw, size_t M) {
    // integer zero-points (subtracted exactly in the integer domain)
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
