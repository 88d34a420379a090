// tp.cu — tensor-parallel plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch, one
// collective per column->row linear pair at the layer boundary.  The
// communicator is created from a 128-byte unique id that the launcher broadcasts
// (torch.distributed / MPI / a file), so this library needs no process manager of its own.
#include <nccl.h>
#include <cuda_bf16.h>
#include <vector>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

#define NCCL_TRY(ctx, expr)                                                                   \
    do {                                                                                      \
        ncclResult_t _r = (expr);                                                             \
        if (_r != ncclSuccess) {                                                              \
            DLLM_SET_ERR(ctx, "NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(_r)); \
            return DLLM_ERR_NCCL;                                                             \
        }                                                                                     \
    } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");

// ==========================================================================================
// Peer-to-peer all-reduce over NVLink / NVSwitch (this library's own kernel, no NCCL on the data path)
// ==========================================================================================
// Every rank allocates one arena of the same size and maps all the others' (CUDA IPC; the 64-byte handles travel over the
// NCCL communicator once, at set-up).  A tensor at byte offset `off` of the arena is reduced in place, "two-shot":
//   barrier  — every rank's partial sums are complete (a rank's kernel follows its GEMM in stream order);
//   reduce   — rank r owns the r-th 1/W of the tensor: it loads that slice from all W arenas (own HBM + W-1 peer loads over
//              NVLink), sums in f32 in rank order, and stores the result into all W arenas (W-1 peer stores);
//   barrier  — every slice has landed everywhere.
// Per GPU and direction that is (W-1)/W of the tensor on NVLink, each way — the same bytes as a ring or an in-switch reduction
// needs — in one pass with nothing staged.  Every element is summed by exactly one rank, so all ranks hold identical bits.
// Barriers are per block: block b of rank r raises flag [phase][r][b] in every rank's flag block (st.release.sys) and waits
// for the W flags [phase][*][b] of its own (ld.acquire.sys on local memory); the generation counter makes the flags
// reusable without clearing.  A wait that lasts longer than ~4 s (a peer died) gives up and raises ctx->p2p_err instead of
// hanging the GPU.  The grid must be co-resident (it is: <= 128 blocks of 512 threads, no shared memory, and while it
// overlaps a dense kernel the latter leaves `sm_reserve` SMs free) and identical on all ranks (same code path).
namespace {
constexpr int kP2PMaxBlocks = 256;
constexpr int kP2PMaxWorld = 16;
constexpr size_t kP2PBarrierBytes = 2 * kP2PMaxWorld * kP2PMaxBlocks * sizeof(uint32_t);
constexpr int kP2PGateSub = 8;                                 // arrival counters per source rank (sub-slices of its token slice)
constexpr size_t kP2PFlagBytes = kP2PBarrierBytes + kP2PMaxWorld * kP2PGateSub * sizeof(uint32_t);   // + the gated all-gather's counters

struct P2PArgs {
    unsigned char *base[kP2PMaxWorld];     // every rank's arena
    size_t off;                            // byte offset of the tensor
    size_t n16;                            // 16-byte vectors
    size_t flag_off;                       // byte offset of the flag block
    uint32_t epoch;                        // barrier generation of the first barrier (second: + 1)
    int rank, world;
    unsigned int *err;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Data accesses of the exchange: WEAK 16-byte loads / stores.  Ordering comes from the barriers
// alone — a rank's stores precede its st.release.sys of the flag (fence + cumulativity through bar.sync), and loads follow the
// ld.acquire.sys that observed the peers' flags (which also invalidates L1).  (Strong .sys accesses per 16 bytes — the first
// version — are each individually coherent: LDG/STG.STRONG.SYS.)
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void *p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(void *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void p2p_barrier(const P2PArgs &a, int phase, uint32_t gen) {
    __syncthreads();                                   // the block's stores precede thread t's fence below (causality through the barrier)
    if ((int)threadIdx.x < a.world) {
        __threadfence_system();
        const int peer = threadIdx.x;
        uint32_t *theirs = reinterpret_cast<uint32_t *>(a.base[peer] + a.flag_off) + ((size_t)phase * kP2PMaxWorld + a.rank) * kP2PMaxBlocks + blockIdx.x;
        st_release_sys(theirs, gen);
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(a.base[a.rank] + a.flag_off) + ((size_t)phase * kP2PMaxWorld + peer) * kP2PMaxBlocks + blockIdx.x;
        unsigned long long t0 = 0;
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(mine) - gen) < 0) {
            if ((++spins & 0x3ff) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) { atomicExch(a.err, 1u + (unsigned)peer); break; }
            }
        }
    }
    __syncthreads();
}

template <bool BF16>
__device__ __forceinline__ void p2p_accumulate(float *acc, uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (BF16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {                  // bf16 -> f32 is a 16-bit shift
            acc[2 * i] += __uint_as_float(w[i] << 16);
            acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += __uint_as_float(w[i]);
    }
}

// W = slots of the rank loop (>= world), U = vectors per thread and pass: U x W 16-byte loads are in flight per thread before
// the first add — what a handful of SMs needs to keep NVLink busy while the rest of the GPU runs a GEMM
template <bool BF16, int W, int U>
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(const P2PArgs a) {
    p2p_barrier(a, 0, a.epoch);
    const size_t per = (a.n16 + a.world - 1) / a.world;
    const size_t i0 = per * a.rank < a.n16 ? per * a.rank : a.n16, i1 = i0 + per < a.n16 ? i0 + per : a.n16;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += U * T) {
        uint4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int r = 0; r < W; ++r)
                if (r < a.world && i + u * T < i1) v[u][r] = ld_relaxed_sys_v4(a.base[r] + a.off + (i + u * T) * 16);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * T >= i1) break;
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
            for (int r = 0; r < W; ++r)
                if (r < a.world) p2p_accumulate<BF16>(acc, v[u][r]);                     // rank order: the same sum on every rank
            uint4 o;
            if (BF16) {
                __nv_bfloat162 h[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
                o = make_uint4(*reinterpret_cast<uint32_t *>(&h[0]), *reinterpret_cast<uint32_t *>(&h[1]),
                               *reinterpret_cast<uint32_t *>(&h[2]), *reinterpret_cast<uint32_t *>(&h[3]));
            } else {
                o = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
            }
#pragma unroll
            for (int r = 0; r < W; ++r)
                if (r < a.world) st_relaxed_sys_v4(a.base[r] + a.off + (i + u * T) * 16, o);
        }
    }
    p2p_barrier(a, 1, a.epoch + 1);
}

// Second half of the fused reduce-scatter + all-gather: the dense kernel of a row-parallel layer has already pushed every
// rank's partial sums for MY tokens into my receive buffer (row block s = rank s's partial, csrc/umma_gemm.cu epilogue); this
// kernel sums the W row blocks (local loads only), and stores the finished rows into every rank's activation buffer (the
// all-gather: W-1 peer stores per element).  NVLink carries (W-1)/W of the tensor per direction here — the other half went out
// under the GEMM.  a.off = receive buffer, a.off2 = destination tensor [M, N], a.n16 = 16-byte vectors of one rank's slice.
// signal = G > 0: no closing barrier.  The slice is processed in G equal sub-slices; after each one every block bumps, on every
// rank, arrival counter [this rank][g] — the consuming GEMM on each rank gates its loads of those token rows on it
// (umma_gemm.cu), so the all-gather runs under that GEMM instead of in front of it, and the GEMM can start on sub-slice 0 while
// the rest is still being reduced.
template <int W, int U>
__global__ void __launch_bounds__(512) p2p_reduce_gather_kernel(const P2PArgs a, size_t off2, int signal) {
    p2p_barrier(a, 0, a.epoch);
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const unsigned char *recv = a.base[a.rank] + a.off;
    const int G = signal > 0 ? signal : 1;
    const size_t part = a.n16 / G;                              // (the host picks G so that it divides the slice's rows)
    for (int g = 0; g < G; ++g) {
        const size_t p0 = (size_t)g * part, p1 = g + 1 == G ? a.n16 : p0 + part;
        for (size_t i = p0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p1; i += U * T) {
            uint4 v[U][W];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < W; ++r)
                    if (r < a.world && i + u * T < p1) v[u][r] = ld_relaxed_sys_v4(recv + ((size_t)r * a.n16 + i + u * T) * 16);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i + u * T >= p1) break;
                float acc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
                for (int r = 0; r < W; ++r)
                    if (r < a.world) p2p_accumulate<true>(acc, v[u][r]);                 // rank order, as in the two-shot kernel
                __nv_bfloat162 h[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
                const uint4 o = make_uint4(*reinterpret_cast<uint32_t *>(&h[0]), *reinterpret_cast<uint32_t *>(&h[1]),
                                           *reinterpret_cast<uint32_t *>(&h[2]), *reinterpret_cast<uint32_t *>(&h[3]));
#pragma unroll
                for (int r = 0; r < W; ++r)
                    if (r < a.world) st_relaxed_sys_v4(a.base[r] + off2 + ((size_t)a.rank * a.n16 + i + u * T) * 16, o);
            }
        }
        if (signal > 0) {
            __syncthreads();
            if ((int)threadIdx.x < a.world) {
                __threadfence_system();
                uint32_t *cnt = reinterpret_cast<uint32_t *>(a.base[threadIdx.x] + a.flag_off + kP2PBarrierBytes) + a.rank * kP2PGateSub + g;
                asm volatile("red.release.sys.global.add.u32 [%0], 1;" :: "l"(cnt) : "memory");
            }
        }
    }
    if (signal <= 0) p2p_barrier(a, 1, a.epoch + 1);
}

// The same reduce / all-gather with the peer stores issued as BULK copies: the finished rows are staged in shared memory
// (2 x 16 KB) and leave as one cp.async.bulk per peer and 16 KB chunk.  16-byte st.global to peer memory travels as 32-byte
// packets (measured 575 GB/s per direction = 64 % of the link's 900 GB/s, the payload share of a 32-byte packet); bulk copies move
// full 128-byte lines.  (DLLM_P2P_BULK=0 selects the plain-store kernel above.)
template <int W>
__global__ void __launch_bounds__(256, 2) p2p_reduce_gather_bulk_kernel(const P2PArgs a, size_t off2) {
    constexpr int U = W <= 4 ? 4 : 2, kChunk16 = 256 * U;      // 16-byte vectors per chunk = 16 KB (8 KB at 8 ranks: registers)
    __shared__ __align__(128) uint4 stage[2][kChunk16];
    p2p_barrier(a, 0, a.epoch);
    const unsigned char *recv = a.base[a.rank] + a.off;
    const size_t n_chunks = (a.n16 + kChunk16 - 1) / kChunk16;
    uint32_t it = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
        const uint32_t b = it & 1;
        // the bulk copies that read this buffer two rounds ago have read it (each issuing thread waits for its own groups)
        if ((int)threadIdx.x < a.world) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        const size_t i0 = c * kChunk16;
        const uint32_t n_here = (uint32_t)(a.n16 - i0 < (size_t)kChunk16 ? a.n16 - i0 : (size_t)kChunk16);
        uint4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int r = 0; r < W; ++r)
                if (r < a.world && u * 256 + threadIdx.x < n_here)
                    v[u][r] = ld_relaxed_sys_v4(recv + ((size_t)r * a.n16 + i0 + u * 256 + threadIdx.x) * 16);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u * 256 + threadIdx.x >= n_here) break;
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
            for (int r = 0; r < W; ++r)
                if (r < a.world) p2p_accumulate<true>(acc, v[u][r]);
            __nv_bfloat162 h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
            stage[b][u * 256 + threadIdx.x] = make_uint4(*reinterpret_cast<uint32_t *>(&h[0]), *reinterpret_cast<uint32_t *>(&h[1]),
                                                          *reinterpret_cast<uint32_t *>(&h[2]), *reinterpret_cast<uint32_t *>(&h[3]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk-copy engine
        __syncthreads();
        if ((int)threadIdx.x < a.world) {
            unsigned char *dst = a.base[threadIdx.x] + off2 + ((size_t)a.rank * a.n16 + i0) * 16;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(&stage[b][0])), "r"(n_here * 16u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if ((int)threadIdx.x < a.world) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // written, not just read
    p2p_barrier(a, 1, a.epoch + 1);
}

template <bool BF16>
void p2p_launch(const P2PArgs &a, int blocks, cudaStream_t stream) {
    if (a.world <= 2) p2p_allreduce_kernel<BF16, 2, 4><<<blocks, 512, 0, stream>>>(a);
    else if (a.world <= 4) p2p_allreduce_kernel<BF16, 4, 2><<<blocks, 512, 0, stream>>>(a);
    else if (a.world <= 8) p2p_allreduce_kernel<BF16, 8, 1><<<blocks, 512, 0, stream>>>(a);
    else p2p_allreduce_kernel<BF16, kP2PMaxWorld, 1><<<blocks, 512, 0, stream>>>(a);
}
}  // namespace

// `buf` lies inside the symmetric arena and is 16-byte aligned with a multiple of 16 bytes: reduce it with the kernel above
static bool p2p_covers(const dllm_ctx *ctx, const void *buf, size_t bytes) {
    if (!ctx->p2p_arena || ctx->tp_world <= 1) return false;
    const uintptr_t b = (uintptr_t)buf, a0 = (uintptr_t)ctx->p2p_arena;
    return b >= a0 && b + bytes <= a0 + ctx->p2p_bytes && ((b - a0) & 15u) == 0 && (bytes & 15u) == 0;
}

static int p2p_blocks(const dllm_ctx *ctx, bool overlapped) {
    // alone on the GPU: one block per SM up to 128; under a dense kernel: two blocks per SM that kernel leaves free
    const int reserve = ctx->sm_limit > 0 ? ctx->sm_count - ctx->sm_limit : 0;
    int blocks = overlapped ? 2 * (reserve > 0 ? reserve : 4) : (ctx->sm_count < 128 ? ctx->sm_count : 128);
    static const int blocks_env = getenv("DLLM_P2P_BLOCKS") ? atoi(getenv("DLLM_P2P_BLOCKS")) : 0;      // experiments only
    if (blocks_env > 0) blocks = blocks_env;
    return blocks > kP2PMaxBlocks ? kP2PMaxBlocks : blocks;
}

static int32_t p2p_allreduce(dllm_ctx *ctx, void *buf, size_t bytes, bool bf16, cudaStream_t stream, bool overlapped) {
    P2PArgs a;
    for (int r = 0; r < kP2PMaxWorld; ++r) a.base[r] = (unsigned char *)(r < ctx->tp_world ? ctx->p2p_peer[r] : nullptr);
    a.off = (size_t)((uintptr_t)buf - (uintptr_t)ctx->p2p_arena);
    a.n16 = bytes / 16;
    a.flag_off = ctx->p2p_bytes;
    a.epoch = ctx->p2p_epoch + 1;
    ctx->p2p_epoch += 2;
    a.rank = ctx->tp_rank; a.world = ctx->tp_world;
    a.err = ctx->p2p_err;
    const int blocks = p2p_blocks(ctx, overlapped);
    if (bf16) p2p_launch<true>(a, blocks, stream);
    else p2p_launch<false>(a, blocks, stream);
    LAUNCH_CHECK(ctx);
    ctx->p2p_calls++;
    return DLLM_OK;
}

// The three regions of the arena a tensor-parallel tcgen05 stack uses: two activation buffers and the receive buffer of the
// fused reduce-scatter (absent when the arena only has room for the first two).
bool tp_p2p_regions(const dllm_ctx *ctx, size_t bytes_each, char **b0, char **b1, char **recv) {
    if (!ctx->p2p_arena || ctx->tp_world <= 1) return false;
    const size_t need = (bytes_each + 255) & ~(size_t)255;
    if (3 * need <= ctx->p2p_bytes) {
        const size_t R = (ctx->p2p_bytes / 3) & ~(size_t)255;
        *b0 = (char *)ctx->p2p_arena; *b1 = *b0 + R; *recv = *b0 + 2 * R;
        return true;
    }
    if (2 * need <= ctx->p2p_bytes) {
        *b0 = (char *)ctx->p2p_arena; *b1 = *b0 + ((ctx->p2p_bytes / 2) & ~(size_t)255); *recv = nullptr;
        return true;
    }
    return false;
}

// every rank's view of the address `local` of this rank's arena
void tp_p2p_peer_ptrs(const dllm_ctx *ctx, const void *local, void **out8) {
    const size_t off = (size_t)((const char *)local - (const char *)ctx->p2p_arena);
    for (int r = 0; r < 8; ++r) out8[r] = r < ctx->tp_world ? (void *)((char *)ctx->p2p_peer[r] + off) : nullptr;
}

// recv (this rank's receive buffer: [world][rows, N] bf16, filled by all ranks' fused epilogues) -> dst [world * rows, N] on every rank
// signal: 0 = closing barrier; 1 = arrival counters, consumer follows on the same stream (its own slice needs no gate);
// 2 = arrival counters with the kernel on the communication stream UNDER the consumer (32 blocks on the 16 SMs the consumer leaves
// free keep NVLink as busy as 128 do: scripts/p2p_blocks_probe.py; the consumer gates its own slice too)
int32_t tp_reduce_gather(dllm_ctx *ctx, const void *recv, void *dst, size_t rows, size_t N, cudaStream_t stream, int signal) {
    const size_t slice = rows * N * 2;
    if (!p2p_covers(ctx, recv, slice * ctx->tp_world) || !p2p_covers(ctx, dst, slice * ctx->tp_world))
        DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "reduce-gather: buffers must lie in the peer-to-peer arena");
    P2PArgs a;
    for (int r = 0; r < kP2PMaxWorld; ++r) a.base[r] = (unsigned char *)(r < ctx->tp_world ? ctx->p2p_peer[r] : nullptr);
    a.off = (size_t)((uintptr_t)recv - (uintptr_t)ctx->p2p_arena);
    const size_t off2 = (size_t)((uintptr_t)dst - (uintptr_t)ctx->p2p_arena);
    a.n16 = slice / 16;
    a.flag_off = ctx->p2p_bytes;
    a.epoch = ctx->p2p_epoch + 1;
    ctx->p2p_epoch += 2;
    a.rank = ctx->tp_rank; a.world = ctx->tp_world;
    a.err = ctx->p2p_err;
    const int blocks = signal == 2 ? 32 : p2p_blocks(ctx, stream != ctx->stream);
    // (bulk-copy peer stores measured no faster than 16-byte stores at 2 and 4 ranks: opt-in, DLLM_P2P_BULK=1)
    static const bool bulk = getenv("DLLM_P2P_BULK") && atoi(getenv("DLLM_P2P_BULK")) == 1;
    int G = 0;
    if (signal) {
        // sub-slices of whole 256-token tiles: 4 where the slice allows
        G = rows % 1024 == 0 ? 4 : rows % 512 == 0 ? 2 : 1;
        // the consumer is armed with the value the counters reach once every block of this launch has signalled
        ctx->gate_signals += (uint32_t)blocks;
        ctx->gate_counters = reinterpret_cast<const uint32_t *>((const char *)ctx->p2p_arena + ctx->p2p_bytes + kP2PBarrierBytes);
        ctx->gate_target = ctx->gate_signals;
        ctx->gate_rows = rows;
        ctx->gate_sub = rows / (size_t)G;
        ctx->gate_self = signal == 2 ? 0xffffffffu : (uint32_t)ctx->tp_rank;
        ctx->gate_armed = true;
    }
    if (bulk && !signal) {
        // 256-thread blocks with <= 32 KB of staging, two per SM; every block must be resident at once (a block waits for the
        // peers' block of the same index): alone on the GPU one per SM, under a dense kernel two per SM that kernel leaves free
        const int reserve = ctx->sm_limit > 0 ? ctx->sm_count - ctx->sm_limit : 0;
        int b2 = stream != ctx->stream ? 2 * (reserve > 0 ? reserve : 4) : ctx->sm_count;
        if (b2 > kP2PMaxBlocks) b2 = kP2PMaxBlocks;
        if (a.world <= 2) p2p_reduce_gather_bulk_kernel<2><<<b2, 256, 0, stream>>>(a, off2);
        else if (a.world <= 4) p2p_reduce_gather_bulk_kernel<4><<<b2, 256, 0, stream>>>(a, off2);
        else p2p_reduce_gather_bulk_kernel<8><<<b2, 256, 0, stream>>>(a, off2);
    } else if (a.world <= 2) p2p_reduce_gather_kernel<2, 4><<<blocks, 512, 0, stream>>>(a, off2, G);
    else if (a.world <= 4) p2p_reduce_gather_kernel<4, 2><<<blocks, 512, 0, stream>>>(a, off2, G);
    else p2p_reduce_gather_kernel<8, 1><<<blocks, 512, 0, stream>>>(a, off2, G);
    LAUNCH_CHECK(ctx);
    ctx->p2p_calls++;
    return DLLM_OK;
}

int32_t tp_allreduce(dllm_ctx *ctx, float *buf, size_t n) {
    if (ctx->tp_world <= 1 || n == 0 || ctx->tp_skip_comm) return DLLM_OK;
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    if (p2p_covers(ctx, buf, n * 4)) return p2p_allreduce(ctx, buf, n * 4, false, ctx->stream, false);
    NCCL_TRY(ctx, ncclAllReduce(buf, buf, n, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return DLLM_OK;
}

// bf16 partial sums of a row-parallel linear, reduced in place (half the NVLink bytes of the f32 form; the result
// feeds the next tcgen05 linear, which reads bf16 anyway)
int32_t tp_allreduce_bf16(dllm_ctx *ctx, void *buf, size_t n) {
    if (ctx->tp_world <= 1 || n == 0 || ctx->tp_skip_comm) return DLLM_OK;
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    if (p2p_covers(ctx, buf, n * 2)) return p2p_allreduce(ctx, buf, n * 2, true, ctx->stream, false);
    NCCL_TRY(ctx, ncclAllReduce(buf, buf, n, ncclBfloat16, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return DLLM_OK;
}

// the same all-reduces on an explicit stream (the overlapped tensor-parallel forward issues them on ctx->comm_stream)
int32_t tp_allreduce_on(dllm_ctx *ctx, void *buf, size_t n, bool bf16, cudaStream_t stream) {
    if (ctx->tp_world <= 1 || n == 0 || ctx->tp_skip_comm) return DLLM_OK;
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    if (p2p_covers(ctx, buf, n * (bf16 ? 2 : 4))) return p2p_allreduce(ctx, buf, n * (bf16 ? 2 : 4), bf16, stream, stream != ctx->stream);
    NCCL_TRY(ctx, ncclAllReduce(buf, buf, n, bf16 ? ncclBfloat16 : ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, stream));
    return DLLM_OK;
}

// params_dev = {scale, zp, min, max} of this rank's shard -> min / max over all ranks of the group (two 1-float all-reduces)
int32_t tp_allreduce_minmax(dllm_ctx *ctx, float *params_dev) {
    if (ctx->tp_world <= 1) return DLLM_OK;
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    NCCL_TRY(ctx, ncclGroupStart());
    NCCL_TRY(ctx, ncclAllReduce(params_dev + 2, params_dev + 2, 1, ncclFloat32, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    NCCL_TRY(ctx, ncclAllReduce(params_dev + 3, params_dev + 3, 1, ncclFloat32, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    NCCL_TRY(ctx, ncclGroupEnd());
    return DLLM_OK;
}

namespace {
// gathered [world][M][n_local] -> out [M][world*n_local]
__global__ void interleave_cols_kernel(const float *__restrict__ gathered, size_t M, size_t n_local, int world,
                                       float *__restrict__ out) {
    const size_t total = (size_t)world * M * n_local;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / (M * n_local), rem = i - r * M * n_local;
        const size_t m = rem / n_local, c = rem - m * n_local;
        out[m * ((size_t)world * n_local) + r * n_local + c] = gathered[i];
    }
}
}  // namespace

int32_t tp_allgather_cols(dllm_ctx *ctx, const float *in, size_t M, size_t n_local, float *out) {
    const size_t n = M * n_local;
    if (ctx->tp_world <= 1) {
        if (n && in != out) CUDA_TRY(ctx, cudaMemcpyAsync(out, in, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        return DLLM_OK;
    }
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    if (n == 0) return DLLM_OK;
    // own scratch: ws[7] is the noise_pred buffer of dllm_denoise_step_dev, i.e. possibly `out` itself
    DLLM_TRY(ensure_buf(ctx, ctx->tp_ws, (size_t)ctx->tp_world * n * sizeof(float)));
    float *gathered = (float *)ctx->tp_ws.p;
    NCCL_TRY(ctx, ncclAllGather(in, gathered, n, ncclFloat32, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    size_t blocks = ((size_t)ctx->tp_world * n + 255) / 256, cap = (size_t)ctx->sm_count * 16;
    interleave_cols_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, ctx->stream>>>(gathered, M, n_local, ctx->tp_world, out);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

extern "C" {

int32_t dllm_tp_unique_id(uint8_t id_out[128]) {
    if (!id_out) return DLLM_ERR_NULL;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return DLLM_ERR_NCCL;
    memcpy(id_out, &id, 128);
    return DLLM_OK;
}

int32_t dllm_tp_init(dllm_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t world) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    if (!id || world < 1 || rank < 0 || rank >= world) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "bad rank %d / world %d", rank, world);
    if (ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "tensor-parallel group already initialised");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclComm_t comm;
    NCCL_TRY(ctx, ncclCommInitRank(&comm, world, uid, rank));
    ctx->nccl_comm = (void *)comm;
    ctx->tp_rank = rank;
    ctx->tp_world = world;
    return DLLM_OK;
}

// tear the arena down: peers' mappings first, then (after everybody has unmapped: the caller's group barrier) the allocation
static void p2p_release(dllm_ctx *ctx) {
    if (!ctx->p2p_arena) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    for (int r = 0; r < ctx->tp_world && r < kP2PMaxWorld; ++r)
        if (r != ctx->tp_rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    if (ctx->nccl_comm) {   // nobody frees while a peer still maps it
        ncclAllReduce(ctx->p2p_err, ctx->p2p_err, 1, ncclUint32, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(ctx->p2p_arena);
    cudaFree(ctx->p2p_err);
    ctx->p2p_arena = nullptr; ctx->p2p_err = nullptr; ctx->p2p_bytes = 0;
    ctx->gate_armed = false; ctx->gate_counters = nullptr; ctx->gate_signals = 0;
    for (auto &q : ctx->p2p_peer) q = nullptr;
    cudaGetLastError();
}

int32_t dllm_tp_p2p_enable(dllm_ctx *ctx, size_t arena_bytes) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    if (ctx->tp_world <= 1) return DLLM_OK;                       // nothing to exchange
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    if (ctx->tp_world > kP2PMaxWorld) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "peer-to-peer all-reduce supports up to %d ranks", kP2PMaxWorld);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    p2p_release(ctx);
    if (arena_bytes == 0) return DLLM_OK;                         // switched off: back to ncclAllReduce
    arena_bytes = (arena_bytes + 4095) & ~(size_t)4095;
    const int W = ctx->tp_world;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    // every rank must end up with the same verdict, so failures are agreed on through the communicator before returning
    void *arena = nullptr;
    unsigned int *err = nullptr;
    unsigned char *d_handles = nullptr;
    uint32_t ok = 1;
    if (cudaMalloc(&arena, arena_bytes + kP2PFlagBytes) != cudaSuccess) ok = 0;
    if (ok && cudaMalloc(&err, 2 * sizeof(unsigned int)) != cudaSuccess) ok = 0;
    if (ok && cudaMalloc(&d_handles, (size_t)W * sizeof(cudaIpcMemHandle_t)) != cudaSuccess) ok = 0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine, arena) != cudaSuccess) ok = 0;
    cudaGetLastError();
    std::vector<cudaIpcMemHandle_t> all((size_t)W);
    if (ok) {
        cudaMemsetAsync(arena, 0, arena_bytes + kP2PFlagBytes, ctx->stream);
        cudaMemsetAsync(err, 0, 2 * sizeof(unsigned int), ctx->stream);
        cudaMemcpyAsync(d_handles + (size_t)ctx->tp_rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream);
    }
    // (a rank that failed above still takes part in the collectives below with ok = 0)
    unsigned int *d_ok = nullptr;
    if (cudaMalloc(&d_ok, sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); DLLM_FAIL(ctx, DLLM_ERR_OOM, "cudaMalloc failed"); }
    auto agree = [&](uint32_t v) -> int {
        cudaMemcpyAsync(d_ok, &v, sizeof(v), cudaMemcpyHostToDevice, ctx->stream);
        if (ncclAllReduce(d_ok, d_ok, 1, ncclUint32, ncclMin, comm, ctx->stream) != ncclSuccess) return -1;
        uint32_t out = 0;
        cudaMemcpyAsync(&out, d_ok, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
        return (int)out;
    };
    int all_ok = agree(ok);
    if (all_ok == 1) {
        if (ncclAllGather(d_handles + (size_t)ctx->tp_rank * sizeof(mine), d_handles, sizeof(mine), ncclUint8, comm, ctx->stream) != ncclSuccess) ok = 0;
        cudaMemcpyAsync(all.data(), d_handles, (size_t)W * sizeof(mine), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) ok = 0;
        for (int r = 0; r < W && ok; ++r) {
            if (r == ctx->tp_rank) { ctx->p2p_peer[r] = arena; continue; }
            void *q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
            ctx->p2p_peer[r] = q;
        }
        all_ok = agree(ok);          // also the barrier after which every arena is zeroed and mapped everywhere
    }
    cudaFree(d_ok);
    if (d_handles) cudaFree(d_handles);
    if (all_ok != 1) {
        for (int r = 0; r < W; ++r) {
            if (r != ctx->tp_rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
            ctx->p2p_peer[r] = nullptr;
        }
        if (arena) cudaFree(arena);
        if (err) cudaFree(err);
        cudaGetLastError();
        DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "peer-to-peer arena could not be set up on every rank (CUDA IPC / peer access unavailable): the NCCL path stays in use");
    }
    ctx->p2p_arena = arena;
    ctx->p2p_bytes = arena_bytes;
    ctx->p2p_err = err;
    ctx->p2p_epoch = 0;
    ctx->gate_signals = 0; ctx->gate_armed = false; ctx->gate_counters = nullptr;
    return DLLM_OK;
}

int32_t dllm_tp_p2p_status(dllm_ctx *ctx, void **arena_dev, size_t *arena_bytes, uint64_t *allreduces, uint32_t *timed_out) {
    if (!ctx) return DLLM_ERR_NULL;
    if (arena_dev) *arena_dev = ctx->p2p_arena;
    if (arena_bytes) *arena_bytes = ctx->p2p_arena ? ctx->p2p_bytes : 0;
    if (allreduces) *allreduces = ctx->p2p_calls;
    if (timed_out) {
        *timed_out = 0;
        if (ctx->p2p_err) {
            cudaSetDevice(ctx->device);
            if (cudaMemcpy(timed_out, ctx->p2p_err, sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return DLLM_ERR_CUDA;
        }
    }
    return DLLM_OK;
}

int32_t dllm_tp_finalize(dllm_ctx *ctx) {
    if (!ctx) return DLLM_ERR_NULL;
    p2p_release(ctx);
    if (ctx->nccl_comm) {
        ncclCommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->tp_rank = 0;
    ctx->tp_world = 1;
    return DLLM_OK;
}

int32_t dllm_tp_configure(dllm_ctx *ctx, int32_t chunks, int32_t reserve_sms, int32_t skip_comm) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    if (chunks < 0 || chunks > 16 || reserve_sms < -1 || reserve_sms > ctx->sm_count / 2)
        DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "chunks in 0..16 (0 = default), reserve_sms in -1..%d (-1 = default)", ctx->sm_count / 2);
    ctx->tp_chunks = chunks;
    ctx->sm_reserve = reserve_sms;
    ctx->tp_skip_comm = skip_comm != 0;
    return DLLM_OK;
}

int32_t dllm_tp_allreduce_dev(dllm_ctx *ctx, float *buf_dev, size_t n) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    return tp_allreduce(ctx, buf_dev, n);
}

int32_t dllm_tp_allgather_cols_dev(dllm_ctx *ctx, const float *in_dev, size_t M, size_t n_local, float *out_dev) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    return tp_allgather_cols(ctx, in_dev, M, n_local, out_dev);
}

}  // extern "C"
