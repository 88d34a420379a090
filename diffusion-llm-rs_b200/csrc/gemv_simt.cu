// gemv_simt.cu — f32 CUDA-core dequant-GEMV / skinny GEMM (K4, "SIMT" path).
//
// y[M,N] = x[M,K] · W + b with W = dequantize_tensor(codes)   (reference composition:
// diffuse-llm-rs/src/quantization.rs:81-85 ∘ diffuse-llm-rs/src/lib.rs:812).
// The dequantized weight is formed exactly as the reference does, (q - zp) * scale in f32,
// so the only difference to the oracle is the f32 summation order.  Memory-bound design:
// lane <-> output column, every lane streams its column's 16-byte chunks with coalesced
// 128-bit loads (wlayout.cuh), activations are staged once per CTA in shared memory and read
// as warp-wide broadcasts, K is split across CTAs so that >= 2 CTAs/SM are resident, and the
// per-split partial sums are reduced in a fixed order (deterministic; no atomics).
#include "common.cuh"
#include "kernels.h"
#include "wlayout.cuh"

namespace {

constexpr int kStageKB = 8;                        // k-blocks of activations staged per step (512 k)
constexpr int kStageK = kStageKB * WL_TILE_K;

template <int CB, int MT>
__global__ void __launch_bounds__(128)
gemv_simt_kernel(const uint8_t *__restrict__ packed, const float *__restrict__ scales,
                 const float *__restrict__ zps, const float *__restrict__ x, size_t M, size_t K,
                 size_t Npad, size_t k_blocks, size_t group_kb, int splits,
                 float *__restrict__ out /* [splits][Mpad][Npad] partials */) {
    constexpr int CH = CB / 2;
    constexpr int EPW = 32 / CB;
    __shared__ __align__(16) float xs[MT][kStageK];

    const size_t nt = blockIdx.x;
    const int split = blockIdx.y;
    const size_t m0 = (size_t)blockIdx.z * MT;     // grid.z walks M in blocks of MT rows
    const int nl = threadIdx.x;
    const size_t kb_begin = k_blocks * (size_t)split / (size_t)splits;
    const size_t kb_end = k_blocks * (size_t)(split + 1) / (size_t)splits;
    const uint4 *tiles = reinterpret_cast<const uint4 *>(packed + nt * k_blocks * wl_tile_bytes(CB));
    const size_t tile_u4 = wl_tile_bytes(CB) / 16;

    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;

    for (size_t kb_s = kb_begin; kb_s < kb_end; kb_s += kStageKB) {
        const size_t kb_e = kb_s + kStageKB < kb_end ? kb_s + kStageKB : kb_end;
        __syncthreads();
        // stage x[m0..m0+MT, kb_s*64 .. kb_e*64) (zero padded) — coalesced along k
        const size_t kbase = kb_s * WL_TILE_K;
        const int klen = (int)((kb_e - kb_s) * WL_TILE_K);
        for (int idx = nl; idx < MT * klen; idx += 128) {
            const int m = idx / klen, kk = idx - m * klen;
            const size_t k = kbase + kk;
            float v = 0.f;
            if (m0 + m < M && k < K) v = __ldg(x + (m0 + m) * K + k);
            xs[m][kk] = v;
        }
        __syncthreads();
        for (size_t kb = kb_s; kb < kb_e; ++kb) {
            const size_t g = kb / group_kb;
            const float s = __ldg(scales + g * Npad + nt * 128 + nl);
            const float z = __ldg(zps + g * Npad + nt * 128 + nl);
            uint4 c[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) c[j] = ldg_stream_u4(tiles + kb * tile_u4 + j * 128 + nl);
            const float *xk = &xs[0][(kb - kb_s) * WL_TILE_K];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const uint32_t words[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
                for (int wd = 0; wd < 4; ++wd) {
                    float q[EPW];
                    wl_decode_word_f32<CB>(words[wd], q);
                    const int koff = j * (4 * EPW) + wd * EPW;
                    float wv[EPW];
#pragma unroll
                    for (int i = 0; i < EPW; ++i)   // reference dequantize: (q - zero_point) * scale, quantization.rs:83
                        wv[i] = __fmul_rn(__fsub_rn(q[i], z), s);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const float4 *xv = reinterpret_cast<const float4 *>(xk + m * kStageK + koff);
#pragma unroll
                        for (int i4 = 0; i4 < EPW / 4; ++i4) {   // warp-wide broadcast LDS.128
                            const float4 xx = xv[i4];
                            acc[m] = fmaf(xx.x, wv[4 * i4 + 0], acc[m]);
                            acc[m] = fmaf(xx.y, wv[4 * i4 + 1], acc[m]);
                            acc[m] = fmaf(xx.z, wv[4 * i4 + 2], acc[m]);
                            acc[m] = fmaf(xx.w, wv[4 * i4 + 3], acc[m]);
                        }
                    }
                }
            }
        }
    }
    const size_t Mpad = gridDim.z * MT;
#pragma unroll
    for (int m = 0; m < MT; ++m)
        out[((size_t)split * Mpad + m0 + m) * Npad + nt * 128 + nl] = acc[m];
}

// y[m][n] = sum_split partial[split][m][n] + bias[n]   (fixed order => deterministic)
__global__ void splitk_reduce_kernel(const float *__restrict__ part, int splits, size_t M, size_t Mpad,
                                     size_t N, size_t Npad, const float *__restrict__ bias,
                                     float *__restrict__ y) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const size_t m = idx / N, n = idx - m * N;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += part[((size_t)s * Mpad + m) * Npad + n];
    if (bias) v += __ldg(bias + n);
    y[idx] = v;
}

template <int CB, int MT>
void launch_gemv(dllm_ctx *ctx, const dllm_qweight *qw, const float *x, size_t M, int splits, size_t mblocks,
                 float *part) {
    dim3 grid((unsigned)qw->n_tiles, (unsigned)splits, (unsigned)mblocks);
    gemv_simt_kernel<CB, MT><<<grid, 128, 0, ctx->stream>>>(qw->d_packed, qw->d_scales, qw->d_zps, x, M, qw->K,
                                                            qw->n_tiles * 128, qw->k_blocks,
                                                            qw->group / WL_TILE_K, splits, part);
}

template <int CB>
void launch_gemv_mt(dllm_ctx *ctx, const dllm_qweight *qw, const float *x, size_t M, int MT, int splits,
                    size_t mblocks, float *part) {
    switch (MT) {
        case 1: launch_gemv<CB, 1>(ctx, qw, x, M, splits, mblocks, part); break;
        case 2: launch_gemv<CB, 2>(ctx, qw, x, M, splits, mblocks, part); break;
        case 4: launch_gemv<CB, 4>(ctx, qw, x, M, splits, mblocks, part); break;
        default: launch_gemv<CB, 8>(ctx, qw, x, M, splits, mblocks, part); break;
    }
}

}  // namespace

int32_t k_qlinear_simt(dllm_ctx *ctx, const dllm_qweight *qw, const float *x_dev, size_t M, float *y_dev) {
    if (M == 0) return DLLM_OK;
    const int cb = wl_container_bits(qw->bits);
    const int MT = M >= 8 ? 8 : (M >= 4 ? 4 : (M >= 2 ? 2 : 1));
    const size_t mblocks = (M + MT - 1) / MT;
    if (mblocks > 65535) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "SIMT path: M=%zu too large", M);
    const size_t Mpad = mblocks * MT, Npad = qw->n_tiles * 128;
    // split K so that about 2 CTAs per SM are resident, at least one activation stage per split
    size_t ctas = qw->n_tiles * mblocks;
    size_t want = ((size_t)ctx->sm_count * 2 + ctas - 1) / ctas;
    size_t max_splits = (qw->k_blocks + kStageKB - 1) / kStageKB;
    int splits = (int)(want < 1 ? 1 : (want > max_splits ? max_splits : want));
    DLLM_TRY(ensure_buf(ctx, ctx->lin_ws, (size_t)splits * Mpad * Npad * sizeof(float)));
    float *part = (float *)ctx->lin_ws.p;
    switch (cb) {
        case 2: launch_gemv_mt<2>(ctx, qw, x_dev, M, MT, splits, mblocks, part); break;
        case 4: launch_gemv_mt<4>(ctx, qw, x_dev, M, MT, splits, mblocks, part); break;
        default: launch_gemv_mt<8>(ctx, qw, x_dev, M, MT, splits, mblocks, part); break;
    }
    LAUNCH_CHECK(ctx);
    const size_t total = M * qw->N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        part, splits, M, Mpad, qw->N, Npad, qw->d_bias, y_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}
