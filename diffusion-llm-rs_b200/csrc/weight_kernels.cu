// weight_kernels.cu — build-time kernels for a quantized weight: per-group quantizer-B
// parameters, quantize + repack into the tile-major layout (wlayout.cuh), export back.
// Quantizer B arithmetic: diffuse-llm-rs/src/quantization.rs:38-68 applied per group of
// `group` consecutive k of each output column (BASELINE.json configs[0]).
#include "common.cuh"
#include "kernels.h"
#include "wlayout.cuh"

namespace {

// thread per (group g, column n): min/max over the group's rows, then B's scale / zero-point
__global__ void __launch_bounds__(128)
wparams_kernel(const float *__restrict__ w, size_t K, size_t N, size_t Npad, size_t group, int bits,
               float *__restrict__ scales, float *__restrict__ zps) {
    const size_t n = (size_t)blockIdx.x * 128 + threadIdx.x;
    const size_t g = blockIdx.y;
    if (n >= Npad) return;
    if (n >= N) { scales[g * Npad + n] = 0.f; zps[g * Npad + n] = 0.f; return; }
    const size_t k0 = g * group, k1 = (k0 + group < K) ? k0 + group : K;
    float mx = -INFINITY, mn = INFINITY;
    for (size_t k = k0; k < k1; ++k) {           // coalesced across the warp (n fastest)
        float v = __ldg(w + k * N + n);
        mx = fmaxf(mx, v);
        mn = fminf(mn, v);
    }
    float s, z;
    params_b(mn, mx, bits, &s, &z);
    scales[g * Npad + n] = s;
    zps[g * Npad + n] = z;
}

__global__ void wparams_broadcast_kernel(const float *__restrict__ params, size_t N, size_t Npad,
                                         float *__restrict__ scales, float *__restrict__ zps) {
    const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Npad) return;
    scales[n] = n < N ? params[0] : 0.f;
    zps[n] = n < N ? params[1] : 0.f;
}

// thread per (k-block kb, column n): 64 codes -> CH chunks of 16 bytes in tile (nt, kb)
template <int CB, typename SRC>  // SRC = float (quantize with B) or uint8_t (adopt codes)
__global__ void __launch_bounds__(128)
wpack_kernel(const SRC *__restrict__ src, size_t K, size_t N, size_t Npad, size_t group, int bits,
             size_t k_blocks, const float *__restrict__ scales, const float *__restrict__ zps,
             uint8_t *__restrict__ packed, unsigned int *__restrict__ flags) {
    constexpr int CH = CB / 2;          // chunks per column per tile
    constexpr int EPW = 32 / CB;        // codes per word
    const size_t nt = blockIdx.x, kb = blockIdx.y;
    const int nl = threadIdx.x;
    const size_t n = nt * 128 + nl;
    const size_t g = (kb * WL_TILE_K) / group;
    const bool col_ok = n < N;
    const float s = scales[g * Npad + n], z = zps[g * Npad + n];
    const int hi = (1 << bits) - 1;
    const uint32_t pad_code = col_ok ? (uint32_t)z : 0u;   // K padding dequantizes to (zp - zp) * s = 0
    uint4 *tile = reinterpret_cast<uint4 *>(packed + (nt * k_blocks + kb) * wl_tile_bytes(CB));
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        uint32_t words[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) {
#pragma unroll
            for (int i = 0; i < EPW; ++i) {
                const size_t k = kb * WL_TILE_K + (size_t)j * (4 * EPW) + wd * EPW + i;
                uint32_t c = pad_code;
                if (col_ok && k < K) {
                    if (sizeof(SRC) == 4) c = code_b((float)src[k * N + n], s, z, hi);
                    else {
                        c = (uint32_t)src[k * N + n];
                        if (c > (uint32_t)hi) atomicOr(flags, 4u);   // a code that does not fit `bits`: rejected by the host, never truncated
                    }
                }
                words[wd] |= (c & ((1u << CB) - 1u)) << wl_bitpos<CB>(i);
            }
        }
        tile[j * 128 + nl] = make_uint4(words[0], words[1], words[2], words[3]);
    }
}

// per (group, column): the two 32-bit operands the tcgen05 dequant needs, so that its inner loop is
// one LDS.64 instead of two dependent global loads and two conversions.
//   2/4-bit: x = bf16x2(128 + zp)   (subtracted from the magic-number form 128 + q, exact)
//   8-bit  : x = f32 bits of zp
//   y = bf16x2(scale)
// and for the GEMV path (int8 tensor path, gemv_mma.cu): {f32 scale, f32 zp}
// flags (OR-ed): 1 = some zero-point is not an integer in [0, 255] (the tensor-core kernels subtract it exactly
// only if it is: quantizer B always produces such, hand-made parameters may not — those weights run on the f32 SIMT path)
__global__ void wdq_params_kernel(const float *__restrict__ scales, const float *__restrict__ zps, size_t n, int cb,
                                  uint2 *__restrict__ out, uint2 *__restrict__ gout, unsigned int *__restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = scales[i], z = zps[i];
    unsigned int f = 0;
    if (!(z >= 0.f && z <= 255.f && z == rintf(z))) f |= 1u;
    if (f) atomicOr(flags, f);
    __nv_bfloat162 sb = __float2bfloat162_rn(s);
    uint2 o;
    o.y = *reinterpret_cast<uint32_t *>(&sb);
    if (cb == 8) {
        o.x = __float_as_uint(z);
    } else {
        __nv_bfloat162 zb = __float2bfloat162_rn(128.0f + z);
        o.x = *reinterpret_cast<uint32_t *>(&zb);
    }
    out[i] = o;
    gout[i] = make_uint2(__float_as_uint(s), __float_as_uint(z));
}

template <int CB>
__global__ void __launch_bounds__(128)
wexport_kernel(const uint8_t *__restrict__ packed, size_t K, size_t N, size_t k_blocks,
               uint8_t *__restrict__ codes) {
    constexpr int CH = CB / 2;
    constexpr int EPW = 32 / CB;
    const size_t nt = blockIdx.x, kb = blockIdx.y;
    const int nl = threadIdx.x;
    const size_t n = nt * 128 + nl;
    if (n >= N) return;
    const uint4 *tile = reinterpret_cast<const uint4 *>(packed + (nt * k_blocks + kb) * wl_tile_bytes(CB));
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const uint4 v = tile[j * 128 + nl];
        const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int wd = 0; wd < 4; ++wd)
#pragma unroll
            for (int i = 0; i < EPW; ++i) {
                const size_t k = kb * WL_TILE_K + (size_t)j * (4 * EPW) + wd * EPW + i;
                if (k < K) codes[k * N + n] = (uint8_t)((words[wd] >> wl_bitpos<CB>(i)) & ((1u << CB) - 1u));
            }
    }
}

}  // namespace

int32_t k_wparams_grouped(dllm_ctx *ctx, const float *w_dev, size_t K, size_t N, size_t group, int bits,
                          float *scales_dev, float *zps_dev) {
    const size_t Npad = (N + 127) / 128 * 128;
    const size_t G = (K + group - 1) / group;
    dim3 grid((unsigned)(Npad / 128), (unsigned)G);
    wparams_kernel<<<grid, 128, 0, ctx->stream>>>(w_dev, K, N, Npad, group, bits, scales_dev, zps_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_wparams_broadcast(dllm_ctx *ctx, const float *params_dev, size_t N, float *scales_dev, float *zps_dev) {
    const size_t Npad = (N + 127) / 128 * 128;
    wparams_broadcast_kernel<<<(unsigned)((Npad + 255) / 256), 256, 0, ctx->stream>>>(params_dev, N, Npad, scales_dev, zps_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

template <typename SRC>
static int32_t wpack_any(dllm_ctx *ctx, const SRC *src, dllm_qweight *qw) {
    const size_t Npad = qw->n_tiles * 128;
    const int cb = wl_container_bits(qw->bits);
    dim3 grid((unsigned)qw->n_tiles, (unsigned)qw->k_blocks);
    unsigned int *flags = reinterpret_cast<unsigned int *>(ctx->d_params + 9);
    if (sizeof(SRC) == 1) CUDA_TRY(ctx, cudaMemsetAsync(flags, 0, sizeof(unsigned int), ctx->stream));
    switch (cb) {
        case 2: wpack_kernel<2, SRC><<<grid, 128, 0, ctx->stream>>>(src, qw->K, qw->N, Npad, qw->group, qw->bits, qw->k_blocks, qw->d_scales, qw->d_zps, qw->d_packed, flags); break;
        case 4: wpack_kernel<4, SRC><<<grid, 128, 0, ctx->stream>>>(src, qw->K, qw->N, Npad, qw->group, qw->bits, qw->k_blocks, qw->d_scales, qw->d_zps, qw->d_packed, flags); break;
        default: wpack_kernel<8, SRC><<<grid, 128, 0, ctx->stream>>>(src, qw->K, qw->N, Npad, qw->group, qw->bits, qw->k_blocks, qw->d_scales, qw->d_zps, qw->d_packed, flags); break;
    }
    LAUNCH_CHECK(ctx);
    if (sizeof(SRC) == 1) {   // adopted codes: one that does not fit `bits` is a format error (InvalidDataFormat, error.rs:38)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_params + 9, flags, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (*reinterpret_cast<unsigned int *>(ctx->h_params + 9) & 4u)
            DLLM_FAIL(ctx, DLLM_ERR_INVALID_DATA_FORMAT, "a code exceeds 2^%d - 1", qw->bits);
    }
    return DLLM_OK;
}

int32_t k_wpack_from_f32(dllm_ctx *ctx, const float *w_dev, dllm_qweight *qw) { return wpack_any<float>(ctx, w_dev, qw); }
int32_t k_wpack_from_codes(dllm_ctx *ctx, const uint8_t *codes_dev, dllm_qweight *qw) { return wpack_any<uint8_t>(ctx, codes_dev, qw); }

int32_t k_wdq_params(dllm_ctx *ctx, dllm_qweight *qw) {
    const size_t G = qw->per_tensor ? 1 : qw->K / qw->group;
    const size_t n = G * qw->n_tiles * 128;
    unsigned int *flags = reinterpret_cast<unsigned int *>(ctx->d_params + 8);
    CUDA_TRY(ctx, cudaMemsetAsync(flags, 0, sizeof(unsigned int), ctx->stream));
    wdq_params_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(qw->d_scales, qw->d_zps, n,
                                                                            wl_container_bits(qw->bits), qw->d_dqparams, qw->d_gparams, flags);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_params + 8, flags, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned int f = *reinterpret_cast<unsigned int *>(ctx->h_params + 8);
    qw->int_zps = (f & 1u) == 0;
    return DLLM_OK;
}

int32_t k_wexport_codes(dllm_ctx *ctx, const dllm_qweight *qw, uint8_t *codes_dev) {
    const int cb = wl_container_bits(qw->bits);
    dim3 grid((unsigned)qw->n_tiles, (unsigned)qw->k_blocks);
    switch (cb) {
        case 2: wexport_kernel<2><<<grid, 128, 0, ctx->stream>>>(qw->d_packed, qw->K, qw->N, qw->k_blocks, codes_dev); break;
        case 4: wexport_kernel<4><<<grid, 128, 0, ctx->stream>>>(qw->d_packed, qw->K, qw->N, qw->k_blocks, codes_dev); break;
        default: wexport_kernel<8><<<grid, 128, 0, ctx->stream>>>(qw->d_packed, qw->K, qw->N, qw->k_blocks, codes_dev); break;
    }
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}
