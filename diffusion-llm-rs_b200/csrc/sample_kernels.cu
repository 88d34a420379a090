// sample_kernels.cu — elementwise kernels of the denoising step: activation casts for the
// tcgen05 path and the fused p_sample update (diffuse-llm-rs/src/lib.rs:1152-1215).
#include "common.cuh"
#include "kernels.h"
#include "noise.cuh"

namespace {

// out[i] = element i0 + i of stream `stream` ("dllm_noise v1", noise.cuh); two elements (one Box-Muller pair) per thread
__global__ void __launch_bounds__(256)
noise_fill_kernel(unsigned long long seed, unsigned long long stream, unsigned long long i0, size_t n, float *__restrict__ out) {
    const uint64_t key = dn_key(seed, stream);
    const uint64_t first_pair = i0 >> 1, last = i0 + n;                      // elements [i0, last)
    const uint64_t n_pairs = ((last + 1) >> 1) - first_pair;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (uint64_t)gridDim.x * blockDim.x) {
        float z0, z1;
        dn_normal_pair(key, first_pair + p, &z0, &z1);
        const uint64_t e0 = 2 * (first_pair + p);
        if (e0 >= i0 && e0 < last) out[e0 - i0] = z0;
        if (e0 + 1 >= i0 && e0 + 1 < last) out[e0 + 1 - i0] = z1;
    }
}

// one denoise step's update with the noise generated in the kernel (no z tensor in HBM):
//   x_prev = (c1 * x_t + c2 * noise_pred) + std * z,   z = element i of stream t under `seed`, z = 0 when t == 0
// t and seed come from device memory when `state` is given ({int t; pad; u64 seed}: the CUDA-graph replay of the step
// reads them there, and `advance` != 0 makes the last thread decrement t for the next replay).
struct SampleState { int t; int pad; unsigned long long seed; };
__global__ void __launch_bounds__(256)
p_sample_seeded_kernel(const float *x /* may alias out */, const float *__restrict__ pred, const float *__restrict__ coef_table,
                       const SampleState *__restrict__ state, int t_arg, unsigned long long seed_arg, int T, size_t total,
                       float *out) {
    const int t = state ? state->t : t_arg;
    const unsigned long long seed = state ? state->seed : seed_arg;
    const float *coef = coef_table + 4 * (size_t)(t < T - 1 ? t : T - 1);                    // lib.rs:1174 clamp
    const float c1 = __ldg(coef), c2 = __ldg(coef + 1), sd = __ldg(coef + 2), degenerate = __ldg(coef + 3);
    const uint64_t key = dn_key(seed, (uint64_t)t);
    const size_t n_pairs = (total + 1) >> 1;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (size_t)gridDim.x * blockDim.x) {
        float z0 = 0.f, z1 = 0.f;
        if (t > 0) dn_normal_pair(key, p, &z0, &z1);                                         // lib.rs:1199-1205
        const size_t i = 2 * p;
        const bool two = i + 1 < total;
        float xv0 = x[i], xv1 = two ? x[i + 1] : 0.f;
        if (degenerate != 0.f) { out[i] = xv0; if (two) out[i + 1] = xv1; continue; }
        const float m0 = __fadd_rn(__fmul_rn(c1, xv0), __fmul_rn(c2, pred[i]));
        out[i] = __fadd_rn(m0, __fmul_rn(sd, z0));
        if (two) {
            const float m1 = __fadd_rn(__fmul_rn(c1, xv1), __fmul_rn(c2, pred[i + 1]));
            out[i + 1] = __fadd_rn(m1, __fmul_rn(sd, z1));
        }
    }
}

__global__ void sample_state_step_kernel(SampleState *state) { state->t -= 1; }

__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float *__restrict__ in, size_t n, __nv_bfloat16 *__restrict__ out) {
    const size_t n4 = n >> 2;
    const float4 *i4 = reinterpret_cast<const float4 *>(in);
    uint2 *o4 = reinterpret_cast<uint2 *>(out);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = ldg_stream_f4(i4 + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        o4[i] = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t j = n4 << 2; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
}

__global__ void __launch_bounds__(256)
bf16_to_f32_kernel(const __nv_bfloat16 *__restrict__ in, size_t n, float *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __bfloat162float(in[i]);
}

// x_prev = (c1 * x_t + c2 * noise_pred) + std * z       lib.rs:1195-1196, 1212
// two products and an add, then a product and an add: no FMA contraction (Rust does not contract)
__global__ void __launch_bounds__(256)
p_sample_kernel(const float *x /* may alias out */, const float *__restrict__ pred, const float *__restrict__ z,
                const float *__restrict__ coef_table, const int *__restrict__ rowmap, int row, size_t batch,
                size_t feat, float *out) {
    const size_t total = batch * feat;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / feat;
        const float *coef = coef_table + 4 * (size_t)(rowmap ? __ldg(rowmap + b) : row);
        const float c1 = __ldg(coef), c2 = __ldg(coef + 1), sd = __ldg(coef + 2);
        const float degenerate = __ldg(coef + 3);
        const float xv = x[i];
        if (degenerate != 0.f) { out[i] = xv; continue; }
        const float mean = __fadd_rn(__fmul_rn(c1, xv), __fmul_rn(c2, pred[i]));
        const float nz = z ? z[i] : 0.f;
        out[i] = __fadd_rn(mean, __fmul_rn(sd, nz));
    }
}

// noisy = x_start * sqrt(alpha_bar_t) + noise * sqrt(1 - alpha_bar_t)      lib.rs:1131-1133
// (two products, one add: no contraction).  tab rows are {sqrt(alpha_bar), sqrt(1 - alpha_bar)} per timestep.
__global__ void __launch_bounds__(256)
add_noise_kernel(const float *__restrict__ x, const float *__restrict__ noise, const float *__restrict__ tab,
                 const int *__restrict__ rowmap, int row, size_t batch, size_t feat, float *__restrict__ out) {
    const size_t total = batch * feat;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / feat;
        const float *c = tab + 2 * (size_t)(rowmap ? __ldg(rowmap + b) : row);
        out[i] = __fadd_rn(__fmul_rn(x[i], __ldg(c)), __fmul_rn(noise[i], __ldg(c + 1)));
    }
}

inline int grid1d(const dllm_ctx *ctx, size_t items) {
    size_t b = (items + 255) / 256, cap = (size_t)ctx->sm_count * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int32_t k_f32_to_bf16(dllm_ctx *ctx, const float *in_dev, size_t n, void *out_bf16_dev) {
    if (n == 0) return DLLM_OK;
    if (!aligned16(in_dev) || (reinterpret_cast<uintptr_t>(out_bf16_dev) & 7u))
        DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "f32->bf16 cast needs aligned device buffers");
    f32_to_bf16_kernel<<<grid1d(ctx, (n + 3) / 4), 256, 0, ctx->stream>>>(in_dev, n, (__nv_bfloat16 *)out_bf16_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

namespace {
// one CTA per token row: absmax, then codes / row sum.  V > 0: the row (<= 2048 V elements) stays in registers between the two
// passes (one HBM read); V == 0: any K, the second pass re-reads the row (from L1 / L2)
template <int V>
__global__ void __launch_bounds__(256) rowquant_i8_kernel(const __nv_bfloat16 *__restrict__ x, uint32_t K, float wscale,
                                                          int8_t *__restrict__ xq, float *__restrict__ rowscale, int32_t *__restrict__ rowsum) {
    __shared__ float s_max[8];
    __shared__ int s_sum[8];
    const size_t row = blockIdx.x;
    const uint4 *src = reinterpret_cast<const uint4 *>(x + row * K);
    const uint32_t n8 = K / 8;
    constexpr int R = V > 0 ? V : 1;
    uint4 reg[R];
    float m = 0.f;
    // programmatic dependent launch on both sides: the linear that consumes the codes may set up while this kernel runs, and this
    // kernel's CTAs are placed while the linear that produces x drains (they wait here until its writes are visible)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    auto absmax8 = [&](const uint4 &v) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            m = fmaxf(m, fabsf(__uint_as_float(w[e] << 16)));
            m = fmaxf(m, fabsf(__uint_as_float(w[e] & 0xffff0000u)));
        }
    };
    if constexpr (V > 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const uint32_t idx = threadIdx.x + 256u * i;
            reg[i] = idx < n8 ? __ldg(src + idx) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) absmax8(reg[i]);
    } else {
        for (uint32_t i = threadIdx.x; i < n8; i += 256) absmax8(__ldg(src + i));
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sh));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_max[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s_max[i]);
    m = fminf(m, 3.0e38f);                                   // (an infinite activation saturates instead of zeroing the row)
    const float inv = m > 0.f ? 127.f / m : 0.f;
    int acc = 0;
    uint2 *dst = reinterpret_cast<uint2 *>(xq + row * K);
    auto encode8 = [&](const uint4 &v, uint32_t idx) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        int q[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            q[2 * e] = max(-127, min(127, __float2int_rn(__uint_as_float(w[e] << 16) * inv)));
            q[2 * e + 1] = max(-127, min(127, __float2int_rn(__uint_as_float(w[e] & 0xffff0000u) * inv)));
        }
        uint2 o;
        o.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
        o.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
        dst[idx] = o;
        acc = __dp4a((int)o.x, 0x01010101, acc);
        acc = __dp4a((int)o.y, 0x01010101, acc);
    };
    if constexpr (V > 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const uint32_t idx = threadIdx.x + 256u * i;
            if (idx < n8) encode8(reg[i], idx);
        }
    } else {
        for (uint32_t i = threadIdx.x; i < n8; i += 256) encode8(__ldg(src + i), i);
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_sum[i];
        rowsum[row] = t;
        rowscale[row] = wscale * (m > 0.f ? m / 127.f : 1.f);
    }
}

// K <= 2048: one WARP per token row — the row's <= 256 16-byte words stay in registers (8 per lane), the reductions are warp
// shuffles, no block barrier: 8 loads in flight per thread instead of 1 (the CTA-per-row kernel took 15 us for [8192, 2048]: it is
// a latency chain per CTA, 7 waves of them)
__global__ void __launch_bounds__(256) rowquant_i8_warp_kernel(const __nv_bfloat16 *__restrict__ x, uint32_t M, uint32_t K, float wscale,
                                                               int8_t *__restrict__ xq, float *__restrict__ rowscale,
                                                               int32_t *__restrict__ rowsum) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t lane = threadIdx.x & 31;
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(x + row * K);
    const uint32_t n8 = K / 8;
    uint4 reg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t idx = lane + 32u * i;
        reg[i] = idx < n8 ? __ldg(src + idx) : make_uint4(0, 0, 0, 0);
    }
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t w[4] = {reg[i].x, reg[i].y, reg[i].z, reg[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            m = fmaxf(m, fabsf(__uint_as_float(w[e] << 16)));
            m = fmaxf(m, fabsf(__uint_as_float(w[e] & 0xffff0000u)));
        }
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sh));
    m = fminf(m, 3.0e38f);
    const float inv = m > 0.f ? 127.f / m : 0.f;
    int acc = 0;
    uint2 *dst = reinterpret_cast<uint2 *>(xq + row * K);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t idx = lane + 32u * i;
        if (idx < n8) {
            const uint32_t w[4] = {reg[i].x, reg[i].y, reg[i].z, reg[i].w};
            int q[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                q[2 * e] = max(-127, min(127, __float2int_rn(__uint_as_float(w[e] << 16) * inv)));
                q[2 * e + 1] = max(-127, min(127, __float2int_rn(__uint_as_float(w[e] & 0xffff0000u) * inv)));
            }
            uint2 o;
            o.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
            o.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
            dst[idx] = o;
            acc = __dp4a((int)o.x, 0x01010101, acc);
            acc = __dp4a((int)o.y, 0x01010101, acc);
        }
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
    if (lane == 0) {
        rowsum[row] = acc;
        rowscale[row] = wscale * (m > 0.f ? m / 127.f : 1.f);
    }
}
}  // namespace

int32_t k_rowquant_i8(dllm_ctx *ctx, const void *x_bf16_dev, size_t M, size_t K, float wscale, int8_t *xq_dev, float *rowscale_dev,
                      int32_t *rowsum_dev) {
    if (M == 0) return DLLM_OK;
    if (K % 8 != 0 || (reinterpret_cast<uintptr_t>(x_bf16_dev) & 15u) || (reinterpret_cast<uintptr_t>(xq_dev) & 7u))
        DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "int8 activation quantizer: K %% 8 == 0 and aligned buffers required");
    const __nv_bfloat16 *x = (const __nv_bfloat16 *)x_bf16_dev;
    const uint32_t k = (uint32_t)K;
    const unsigned g = (unsigned)M;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g);
    cfg.blockDim = dim3(256);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool no_pdl = getenv("DLLM_UMMA_NO_PDL") != nullptr;      // experiments only
    cfg.numAttrs = no_pdl ? 0 : 1;
    if (K <= 2048) {
        cfg.gridDim = dim3((unsigned)((M + 7) / 8));
        CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, rowquant_i8_warp_kernel, x, (uint32_t)M, k, wscale, xq_dev, rowscale_dev, rowsum_dev));
        LAUNCH_CHECK(ctx);
        return DLLM_OK;
    }
    auto kern = K <= 2048 ? rowquant_i8_kernel<1> : K <= 4096 ? rowquant_i8_kernel<2> : K <= 8192 ? rowquant_i8_kernel<4>
              : K <= 16384 ? rowquant_i8_kernel<8> : rowquant_i8_kernel<0>;
    CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, x, k, wscale, xq_dev, rowscale_dev, rowsum_dev));
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_bf16_to_f32(dllm_ctx *ctx, const void *in_bf16_dev, size_t n, float *out_dev) {
    if (n == 0) return DLLM_OK;
    bf16_to_f32_kernel<<<grid1d(ctx, n), 256, 0, ctx->stream>>>((const __nv_bfloat16 *)in_bf16_dev, n, out_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_p_sample(dllm_ctx *ctx, const float *x_dev, const float *pred_dev, const float *z_dev,
                   const float *coef_table_dev, const int *rowmap_dev, int row, size_t batch, size_t feat,
                   float *out_dev) {
    if (batch * feat == 0) return DLLM_OK;
    p_sample_kernel<<<grid1d(ctx, batch * feat), 256, 0, ctx->stream>>>(x_dev, pred_dev, z_dev, coef_table_dev,
                                                                        rowmap_dev, row, batch, feat, out_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_add_noise(dllm_ctx *ctx, const float *x_dev, const float *noise_dev, const float *tab_dev, const int *rowmap_dev,
                    int row, size_t batch, size_t feat, float *out_dev) {
    if (batch * feat == 0) return DLLM_OK;
    add_noise_kernel<<<grid1d(ctx, batch * feat), 256, 0, ctx->stream>>>(x_dev, noise_dev, tab_dev, rowmap_dev, row, batch,
                                                                         feat, out_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_noise_fill(dllm_ctx *ctx, unsigned long long seed, unsigned long long stream, unsigned long long i0, size_t n,
                     float *out_dev) {
    if (n == 0) return DLLM_OK;
    noise_fill_kernel<<<grid1d(ctx, (n + 1) / 2 + 1), 256, 0, ctx->stream>>>(seed, stream, i0, n, out_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_p_sample_seeded(dllm_ctx *ctx, const float *x_dev, const float *pred_dev, const float *coef_table_dev,
                          const void *state_dev, int t, unsigned long long seed, int T, size_t total, float *out_dev) {
    if (total == 0) return DLLM_OK;
    p_sample_seeded_kernel<<<grid1d(ctx, (total + 1) / 2), 256, 0, ctx->stream>>>(x_dev, pred_dev, coef_table_dev,
                                                                                 (const SampleState *)state_dev, t, seed, T, total, out_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_sample_state_step(dllm_ctx *ctx, void *state_dev) {
    sample_state_step_kernel<<<1, 1, 0, ctx->stream>>>((SampleState *)state_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}
