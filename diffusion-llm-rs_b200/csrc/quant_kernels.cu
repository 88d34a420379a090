// quant_kernels.cu — HBM-bound quantize / dequantize / pack / unpack kernels (K1, K2, K3, K6).
//
// Reference arithmetic restated on the device (file:line under the reference root):
//   B  diffuse-llm-rs/src/quantization.rs:38-68, 81-85
//   A  quantization/src/quantize.rs:111-154, 172-184
//   C  prefill-kvquant-rs/lib.rs:39-53
//   D  diffusion_prefill/src/prefill_kv.rs:53-67, 104-121
// All of them are streaming kernels: one coalesced 128-bit load per lane per step, codes
// produced in registers (IEEE division via __fdiv_rn, no FMA contraction), narrow coalesced
// stores.  Grids are multiples of the SM count.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------
// min / max fold (f32::max / f32::min ignore NaN == fmaxf / fminf) + optional B parameters
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
minmax_kernel(const float *__restrict__ x, size_t n, int vec_ok, float *partials,
              unsigned int *ticket, float *out, int bits /* 0: min/max only */) {
    float mx = -INFINITY, mn = INFINITY;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    if (vec_ok) {
        const size_t n4 = n >> 2;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        size_t i = tid;
        for (; i + 3 * nth < n4; i += 4 * nth) {
            float4 a = ldg_stream_f4(x4 + i), b = ldg_stream_f4(x4 + i + nth);
            float4 c = ldg_stream_f4(x4 + i + 2 * nth), d = ldg_stream_f4(x4 + i + 3 * nth);
            mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
            mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w)), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w))));
            mn = fminf(mn, fminf(fminf(fminf(a.x, a.y), fminf(a.z, a.w)), fminf(fminf(b.x, b.y), fminf(b.z, b.w))));
            mn = fminf(mn, fminf(fminf(fminf(c.x, c.y), fminf(c.z, c.w)), fminf(fminf(d.x, d.y), fminf(d.z, d.w))));
        }
        for (; i < n4; i += nth) {
            float4 a = ldg_stream_f4(x4 + i);
            mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
            mn = fminf(mn, fminf(fminf(a.x, a.y), fminf(a.z, a.w)));
        }
        for (size_t j = (n4 << 2) + tid; j < n; j += nth) {
            mx = fmaxf(mx, x[j]);
            mn = fminf(mn, x[j]);
        }
    } else {
        for (size_t j = tid; j < n; j += nth) {
            mx = fmaxf(mx, x[j]);
            mn = fminf(mn, x[j]);
        }
    }
    __shared__ float smx[kThreads / 32], smn[kThreads / 32];
    __shared__ bool is_last;
    mx = warp_max(mx);
    mn = warp_min(mn);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { smx[w] = mx; smn[w] = mn; }
    __syncthreads();
    if (w == 0) {
        mx = lane < kThreads / 32 ? smx[lane] : -INFINITY;
        mn = lane < kThreads / 32 ? smn[lane] : INFINITY;
        mx = warp_max(mx);
        mn = warp_min(mn);
        if (lane == 0) {
            partials[blockIdx.x] = mx;
            partials[kMaxPartials + blockIdx.x] = mn;
            __threadfence();
            unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    mx = -INFINITY; mn = INFINITY;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        mx = fmaxf(mx, __ldcg(partials + i));
        mn = fminf(mn, __ldcg(partials + kMaxPartials + i));
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    __syncthreads();
    if (lane == 0) { smx[w] = mx; smn[w] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kThreads / 32; ++i) { mx = fmaxf(mx, smx[i]); mn = fminf(mn, smn[i]); }
        float scale = 0.f, zp = 0.f;
        if (bits > 0) params_b(mn, mx, bits, &scale, &zp);
        out[0] = scale; out[1] = zp; out[2] = mn; out[3] = mx;
        *ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------
// elementwise encode: f32 -> codes (one per u8, or bit-packed LSB-first)
// ------------------------------------------------------------------------------------------
struct EncArgs {
    float scale, zp;      // used when dev == nullptr
    float lo, hi;         // A: clamp range; C/D: hi = levels
    int ihi;              // B: (1<<bits)-1
    const float *dev;     // B: {scale, zp} resident on the device
};

template <int SCHEME>  // 0 = B, 1 = A, 2 = C/D
__device__ __forceinline__ uint8_t encode_one(float x, float s, float z, const EncArgs &a) {
    if (SCHEME == 0) return code_b(x, s, z, a.ihi);
    if (SCHEME == 1) return code_a(x, s, z, a.lo, a.hi);
    return code_cd(x, s, z, a.hi);
}

// store the 4 codes of one lane's float4.  PACK: 0 -> 4 bytes; 8 -> 4 bytes; 4 -> 2 bytes;
// 2 -> 1 byte; 1 -> half a byte (even lane merges its neighbour's nibble).
template <int PACK>
__device__ __forceinline__ void store_codes4(uint8_t *out, size_t i4, uint32_t c0, uint32_t c1,
                                             uint32_t c2, uint32_t c3, bool active) {
    if (PACK == 0 || PACK == 8) {
        if (active) reinterpret_cast<uint32_t *>(out)[i4] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
    } else if (PACK == 4) {
        if (active) reinterpret_cast<uint16_t *>(out)[i4] = (uint16_t)(c0 | (c1 << 4) | (c2 << 8) | (c3 << 12));
    } else if (PACK == 2) {
        if (active) out[i4] = (uint8_t)(c0 | (c1 << 2) | (c2 << 4) | (c3 << 6));
    } else {  // 1 bit: lane pair -> one byte
        uint32_t nib = active ? (c0 | (c1 << 1) | (c2 << 2) | (c3 << 3)) : 0u;
        uint32_t other = __shfl_xor_sync(0xffffffffu, nib, 1);
        if (active && !(threadIdx.x & 1)) out[i4 >> 1] = (uint8_t)(nib | (other << 4));
    }
}

template <int SCHEME, int PACK>
__global__ void __launch_bounds__(kThreads)
encode_kernel(const float *__restrict__ x, size_t n, uint8_t *__restrict__ out, EncArgs a) {
    float s = a.scale, z = a.zp;
    // quantizer B with its parameters on the device ({scale, zp, min, max} from minmax_kernel): one divisor for the
    // whole tensor and a known numerator range -> hoisted-reciprocal division (common.cuh), bit-identical quotients
    RowDivisor rd;
    rd.fast = false;
    if (a.dev) {
        s = __ldg(a.dev); z = __ldg(a.dev + 1);
        if (SCHEME == 0) rd = make_row_divisor(s, fmaxf(fabsf(__ldg(a.dev + 2)), fabsf(__ldg(a.dev + 3))));
    }
    const size_t n4 = n >> 2;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    // every lane of a warp runs the same number of iterations (the 1-bit path shuffles)
    const size_t iters = (n4 + nth - 1) / nth;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t it = 0; it < iters; ++it, i += nth) {
        const bool act = i < n4;
        float4 v = act ? ldg_stream_f4(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t c0, c1, c2, c3;
        if (SCHEME == 0 && rd.fast) {          // code_b with the division replaced: round(x / scale + zp), clamp
            c0 = (uint32_t)rs_clampi(rs_as_i32(roundf(__fadd_rn(div_row_checked(rd, v.x), z))), 0, a.ihi);
            c1 = (uint32_t)rs_clampi(rs_as_i32(roundf(__fadd_rn(div_row_checked(rd, v.y), z))), 0, a.ihi);
            c2 = (uint32_t)rs_clampi(rs_as_i32(roundf(__fadd_rn(div_row_checked(rd, v.z), z))), 0, a.ihi);
            c3 = (uint32_t)rs_clampi(rs_as_i32(roundf(__fadd_rn(div_row_checked(rd, v.w), z))), 0, a.ihi);
        } else {
            c0 = encode_one<SCHEME>(v.x, s, z, a); c1 = encode_one<SCHEME>(v.y, s, z, a);
            c2 = encode_one<SCHEME>(v.z, s, z, a); c3 = encode_one<SCHEME>(v.w, s, z, a);
        }
        store_codes4<PACK>(out, i, c0, c1, c2, c3, act);
    }
}

// ragged tail (< 4 elements) of encode_kernel: launched after it on the same stream (the 1-bit
// tail shares a byte with the last full float4, so it must not race with the main kernel)
template <int SCHEME, int PACK>
__global__ void encode_tail_kernel(const float *__restrict__ x, size_t n, uint8_t *__restrict__ out, EncArgs a) {
    float s = a.scale, z = a.zp;
    if (a.dev) { s = __ldg(a.dev); z = __ldg(a.dev + 1); }
    const size_t base = (n >> 2) << 2;
    uint32_t acc = 0;
    for (size_t j = base; j < n; ++j) {
        uint32_t c = encode_one<SCHEME>(x[j], s, z, a);
        if (PACK == 0 || PACK == 8) out[j] = (uint8_t)c;
        else acc |= c << ((j - base) * PACK);
    }
    if (PACK == 4) { out[base >> 1] = (uint8_t)acc; if ((n & 3) == 3) out[(base >> 1) + 1] = (uint8_t)(acc >> 8); }
    if (PACK == 2) out[base >> 2] = (uint8_t)acc;
    if (PACK == 1) {
        const size_t byte = base >> 3;
        if (base & 4) out[byte] = (uint8_t)((out[byte] & 0x0f) | (acc << 4));
        else out[byte] = (uint8_t)acc;
    }
}

// unaligned / generic scalar fallback (device pointers that are not 16-byte aligned)
template <int SCHEME>
__global__ void encode_scalar_kernel(const float *__restrict__ x, size_t n, uint8_t *out, EncArgs a) {
    float s = a.scale, z = a.zp;
    if (a.dev) { s = __ldg(a.dev); z = __ldg(a.dev + 1); }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = encode_one<SCHEME>(x[i], s, z, a);
}

// ------------------------------------------------------------------------------------------
// elementwise decode: codes -> f32
// ------------------------------------------------------------------------------------------
struct DecArgs {
    float scale, zp;
    const float *dev;         // {scale, zp} on the device (per tensor)
    const float *row_scales;  // per-row parameters (D); row = element / dim
    const float *row_zps;
    size_t dim;
};

template <int PACK>
__device__ __forceinline__ void load_codes4(const uint8_t *in, size_t i4, uint32_t &c0, uint32_t &c1,
                                            uint32_t &c2, uint32_t &c3) {
    if (PACK == 0 || PACK == 8) {
        uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(in) + i4);
        c0 = w & 255u; c1 = (w >> 8) & 255u; c2 = (w >> 16) & 255u; c3 = w >> 24;
    } else if (PACK == 4) {
        uint32_t w = __ldg(reinterpret_cast<const uint16_t *>(in) + i4);
        c0 = w & 15u; c1 = (w >> 4) & 15u; c2 = (w >> 8) & 15u; c3 = (w >> 12) & 15u;
    } else if (PACK == 2) {
        uint32_t w = __ldg(in + i4);
        c0 = w & 3u; c1 = (w >> 2) & 3u; c2 = (w >> 4) & 3u; c3 = (w >> 6) & 3u;
    } else {
        uint32_t w = __ldg(in + (i4 >> 1));
        w = (i4 & 1) ? (w >> 4) : (w & 15u);
        c0 = w & 1u; c1 = (w >> 1) & 1u; c2 = (w >> 2) & 1u; c3 = (w >> 3) & 1u;
    }
}

template <int FORM /*0: (q-zp)*s  1: q*s+zp*/, int PACK>
__global__ void __launch_bounds__(kThreads)
decode_kernel(const uint8_t *__restrict__ in, size_t n, float *__restrict__ out, DecArgs a) {
    float s = a.scale, z = a.zp;
    if (a.dev) { s = __ldg(a.dev); z = __ldg(a.dev + 1); }
    const size_t n4 = n >> 2;
    float4 *o4 = reinterpret_cast<float4 *>(out);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (a.row_scales) {  // dim % 4 == 0 guaranteed by the launcher
            const size_t r = (i << 2) / a.dim;
            s = __ldg(a.row_scales + r);
            z = __ldg(a.row_zps + r);
        }
        uint32_t c0, c1, c2, c3;
        load_codes4<PACK>(in, i, c0, c1, c2, c3);
        float4 v;
        if (FORM == 0) {
            v.x = deq_ab((uint8_t)c0, s, z); v.y = deq_ab((uint8_t)c1, s, z);
            v.z = deq_ab((uint8_t)c2, s, z); v.w = deq_ab((uint8_t)c3, s, z);
        } else {
            v.x = deq_cd((uint8_t)c0, s, z); v.y = deq_cd((uint8_t)c1, s, z);
            v.z = deq_cd((uint8_t)c2, s, z); v.w = deq_cd((uint8_t)c3, s, z);
        }
        stg_stream_f4(o4 + i, v);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 3)) {
        for (size_t j = n4 << 2; j < n; ++j) {
            if (a.row_scales) { s = a.row_scales[j / a.dim]; z = a.row_zps[j / a.dim]; }
            uint32_t q;
            if (PACK == 0 || PACK == 8) q = in[j];
            else { const int per = 8 / (PACK ? PACK : 8); q = (in[j / per] >> ((j % per) * PACK)) & ((1u << PACK) - 1u); }
            out[j] = FORM == 0 ? deq_ab((uint8_t)q, s, z) : deq_cd((uint8_t)q, s, z);
        }
    }
}

// Per-row dequantize, `x * scale + zp` (prefill_kv.rs:123-131): a CTA walks whole rows — the row's scale / zero-point are
// read once, and there is no per-element 64-bit division to find the row (the flat kernel above spends ~100 instructions
// per float4 on it).  Lane i of a warp handles the float4 i, i + 32, ...: every store instruction writes 512 contiguous
// bytes.  (A variant with 16 codes per thread from one wide load was 60 % slower: its stores were 64 bytes apart.)
template <int PACK>
__global__ void __launch_bounds__(256)
decode_rows_kernel(const uint8_t *__restrict__ in, size_t rows, uint32_t dim, const float *__restrict__ row_scales,
                   const float *__restrict__ row_zps, float *__restrict__ out) {
    constexpr int B = PACK == 0 ? 8 : PACK;                 // bits per code in memory
    const uint32_t d4 = dim >> 2;
    const size_t row_bytes = (size_t)dim * B / 8;
    for (size_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const float s = __ldg(row_scales + r), z = __ldg(row_zps + r);
        const uint8_t *irow = in + r * row_bytes;
        float4 *orow = reinterpret_cast<float4 *>(out + r * dim);
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < d4; i += 256) {
            uint32_t c0, c1, c2, c3;
            load_codes4<PACK>(irow, i, c0, c1, c2, c3);
            float4 v;
            v.x = deq_cd((uint8_t)c0, s, z); v.y = deq_cd((uint8_t)c1, s, z);
            v.z = deq_cd((uint8_t)c2, s, z); v.w = deq_cd((uint8_t)c3, s, z);
            stg_stream_f4(orow + i, v);
        }
    }
}

template <int FORM>
__global__ void decode_scalar_kernel(const uint8_t *__restrict__ in, size_t n, float *out, DecArgs a, int pack) {
    float s = a.scale, z = a.zp;
    if (a.dev) { s = __ldg(a.dev); z = __ldg(a.dev + 1); }
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        if (a.row_scales) { s = a.row_scales[j / a.dim]; z = a.row_zps[j / a.dim]; }
        uint32_t q;
        if (pack == 0 || pack == 8) q = in[j];
        else { const int per = 8 / pack; q = (in[j / per] >> ((j % per) * pack)) & ((1u << pack) - 1u); }
        out[j] = FORM == 0 ? deq_ab((uint8_t)q, s, z) : deq_cd((uint8_t)q, s, z);
    }
}

// ------------------------------------------------------------------------------------------
// pack / unpack of existing byte codes (K1): 16 codes per lane per step, 128-bit loads
// ------------------------------------------------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(kThreads)
pack_kernel(const uint8_t *__restrict__ codes, size_t n, uint8_t *__restrict__ packed) {
    constexpr uint32_t M = (1u << BITS) - 1u;
    const size_t n16 = n >> 4;
    const uint4 *c16 = reinterpret_cast<const uint4 *>(codes);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = ldg_stream_u4(c16 + i);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if (BITS == 8) {
            reinterpret_cast<uint4 *>(packed)[i] = v;
        } else if (BITS == 4) {
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                lo |= (((w[0] >> (8 * j)) & M) << (4 * j)) | (((w[1] >> (8 * j)) & M) << (16 + 4 * j));
                hi |= (((w[2] >> (8 * j)) & M) << (4 * j)) | (((w[3] >> (8 * j)) & M) << (16 + 4 * j));
            }
            reinterpret_cast<uint2 *>(packed)[i] = make_uint2(lo, hi);
        } else if (BITS == 2) {
            uint32_t o = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) o |= ((w[q] >> (8 * j)) & M) << (8 * q + 2 * j);
            reinterpret_cast<uint32_t *>(packed)[i] = o;
        } else {
            uint32_t o = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) o |= ((w[q] >> (8 * j)) & M) << (4 * q + j);
            reinterpret_cast<uint16_t *>(packed)[i] = (uint16_t)o;
        }
    }
    // ragged tail: < 16 codes, one thread, byte by byte
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 15)) {
        constexpr int per = 8 / BITS;
        const size_t base = n16 << 4;
        for (size_t b = base / per; b < (n + per - 1) / per; ++b) {
            uint32_t o = 0;
            for (int j = 0; j < per; ++j) {
                size_t idx = b * per + j;
                if (idx < n) o |= (codes[idx] & M) << (j * BITS);
            }
            packed[b] = (uint8_t)o;
        }
    }
}

template <int BITS>
__global__ void __launch_bounds__(kThreads)
unpack_kernel(const uint8_t *__restrict__ packed, size_t n, uint8_t *__restrict__ codes) {
    constexpr uint32_t M = (1u << BITS) - 1u;
    constexpr int per = 8 / BITS;
    const size_t n16 = n >> 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[4];
        if (BITS == 8) {
            uint4 v = ldg_stream_u4(reinterpret_cast<const uint4 *>(packed) + i);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            uint64_t bitsv;
            if (BITS == 4) { uint2 v = __ldg(reinterpret_cast<const uint2 *>(packed) + i); bitsv = ((uint64_t)v.y << 32) | v.x; }
            else if (BITS == 2) bitsv = __ldg(reinterpret_cast<const uint32_t *>(packed) + i);
            else bitsv = __ldg(reinterpret_cast<const uint16_t *>(packed) + i);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t o = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) o |= (uint32_t)((bitsv >> ((4 * q + j) * BITS)) & M) << (8 * j);
                w[q] = o;
            }
        }
        stg_stream_u4(reinterpret_cast<uint4 *>(codes) + i, make_uint4(w[0], w[1], w[2], w[3]));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 15))
        for (size_t j = n16 << 4; j < n; ++j) codes[j] = (uint8_t)((packed[j / per] >> ((j % per) * BITS)) & M);
}

// ------------------------------------------------------------------------------------------
// K6: fused per-token (row) quantizer D.  One CTA per row: the row is read from HBM once
// (kept in registers), min/max reduced with warp shuffles, codes written bit-packed.
// ------------------------------------------------------------------------------------------
template <int T, int NV, int PACK>  // T threads, up to NV float4 per thread cached in registers
__global__ void __launch_bounds__(T)
quant_d_rows_kernel(const float *__restrict__ x, size_t rows, size_t dim, const uint8_t *__restrict__ bits_tab,
                    int nbits, int uniform_bits, uint8_t *__restrict__ out, float *__restrict__ scales,
                    float *__restrict__ zps) {
    __shared__ float smx[T / 32], smn[T / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t d4 = dim >> 2;
    for (size_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const int bits = bits_tab ? (int)bits_tab[r % (size_t)nbits] : uniform_bits;
        const float4 *x4 = reinterpret_cast<const float4 *>(x + r * dim);
        float4 v[NV];
        float mx = -INFINITY, mn = INFINITY;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const size_t i = (size_t)j * T + threadIdx.x;
            if (i < d4) {
                v[j] = ldg_stream_f4(x4 + i);
                mx = fmaxf(mx, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
                mn = fminf(mn, fminf(fminf(v[j].x, v[j].y), fminf(v[j].z, v[j].w)));
            }
        }
        mx = warp_max(mx);
        mn = warp_min(mn);
        if (T > 32) {
            __syncthreads();  // previous row's readers are done with smx/smn
            if (lane == 0) { smx[w] = mx; smn[w] = mn; }
            __syncthreads();
            mx = smx[0]; mn = smn[0];
#pragma unroll
            for (int i = 1; i < T / 32; ++i) { mx = fmaxf(mx, smx[i]); mn = fminf(mn, smn[i]); }
        }
        const float levels = (float)((1u << bits) - 1u);
        const float scale = __fdiv_rn(__fsub_rn(mx, mn), levels);   // prefill_kv.rs:107
        const float zp = mn;                                        // :108
        if (threadIdx.x == 0) { scales[r] = scale; zps[r] = zp; }
        uint8_t *orow = out + (PACK ? r * (dim * PACK / 8) : r * dim);
        // every numerator x - zp of the row lies in [0, max - min] (NaN elements stay NaN -> code 0): one divisor per
        // row, reciprocal and range check hoisted (common.cuh); rows with inf / NaN / extreme scales take __fdiv_rn
        const RowDivisor rd = make_row_divisor(scale, __fsub_rn(mx, mn));
        const unsigned int ulevels = (unsigned int)levels;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const size_t i = (size_t)j * T + threadIdx.x;
            const bool act = i < d4;
            uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            if (act) {
                if (rd.fast) {
                    c0 = clamp_trunc_u8(div_row(rd, __fsub_rn(v[j].x, zp)), ulevels);
                    c1 = clamp_trunc_u8(div_row(rd, __fsub_rn(v[j].y, zp)), ulevels);
                    c2 = clamp_trunc_u8(div_row(rd, __fsub_rn(v[j].z, zp)), ulevels);
                    c3 = clamp_trunc_u8(div_row(rd, __fsub_rn(v[j].w, zp)), ulevels);
                } else {
                    c0 = code_cd(v[j].x, scale, zp, levels); c1 = code_cd(v[j].y, scale, zp, levels);
                    c2 = code_cd(v[j].z, scale, zp, levels); c3 = code_cd(v[j].w, scale, zp, levels);
                }
            }
            store_codes4<PACK>(orow, i, c0, c1, c2, c3, act);
        }
    }
}

// The same quantizer with the rows streamed through a shared-memory ring by bulk copies (cp.async.bulk, one per row,
// completion on an mbarrier): R rows per CTA are in flight while one is being reduced and encoded, instead of the load
// of a row waiting for the stores of the previous one in every CTA.  Rows of >= 4 KB, dim % 4 == 0.
template <int T, int NV, int PACK, int R>
__global__ void __launch_bounds__(T)
quant_d_rows_ring_kernel(const float *__restrict__ x, size_t rows, size_t dim, int bits, uint8_t *__restrict__ out,
                         float *__restrict__ scales, float *__restrict__ zps) {
    extern __shared__ __align__(16) float4 ring[];          // R rows of dim floats
    __shared__ uint64_t bar[R];
    __shared__ float smx[T / 32], smn[T / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t d4 = dim >> 2;
    const uint32_t row_bytes = (uint32_t)(dim * sizeof(float));
    auto bar_addr = [&](int sl) { return (uint32_t)__cvta_generic_to_shared(bar + sl); };
    auto fill = [&](int sl, size_t r) {                     // one elected thread
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_addr(sl)), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((uint32_t)__cvta_generic_to_shared(ring + (size_t)sl * d4)), "l"(x + r * dim), "r"(row_bytes), "r"(bar_addr(sl)) : "memory");
    };
    if (threadIdx.x == 0) {
        for (int sl = 0; sl < R; ++sl) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr(sl)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int sl = 0; sl < R; ++sl) {
            const size_t r = blockIdx.x + (size_t)sl * gridDim.x;
            if (r < rows) fill(sl, r);
        }
    const float levels = (float)((1u << bits) - 1u);
    const unsigned int ulevels = (unsigned int)levels;
    size_t it = 0;
    for (size_t r = blockIdx.x; r < rows; r += gridDim.x, ++it) {
        const int sl = (int)(it % R);
        const uint32_t ph = (uint32_t)((it / R) & 1);
        asm volatile(
            "{\n\t.reg .pred p;\n\tQR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra QR_DONE;\n\tbra QR_WAIT;\n\tQR_DONE:\n\t}\n"
            :: "r"(bar_addr(sl)), "r"(ph) : "memory");
        const float4 *x4 = ring + (size_t)sl * d4;
        float4 v[NV];
        float mx = -INFINITY, mn = INFINITY;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const size_t i = (size_t)j * T + threadIdx.x;
            if (i < d4) {
                v[j] = x4[i];
                mx = fmaxf(mx, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
                mn = fminf(mn, fminf(fminf(v[j].x, v[j].y), fminf(v[j].z, v[j].w)));
            }
        }
        mx = warp_max(mx);
        mn = warp_min(mn);
        __syncthreads();                                    // the row is in registers everywhere; the previous row's smx/smn readers are done
        if (threadIdx.x == 0) {                             // refill the slot with the row R iterations ahead
            const size_t rn = r + (size_t)R * gridDim.x;
            if (rn < rows) fill(sl, rn);
        }
        if (lane == 0) { smx[w] = mx; smn[w] = mn; }
        __syncthreads();
        mx = smx[0]; mn = smn[0];
#pragma unroll
        for (int i = 1; i < T / 32; ++i) { mx = fmaxf(mx, smx[i]); mn = fminf(mn, smn[i]); }
        const float scale = __fdiv_rn(__fsub_rn(mx, mn), levels);   // prefill_kv.rs:107
        const float zp = mn;                                        // :108
        if (threadIdx.x == 0) { scales[r] = scale; zps[r] = zp; }
        uint8_t *orow = out + (PACK ? r * (dim * PACK / 8) : r * dim);
        const RowDivisor rd = make_row_divisor(scale, __fsub_rn(mx, mn));
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const size_t i = (size_t)j * T + threadIdx.x;
            const bool act = i < d4;
            uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            if (act) {
                if (rd.fast) {
                    c0 = clamp_trunc_u8(div_row(rd, __fsub_rn(v[j].x, zp)), ulevels);
                    c1 = clamp_trunc_u8(div_row(rd, __fsub_rn(v[j].y, zp)), ulevels);
                    c2 = clamp_trunc_u8(div_row(rd, __fsub_rn(v[j].z, zp)), ulevels);
                    c3 = clamp_trunc_u8(div_row(rd, __fsub_rn(v[j].w, zp)), ulevels);
                } else {
                    c0 = code_cd(v[j].x, scale, zp, levels); c1 = code_cd(v[j].y, scale, zp, levels);
                    c2 = code_cd(v[j].z, scale, zp, levels); c3 = code_cd(v[j].w, scale, zp, levels);
                }
            }
            store_codes4<PACK>(orow, i, c0, c1, c2, c3, act);
        }
    }
}

// Self-test of div_row against __fdiv_rn.  Case i: divisor d from a hash of i (random significand, exponent in
// [-40, 40]); numerators around every code boundary: +-(RN(c·d) + k ulps) for c = 1..256, k = -4..4, plus random ones.
// Counts quotients whose bits differ (zeros of either sign are equal) — there must be none.
__global__ void selftest_division_kernel(unsigned long long cases, unsigned long long seed, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < cases;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ull;       // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const int ex = (int)((z >> 40) % 81) - 40;
        const float d = __uint_as_float((uint32_t)(127 + ex) << 23 | (uint32_t)(z & 0x7fffffu));
        const RowDivisor rd = make_row_divisor(d, 1.0f);
        for (int c = 1; c <= 256; ++c) {
            const float base = __fmul_rn((float)c, d);
            for (int k = -4; k <= 4; ++k) {
                const float n = __uint_as_float(__float_as_uint(base) + k);
                const float a = div_row(rd, n), b = __fdiv_rn(n, d);
                if (__float_as_uint(a) != __float_as_uint(b) && !(a == 0.f && b == 0.f)) ++bad;
                const float a2 = div_row_checked(rd, -n), b2 = __fdiv_rn(-n, d);      // quantizer B sees negative numerators
                if (__float_as_uint(a2) != __float_as_uint(b2) && !(a2 == 0.f && b2 == 0.f)) ++bad;
            }
        }
        // a random numerator with quotient in (0, 512)
        const float n = __fmul_rn(d, (float)((z >> 8) & 0xffffff) * (512.0f / 16777216.0f));
        const float a = div_row(rd, n), b = __fdiv_rn(n, d);
        if (__float_as_uint(a) != __float_as_uint(b) && !(a == 0.f && b == 0.f)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// generic fallback: any dim, any alignment, one warp per row, two passes over the row
__global__ void quant_d_rows_generic_kernel(const float *__restrict__ x, size_t rows, size_t dim,
                                            const uint8_t *__restrict__ bits_tab, int nbits, int uniform_bits,
                                            uint8_t *__restrict__ out, float *scales, float *zps) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = warp; r < rows; r += nwarps) {
        const int bits = bits_tab ? (int)bits_tab[r % (size_t)nbits] : uniform_bits;
        const float *row = x + r * dim;
        float mx = -INFINITY, mn = INFINITY;
        for (size_t i = lane; i < dim; i += 32) { mx = fmaxf(mx, row[i]); mn = fminf(mn, row[i]); }
        mx = warp_max(mx);
        mn = warp_min(mn);
        const float levels = (float)((1u << bits) - 1u);
        const float scale = __fdiv_rn(__fsub_rn(mx, mn), levels);
        if (lane == 0) { scales[r] = scale; zps[r] = mn; }
        for (size_t i = lane; i < dim; i += 32) out[r * dim + i] = code_cd(row[i], scale, mn, levels);
    }
}

inline int grid_for(const dllm_ctx *ctx, size_t work_items, int per_sm) {
    size_t blocks = (work_items + kThreads - 1) / kThreads;
    size_t cap = (size_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

__global__ void params_from_minmax_kernel(int bits, float *params) {
    float scale, zp;
    params_b(params[2], params[3], bits, &scale, &zp);
    params[0] = scale;
    params[1] = zp;
}

// ==========================================================================================
// launchers
// ==========================================================================================
int32_t k_params_from_minmax(dllm_ctx *ctx, int bits, float *params_dev) {
    params_from_minmax_kernel<<<1, 1, 0, ctx->stream>>>(bits, params_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_minmax(dllm_ctx *ctx, const float *x_dev, size_t n, int bits, float *out_dev) {
    int grid = grid_for(ctx, (n + 15) / 16, 8);
    if (grid > kMaxPartials) grid = kMaxPartials;
    minmax_kernel<<<grid, kThreads, 0, ctx->stream>>>(x_dev, n, aligned16(x_dev) ? 1 : 0, ctx->d_partials,
                                                      ctx->d_ticket, out_dev, bits);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

template <int SCHEME>
static int32_t launch_encode(dllm_ctx *ctx, const float *x, size_t n, uint8_t *out, EncArgs a, int pack) {
    if (n == 0) return DLLM_OK;
    const bool vec = aligned16(x) && (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
    if (!vec) {
        if (pack != 0 && pack != 8) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "packed encode needs 16-byte aligned input");
        encode_scalar_kernel<SCHEME><<<grid_for(ctx, n, 8), kThreads, 0, ctx->stream>>>(x, n, out, a);
        LAUNCH_CHECK(ctx);
        return DLLM_OK;
    }
    const int grid = grid_for(ctx, (n + 3) / 4, 16);
    const bool tail = (n & 3) != 0;
#define ENC_CASE(P)                                                                          \
    if (n >= 4) { encode_kernel<SCHEME, P><<<grid, kThreads, 0, ctx->stream>>>(x, n, out, a); LAUNCH_CHECK(ctx); } \
    if (tail) { encode_tail_kernel<SCHEME, P><<<1, 1, 0, ctx->stream>>>(x, n, out, a); LAUNCH_CHECK(ctx); }
    switch (pack) {
        case 0: case 8: ENC_CASE(0) break;
        case 4: ENC_CASE(4) break;
        case 2: ENC_CASE(2) break;
        case 1: ENC_CASE(1) break;
        default: DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "pack width %d not in {1,2,4,8}", pack);
    }
#undef ENC_CASE
    return DLLM_OK;
}

int32_t k_encode_b(dllm_ctx *ctx, const float *x_dev, size_t n, int bits, int pack, const float *params_dev,
                   float scale, float zp, uint8_t *out_dev) {
    EncArgs a{scale, zp, 0.f, 0.f, (int)((1u << bits) - 1u), params_dev};
    return launch_encode<0>(ctx, x_dev, n, out_dev, a, pack);
}
int32_t k_encode_a(dllm_ctx *ctx, const float *x_dev, size_t n, float scale, float zp, float lo, float hi,
                   uint8_t *out_dev) {
    EncArgs a{scale, zp, lo, hi, 0, nullptr};
    return launch_encode<1>(ctx, x_dev, n, out_dev, a, 0);
}
int32_t k_encode_cd(dllm_ctx *ctx, const float *x_dev, size_t n, int bits, int pack, float scale, float zp,
                    uint8_t *out_dev) {
    EncArgs a{scale, zp, 0.f, (float)(int)((1u << bits) - 1u), 0, nullptr};
    return launch_encode<2>(ctx, x_dev, n, out_dev, a, pack);
}

template <int FORM>
static int32_t launch_decode(dllm_ctx *ctx, const uint8_t *in, size_t n, float *out, DecArgs a, int pack) {
    if (n == 0) return DLLM_OK;
    const bool vec = aligned16(out) && (reinterpret_cast<uintptr_t>(in) & 3u) == 0 &&
                     (a.row_scales == nullptr || (a.dim & 3) == 0);
    if (!vec) {
        decode_scalar_kernel<FORM><<<grid_for(ctx, n, 8), kThreads, 0, ctx->stream>>>(in, n, out, a, pack);
        LAUNCH_CHECK(ctx);
        return DLLM_OK;
    }
    const int grid = grid_for(ctx, (n + 3) / 4, 16);
    switch (pack) {
        case 0: case 8: decode_kernel<FORM, 0><<<grid, kThreads, 0, ctx->stream>>>(in, n, out, a); break;
        case 4: decode_kernel<FORM, 4><<<grid, kThreads, 0, ctx->stream>>>(in, n, out, a); break;
        case 2: decode_kernel<FORM, 2><<<grid, kThreads, 0, ctx->stream>>>(in, n, out, a); break;
        case 1: decode_kernel<FORM, 1><<<grid, kThreads, 0, ctx->stream>>>(in, n, out, a); break;
        default: DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "pack width %d not in {1,2,4,8}", pack);
    }
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_decode_ab(dllm_ctx *ctx, const uint8_t *in_dev, size_t n, int pack, const float *params_dev,
                    float scale, float zp, float *out_dev) {
    DecArgs a{scale, zp, params_dev, nullptr, nullptr, 1};
    return launch_decode<0>(ctx, in_dev, n, out_dev, a, pack);
}
int32_t k_decode_cd(dllm_ctx *ctx, const uint8_t *in_dev, size_t n, int pack, float scale, float zp,
                    const float *row_scales, const float *row_zps, size_t dim, float *out_dev) {
    // whole rows of 16-code units with per-row parameters: the row-wise kernel
    // (a row is a whole number of the 4 / 2 / 1-byte loads of load_codes4 and starts aligned to them)
    if (row_scales && row_zps && dim >= 256 && dim % 16 == 0 && dim <= (1u << 24) && n % dim == 0 && n &&
        (pack == 0 || pack == 8 || pack == 4 || pack == 2 || pack == 1) && aligned16(in_dev) && aligned16(out_dev)) {
        const size_t rows = n / dim;
        const size_t cap = (size_t)ctx->sm_count * 8;
        const int grid = (int)(rows < cap ? rows : cap);
        switch (pack) {
            case 0: case 8: decode_rows_kernel<0><<<grid, 256, 0, ctx->stream>>>(in_dev, rows, (uint32_t)dim, row_scales, row_zps, out_dev); break;
            case 4: decode_rows_kernel<4><<<grid, 256, 0, ctx->stream>>>(in_dev, rows, (uint32_t)dim, row_scales, row_zps, out_dev); break;
            case 2: decode_rows_kernel<2><<<grid, 256, 0, ctx->stream>>>(in_dev, rows, (uint32_t)dim, row_scales, row_zps, out_dev); break;
            default: decode_rows_kernel<1><<<grid, 256, 0, ctx->stream>>>(in_dev, rows, (uint32_t)dim, row_scales, row_zps, out_dev); break;
        }
        LAUNCH_CHECK(ctx);
        return DLLM_OK;
    }
    DecArgs a{scale, zp, nullptr, row_scales, row_zps, dim ? dim : 1};
    return launch_decode<1>(ctx, in_dev, n, out_dev, a, pack);
}

int32_t k_pack(dllm_ctx *ctx, const uint8_t *codes_dev, size_t n, int bits, uint8_t *packed_dev) {
    if (n == 0) return DLLM_OK;
    if (!aligned16(codes_dev) || !aligned16(packed_dev))
        DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "pack needs 16-byte aligned device buffers");
    const int grid = grid_for(ctx, (n + 15) / 16, 16);
    switch (bits) {
        case 1: pack_kernel<1><<<grid, kThreads, 0, ctx->stream>>>(codes_dev, n, packed_dev); break;
        case 2: pack_kernel<2><<<grid, kThreads, 0, ctx->stream>>>(codes_dev, n, packed_dev); break;
        case 4: pack_kernel<4><<<grid, kThreads, 0, ctx->stream>>>(codes_dev, n, packed_dev); break;
        case 8: pack_kernel<8><<<grid, kThreads, 0, ctx->stream>>>(codes_dev, n, packed_dev); break;
        default: DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "pack width %d not in {1,2,4,8}", bits);
    }
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_unpack(dllm_ctx *ctx, const uint8_t *packed_dev, size_t n, int bits, uint8_t *codes_dev) {
    if (n == 0) return DLLM_OK;
    if (!aligned16(codes_dev) || !aligned16(packed_dev))
        DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "unpack needs 16-byte aligned device buffers");
    const int grid = grid_for(ctx, (n + 15) / 16, 16);
    switch (bits) {
        case 1: unpack_kernel<1><<<grid, kThreads, 0, ctx->stream>>>(packed_dev, n, codes_dev); break;
        case 2: unpack_kernel<2><<<grid, kThreads, 0, ctx->stream>>>(packed_dev, n, codes_dev); break;
        case 4: unpack_kernel<4><<<grid, kThreads, 0, ctx->stream>>>(packed_dev, n, codes_dev); break;
        case 8: unpack_kernel<8><<<grid, kThreads, 0, ctx->stream>>>(packed_dev, n, codes_dev); break;
        default: DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "pack width %d not in {1,2,4,8}", bits);
    }
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

template <int T, int NV>
static void launch_d_rows(dllm_ctx *ctx, const float *x, size_t rows, size_t dim, const uint8_t *tab, int nbits,
                          int ubits, int pack, uint8_t *out, float *scales, float *zps) {
    const int per_sm = T >= 256 ? 8 : (T >= 128 ? 16 : 32);
    size_t grid = rows < (size_t)ctx->sm_count * per_sm ? rows : (size_t)ctx->sm_count * per_sm;
    switch (pack) {
        case 4: quant_d_rows_kernel<T, NV, 4><<<(int)grid, T, 0, ctx->stream>>>(x, rows, dim, tab, nbits, ubits, out, scales, zps); break;
        case 2: quant_d_rows_kernel<T, NV, 2><<<(int)grid, T, 0, ctx->stream>>>(x, rows, dim, tab, nbits, ubits, out, scales, zps); break;
        case 1: quant_d_rows_kernel<T, NV, 1><<<(int)grid, T, 0, ctx->stream>>>(x, rows, dim, tab, nbits, ubits, out, scales, zps); break;
        default: quant_d_rows_kernel<T, NV, 0><<<(int)grid, T, 0, ctx->stream>>>(x, rows, dim, tab, nbits, ubits, out, scales, zps); break;
    }
}

template <int PACK>
static bool launch_d_rows_ring(dllm_ctx *ctx, const float *x, size_t rows, size_t dim, int bits, uint8_t *out, float *scales, float *zps) {
    constexpr int T = 256, NV = 4, R = 3;
    const size_t smem = (size_t)R * dim * sizeof(float);
    if (ensure_smem_attr(ctx, quant_d_rows_ring_kernel<T, NV, PACK, R>, 3 * 4096 * 4) != DLLM_OK) return false;
    const size_t per_sm = 4;
    const size_t grid = rows < (size_t)ctx->sm_count * per_sm ? rows : (size_t)ctx->sm_count * per_sm;
    quant_d_rows_ring_kernel<T, NV, PACK, R><<<(int)grid, T, smem, ctx->stream>>>(x, rows, dim, bits, out, scales, zps);
    return true;
}

int32_t k_quant_d_rows(dllm_ctx *ctx, const float *x_dev, size_t rows, size_t dim, const uint8_t *bits_tab_dev,
                       int nbits, int uniform_bits, int pack, uint8_t *out_dev, float *scales_dev, float *zps_dev) {
    if (rows == 0) return DLLM_OK;
    if (pack == 8) pack = 0;
    if (pack && (bits_tab_dev != nullptr || (dim * pack) % 8 != 0 || pack != uniform_bits))
        DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "packed per-row quantize needs one bit width and dim*bits %% 8 == 0");
    const bool fast = aligned16(x_dev) && (dim % 4 == 0) && (dim % 8 == 0 || pack != 1) && dim <= 16384 &&
                      (reinterpret_cast<uintptr_t>(out_dev) & 3u) == 0 && ((pack ? dim * pack / 8 : dim) % 4 == 0);
    if (!fast) {
        if (pack) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "packed per-row quantize: unsupported dim/alignment");
        size_t warps = rows;
        size_t blocks = (warps * 32 + 255) / 256;
        if (blocks > (size_t)ctx->sm_count * 8) blocks = (size_t)ctx->sm_count * 8;
        quant_d_rows_generic_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(x_dev, rows, dim, bits_tab_dev, nbits,
                                                                          uniform_bits, out_dev, scales_dev, zps_dev);
        LAUNCH_CHECK(ctx);
        return DLLM_OK;
    }
    const size_t d4 = dim / 4;
    // long rows with one bit width: the bulk-copy ring variant (more bytes in flight per SM)
    static const bool no_ring = getenv("DLLM_KV_NO_RING") != nullptr;       // experiments only
    if (!no_ring && bits_tab_dev == nullptr && dim >= 1024 && dim <= 4096 && rows >= 64) {
        bool ok = false;
        switch (pack) {
            case 4: ok = launch_d_rows_ring<4>(ctx, x_dev, rows, dim, uniform_bits, out_dev, scales_dev, zps_dev); break;
            case 2: ok = launch_d_rows_ring<2>(ctx, x_dev, rows, dim, uniform_bits, out_dev, scales_dev, zps_dev); break;
            case 1: ok = launch_d_rows_ring<1>(ctx, x_dev, rows, dim, uniform_bits, out_dev, scales_dev, zps_dev); break;
            default: ok = launch_d_rows_ring<0>(ctx, x_dev, rows, dim, uniform_bits, out_dev, scales_dev, zps_dev); break;
        }
        if (ok) { LAUNCH_CHECK(ctx); return DLLM_OK; }
    }
    if (d4 <= 32 * 4) launch_d_rows<32, 4>(ctx, x_dev, rows, dim, bits_tab_dev, nbits, uniform_bits, pack, out_dev, scales_dev, zps_dev);
    else if (d4 <= 128 * 4) launch_d_rows<128, 4>(ctx, x_dev, rows, dim, bits_tab_dev, nbits, uniform_bits, pack, out_dev, scales_dev, zps_dev);
    else if (d4 <= 256 * 4) launch_d_rows<256, 4>(ctx, x_dev, rows, dim, bits_tab_dev, nbits, uniform_bits, pack, out_dev, scales_dev, zps_dev);
    else launch_d_rows<256, 16>(ctx, x_dev, rows, dim, bits_tab_dev, nbits, uniform_bits, pack, out_dev, scales_dev, zps_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

int32_t k_selftest_division(dllm_ctx *ctx, unsigned long long cases, unsigned long long seed, unsigned long long *mismatches_dev) {
    selftest_division_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(cases, seed, mismatches_dev);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}
