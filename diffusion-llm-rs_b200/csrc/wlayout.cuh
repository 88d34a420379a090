// wlayout.cuh — the tile-major packed weight layout (private to the library).
//
// W is [K,N] row-major in the reference ([input_dim, output_dim], diffuse-llm-rs/src/lib.rs:777).
// In HBM the codes of a quantized weight live as tiles of 128 output columns (n) x 64 inputs (k):
//
//   tile(nt, kb) at  packed + (nt * k_blocks + kb) * tile_bytes          (k-minor: a CTA that walks
//                                                                          K reads contiguous memory)
//   tile_bytes = 128 * 64 * bits / 8 ;  CH = bits/2 chunks of 16 bytes per column per tile
//   chunk j of column n_local at  tile + (j * 128 + n_local) * 16        (lane <-> column: a warp's
//                                                                          128-bit loads are contiguous)
//   a chunk holds EPC = 128/bits consecutive k, as 4 words of EPW = 32/bits codes; inside a word
//   codes are interleaved so that one shift+mask yields the pair (k, k+1) in the low / high 16 bits:
//       4-bit: code i of the word sits at nibble  (i>>1) + 4*(i&1)
//       2-bit: code i of the word sits at 2-bit field (i>>1) + 8*(i&1)
//       8-bit: natural byte order
//
// Padding: N is padded to a multiple of 128 with scale 0 / code 0; K to a multiple of 64 with
// code == zero-point (so the padded weights dequantize to exactly 0).
#pragma once
#include <stdint.h>

#define WL_TILE_N 128
#define WL_TILE_K 64

__host__ __device__ inline int wl_container_bits(int bits) { return bits <= 2 ? 2 : (bits <= 4 ? 4 : 8); }
__host__ __device__ inline size_t wl_tile_bytes(int cbits) { return (size_t)WL_TILE_N * WL_TILE_K * cbits / 8; }

// bit offset of code i (0..EPW-1) inside its 32-bit word
template <int CB>
__host__ __device__ inline int wl_bitpos(int i) {
    if (CB == 4) return 4 * ((i >> 1) + 4 * (i & 1));
    if (CB == 2) return 2 * ((i >> 1) + 8 * (i & 1));
    return 8 * i;
}

#ifdef __CUDACC__
#include <cuda_fp16.h>
// decode one 32-bit word into its EPW codes as exact floats, in k order.
template <int CB>
__device__ __forceinline__ void wl_decode_word_f32(uint32_t w, float *q) {
    const __half2 k1024 = __halves2half2(__ushort_as_half(0x6400), __ushort_as_half(0x6400));
    if (CB == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t pr = ((w >> (4 * i)) & 0x000f000fu) | 0x64006400u;   // fp16 {1024+q_even, 1024+q_odd}
            __half2 h = __hsub2(*reinterpret_cast<__half2 *>(&pr), k1024);
            float2 f = __half22float2(h);
            q[2 * i] = f.x; q[2 * i + 1] = f.y;
        }
    } else if (CB == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t pr = ((w >> (2 * i)) & 0x00030003u) | 0x64006400u;
            __half2 h = __hsub2(*reinterpret_cast<__half2 *>(&pr), k1024);
            float2 f = __half22float2(h);
            q[2 * i] = f.x; q[2 * i + 1] = f.y;
        }
    } else {
        uint32_t p01 = __byte_perm(w, 0x64646464u, 0x5140);
        uint32_t p23 = __byte_perm(w, 0x64646464u, 0x5342);
        float2 a = __half22float2(__hsub2(*reinterpret_cast<__half2 *>(&p01), k1024));
        float2 b = __half22float2(__hsub2(*reinterpret_cast<__half2 *>(&p23), k1024));
        q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y;
    }
}
#endif
