// kernels.h — host launchers of the CUDA kernels (internal; the public surface is include/dllm_b200.h)
#pragma once
#include <stddef.h>
#include <stdint.h>

struct dllm_ctx;
struct dllm_qweight;

// ---- quant_kernels.cu ----
// out_dev[4] = {scale_B, zp_B, min, max}; bits == 0 computes min/max only
int32_t k_minmax(dllm_ctx *ctx, const float *x_dev, size_t n, int bits, float *out_dev);
// pack: 0 = one code per u8; 1/2/4/8 = bit-packed
int32_t k_encode_b(dllm_ctx *ctx, const float *x_dev, size_t n, int bits, int pack, const float *params_dev,
                   float scale, float zp, uint8_t *out_dev);
int32_t k_encode_a(dllm_ctx *ctx, const float *x_dev, size_t n, float scale, float zp, float lo, float hi,
                   uint8_t *out_dev);
int32_t k_encode_cd(dllm_ctx *ctx, const float *x_dev, size_t n, int bits, int pack, float scale, float zp,
                    uint8_t *out_dev);
int32_t k_decode_ab(dllm_ctx *ctx, const uint8_t *in_dev, size_t n, int pack, const float *params_dev,
                    float scale, float zp, float *out_dev);
int32_t k_decode_cd(dllm_ctx *ctx, const uint8_t *in_dev, size_t n, int pack, float scale, float zp,
                    const float *row_scales, const float *row_zps, size_t dim, float *out_dev);
int32_t k_pack(dllm_ctx *ctx, const uint8_t *codes_dev, size_t n, int bits, uint8_t *packed_dev);
int32_t k_unpack(dllm_ctx *ctx, const uint8_t *packed_dev, size_t n, int bits, uint8_t *codes_dev);
int32_t k_quant_d_rows(dllm_ctx *ctx, const float *x_dev, size_t rows, size_t dim, const uint8_t *bits_tab_dev,
                       int nbits, int uniform_bits, int pack, uint8_t *out_dev, float *scales_dev, float *zps_dev);

// div_row (hoisted-reciprocal division of the row quantizer) against __fdiv_rn: counts differing quotients
int32_t k_selftest_division(dllm_ctx *ctx, unsigned long long cases, unsigned long long seed, unsigned long long *mismatches_dev);

// ---- weight_kernels.cu ----
// per-(group, column) quantizer-B parameters of W[K,N] (row-major f32)
int32_t k_wparams_grouped(dllm_ctx *ctx, const float *w_dev, size_t K, size_t N, size_t group, int bits,
                          float *scales_dev, float *zps_dev);
// fill scales/zps [1, N] from the per-tensor params {scale, zp} on the device
int32_t k_wparams_broadcast(dllm_ctx *ctx, const float *params_dev, size_t N, float *scales_dev, float *zps_dev);
// quantize (from f32) or adopt (from u8 codes) into the tile-major packed layout of qw
int32_t k_wpack_from_f32(dllm_ctx *ctx, const float *w_dev, dllm_qweight *qw);
int32_t k_wpack_from_codes(dllm_ctx *ctx, const uint8_t *codes_dev, dllm_qweight *qw);
// derive the packed dequant operands of the tcgen05 path from scales / zps (call after they are final)
int32_t k_wdq_params(dllm_ctx *ctx, dllm_qweight *qw);
int32_t k_wexport_codes(dllm_ctx *ctx, const dllm_qweight *qw, uint8_t *codes_dev);

// ---- gemv_simt.cu ----
// y[M,N] = x[M,K] · dequant(W) + b, f32 CUDA cores, any M (tiled by 8 rows)
int32_t k_qlinear_simt(dllm_ctx *ctx, const dllm_qweight *qw, const float *x_dev, size_t M, float *y_dev);

// ---- gemv_mma.cu ----
// HBM-bound path for 1..16 tokens: bulk-copy ring + int8 mma.sync (u8 codes x signed-digit activations)
int32_t k_qlinear_gemv(dllm_ctx *ctx, const dllm_qweight *qw, const float *x_dev, size_t M, float *y_dev);
bool k_gemv_supported(const dllm_qweight *qw, size_t M);

// ---- umma_gemm.cu ----
// tcgen05 path.  x_bf16_dev: [M, K] bf16 row-major.  out_f32 / out_bf16: either may be null.
int32_t k_qlinear_umma(dllm_ctx *ctx, const dllm_qweight *qw, const void *x_bf16_dev, size_t M,
                       float *y_f32_dev, void *y_bf16_dev);
bool k_umma_supported(const dllm_qweight *qw, size_t M);
// exact int8 x u8-codes -> int32 linear on tcgen05 kind::i8 (per-tensor quantized weights)
// the same linear as a tensor-parallel ROW layer with the reduce-scatter fused into the epilogue: tile rows go straight into the
// receive buffer [M, N] bf16 of the rank that owns those tokens (recv[r]: rank r's buffer as mapped here), row block `rank`
struct UmmaRs { void *recv[8]; int world, rank; size_t rows; };
bool k_umma_rs_supported(const dllm_ctx *ctx, const dllm_qweight *qw, size_t M, int world);
// can the next k_qlinear_umma(ctx, qw, x, M, nullptr, y_bf16) run with ctx->gate_armed (flag-gated activation loads)?
bool k_umma_gate_supported(const dllm_ctx *ctx, const dllm_qweight *qw, size_t M, int world, const void *y_bf16);
int32_t k_qlinear_umma_rs(dllm_ctx *ctx, const dllm_qweight *qw, const void *x_bf16_dev, size_t M, const UmmaRs *rs);
int32_t k_qlinear_umma_i8(dllm_ctx *ctx, const dllm_qweight *qw, const int8_t *xq_dev, size_t M, int32_t *y_i32_dev);
// ... with the dequantization fused into the epilogue (int8 denoise stack): y = (sum - zp rowsum[m]) * rowscale[m] + bias[n]
int32_t k_qlinear_umma_i8_deq(dllm_ctx *ctx, const dllm_qweight *qw, const int8_t *xq_dev, const int32_t *rowsum_dev,
                              const float *rowscale_dev, size_t M, float *y_f32_dev, void *y_bf16_dev);
// per-token symmetric int8 quantization of bf16 activations [M, K] (K % 8 == 0): xq = rint(x * 127 / max|x_row|), rowscale[m] =
// wscale * max|x_row| / 127 (1 for an all-zero row), rowsum[m] = sum of the row's codes
int32_t k_rowquant_i8(dllm_ctx *ctx, const void *x_bf16_dev, size_t M, size_t K, float wscale, int8_t *xq_dev, float *rowscale_dev,
                      int32_t *rowsum_dev);
bool k_umma_i8_supported(const dllm_qweight *qw, size_t M);

// ---- sample_kernels.cu ----
int32_t k_f32_to_bf16(dllm_ctx *ctx, const float *in_dev, size_t n, void *out_bf16_dev);
int32_t k_bf16_to_f32(dllm_ctx *ctx, const void *in_bf16_dev, size_t n, float *out_dev);
// x_prev = (c1*x + c2*pred) + sd*z per row (coefficients per batch row on the device)
// coefficient table rows are {c1, c2, std, degenerate}; row of batch element b = rowmap ? rowmap[b] : row
int32_t k_p_sample(dllm_ctx *ctx, const float *x_dev, const float *pred_dev, const float *z_dev,
                   const float *coef_table_dev, const int *rowmap_dev, int row, size_t batch, size_t feat,
                   float *out_dev);
// noisy = x * tab[t][0] + noise * tab[t][1] per batch row (add_noise, lib.rs:1131-1133); rows as in k_p_sample
int32_t k_add_noise(dllm_ctx *ctx, const float *x_dev, const float *noise_dev, const float *tab_dev, const int *rowmap_dev,
                    int row, size_t batch, size_t feat, float *out_dev);
// "dllm_noise v1" (noise.cuh): out[j] = element i0 + j of stream `stream` under `seed`
int32_t k_noise_fill(dllm_ctx *ctx, unsigned long long seed, unsigned long long stream, unsigned long long i0, size_t n,
                     float *out_dev);
// p_sample with the step's noise generated in the kernel (stream = t); state_dev (may be null) = {int t; int pad; u64 seed}
// read on the device instead of the t / seed arguments (CUDA-graph replay); k_sample_state_step decrements its t
int32_t k_p_sample_seeded(dllm_ctx *ctx, const float *x_dev, const float *pred_dev, const float *coef_table_dev,
                          const void *state_dev, int t, unsigned long long seed, int T, size_t total, float *out_dev);
int32_t k_sample_state_step(dllm_ctx *ctx, void *state_dev);
// quantizer B's {scale, zp} recomputed on the device from params[2..3] = {min, max} (after a cross-GPU min / max all-reduce)
int32_t k_params_from_minmax(dllm_ctx *ctx, int bits, float *params_dev);
