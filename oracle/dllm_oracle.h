/*
 * dllm_oracle.h — CPU restatement (plain C, f32 where the reference is f32) of the
 * quantized-linear / quantize / KV-quant hot path of zetareticula/diffusion-llm-rs.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The shipped library
 * (libdllm_b200.so) never links, loads or calls anything in this directory.
 *
 * PARITY STATUS: the reference is Rust, does not compile and no Rust
 * toolchain exists in this image, so the reference itself cannot be run here.  The
 * oracle is pinned by (1) every known-answer vector the reference's own unit tests hold
 * for this path (tests/golden/reference_kat.json; the reference asserts only 0.1
 * tolerances / shapes on them), (2) an independent numpy restatement of the same source
 * lines (tests/np_restatement.py) that must agree bit-for-bit.  Exact codes are pinned by
 * the reference's SOURCE TEXT, not by reference test assertions.  The matmul
 * (ndarray -> matrixmultiply, un-vendored, no Cargo.lock), the noise RNG (thread_rng) and
 * AdaptiveQuantizer's CKMS sketch (quantiles 0.7, un-vendored) are "parity unpinned":
 * see DESIGN.md §Oracle.
 *
 * All citations are file:line under /root/reference.
 */
#ifndef DLLM_ORACLE_H
#define DLLM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes shared with include/dllm_b200.h */
#define ORC_OK 0
#define ORC_ERR_INVALID_PARAMS 1   /* reference: assert!/panic or QuantizationError::InvalidParams */
#define ORC_ERR_SHAPE 3
#define ORC_ERR_INDEX 8            /* reference: Vec index out of bounds panic */

/* QuantizationType, quantization/src/quantize.rs:62-78 */
enum { ORC_QT_INT8 = 0, ORC_QT_INT4 = 1, ORC_QT_BINARY = 2, ORC_QT_FLOAT8 = 3 };
/* BetaSchedule, diffuse-llm-rs/src/lib.rs:109-117 */
enum { ORC_BETA_LINEAR = 0, ORC_BETA_QUADRATIC = 1, ORC_BETA_COSINE = 2 };

/* ---- Quantizer B: diffuse-llm-rs/src/quantization.rs:38-68, 81-85 ---- */
int32_t orc_quantize_tensor(const float *x, size_t n, uint8_t bits,
                            uint8_t *codes, float *scale, float *zero_point);
void orc_dequantize_tensor(const uint8_t *codes, size_t n, float scale, float zero_point,
                           float *out);
/* the code step alone, given (scale, zp): quantization.rs:59-65 and AdaptiveQuantizer :225-231 */
void orc_quantize_codes_b(const float *x, size_t n, uint8_t bits, float scale, float zp,
                          uint8_t *codes);
/* compression_ratio, quantization.rs:120-124 */
float orc_compression_ratio(size_t numel, size_t data_len, uint8_t bits);

/* Extension (not in the reference, BASELINE.json configs[0]): quantizer B applied to each
 * group of `group` consecutive k of column n of W[K,N] row-major.  scales/zps are [K/group, N]. */
int32_t orc_quantize_weight_grouped(const float *w, size_t K, size_t N, uint8_t bits,
                                    size_t group, uint8_t *codes /*[K,N]*/,
                                    float *scales, float *zps);
void orc_dequantize_weight_grouped(const uint8_t *codes, size_t K, size_t N, size_t group,
                                   const float *scales, const float *zps, float *w);

/* ---- Quantizer A: quantization/src/quantize.rs:111-154, 172-184; types.rs:71-81 ---- */
int32_t orc_quantize_a(const float *x, size_t n, int32_t qtype, float scale, int32_t zero_point,
                       uint8_t *codes);
void orc_dequantize_a(const uint8_t *codes, size_t n, float scale, int32_t zero_point, float *out);
/* CalibrationData::compute_params, quantization/src/calibrate.rs:72-110.
 * returns ORC_ERR_INVALID_PARAMS for total_samples == 0 (CalibrationRequired). */
int32_t orc_calibrate_params(float min, float max, size_t total_samples, uint8_t bits,
                             int32_t symmetric, float *scale, int32_t *zero_point);

/* ---- Quantizer C: prefill-kvquant-rs/lib.rs:34-53, 101-110, 127-146 ---- */
float orc_bitquantizer_scale_c(uint8_t bits); /* 1.0 / ((1<<bits)-1) as f32, :105 */
int32_t orc_quantize_c(const float *x, size_t n, uint8_t bits, float scale, float zp,
                       uint8_t *codes);
void orc_dequantize_cd(const uint8_t *codes, size_t n, float scale, float zp, float *out);
/* quantize_vectors: vector v uses bits[v % nbits] and quantizers[bits/2] (:132-133) where
 * quantizers[i] was built from cfg_bits[i].  ORC_ERR_INDEX when bits/2 >= ncfg. */
int32_t orc_kvquant_quantize_vectors(const float *emb, size_t nvec, size_t elems_per_vec,
                                     const uint8_t *cfg_bits, size_t ncfg,
                                     const uint8_t *bits, size_t nbits, uint8_t *codes);

/* ---- Quantizer D: diffusion_prefill/src/prefill_kv.rs:48-67, 104-132; fusion_ann.rs:53-88 ---- */
int32_t orc_quantize_d_rows(const float *x, size_t rows, size_t dim,
                            const uint8_t *bits, size_t nbits /* row r uses bits[r % nbits] */,
                            uint8_t *codes, float *scales, float *zps);
void orc_dequantize_d_rows(const uint8_t *codes, size_t rows, size_t dim,
                           const float *scales, const float *zps, float *out);

/* ---- pack / unpack (layout defined by this build) ----
 * element i occupies bits [(i*bits)%8, (i*bits)%8+bits) of byte (i*bits)/8, LSB first.
 * bits in {1,2,4,8}.  packed length = (n*bits+7)/8 (matches quantization.rs:122). */
size_t orc_packed_len(size_t n, uint8_t bits);
int32_t orc_pack(const uint8_t *codes, size_t n, uint8_t bits, uint8_t *packed);
int32_t orc_unpack(const uint8_t *packed, size_t n, uint8_t bits, uint8_t *codes);

/* ---- Linear: diffuse-llm-rs/src/lib.rs:806-813 (x.dot(W) + b) ---- */
void orc_linear_f32(const float *x, const float *w, const float *bias,
                    size_t M, size_t K, size_t N, float *y);   /* sequential-k f32 accumulate */
void orc_linear_f64(const float *x, const float *w, const float *bias,
                    size_t M, size_t K, size_t N, double *y);  /* f64-accumulated truth */
/* the same, multi-threaded over output columns (NOT reference behaviour; "all cores" CPU arm) */
void orc_linear_f32_mt(const float *x, const float *w, const float *bias,
                       size_t M, size_t K, size_t N, float *y, int threads);
/* exact integer path: sum_k (qw - zw)(qx - zx) -> int64 (SURVEY.md §7 "int8 path") */
void orc_linear_i8_exact(const uint8_t *qx, int32_t zx, const uint8_t *qw, int32_t zw,
                         size_t M, size_t K, size_t N, int64_t *acc);

/* ---- schedules and p_sample: lib.rs:554-593, 1152-1215 ---- */
int32_t orc_beta_schedule(int32_t kind, size_t T, float beta_start, float beta_end, float *betas);
/* coefficient triple for timestep t (clamped to T-1):  c1, c2, std  (lib.rs:1160-1192,1208-1209).
 * Decision: `alphas` at :1191 is read as alpha_t. */
void orc_p_sample_coeffs(const float *betas, size_t T, size_t t, float *c1, float *c2, float *std);
/* x_prev = (c1*x_t + c2*noise_pred) + std*z ; z == NULL or t[0]==0 -> zeros (lib.rs:1195-1212).
 * guard_t0 != 0: rows whose (1 - alpha_bar_t) == 0 keep x_prev = x_t (documented guard). */
void orc_p_sample(const float *x_t, const float *noise_pred, const float *z,
                  const size_t *t, size_t batch, size_t feat,
                  const float *betas, size_t T, int32_t guard_t0, float *x_prev);
/* add_noise, lib.rs:1100-1137: noisy = x_start * sqrt(alpha_bar_t) + noise * sqrt(1 - alpha_bar_t), t clamped to T-1 per row */
void orc_add_noise(const float *x_start, const float *noise, const size_t *t, size_t batch, size_t feat,
                   const float *betas, size_t T, float *noisy);
/* "dllm_noise v1": elements [i0, i0+n) of stream `stream` under `seed` (spec: diffusion-llm-rs_b200/csrc/noise.cuh).
 * Not a reference function: the reference's noise (thread_rng) is unseeded; this is the generator the seeded loop uses. */
void orc_noise_normal(uint64_t seed, uint64_t stream, uint64_t i0, size_t n, float *out);
/* progressive decode bits, lib.rs:886-897. returns target bits; *is_prefill set. */
uint8_t orc_progressive_bits(size_t num_steps, size_t t, uint8_t decode_bits, uint8_t min_bits,
                             int32_t *is_prefill);

#ifdef __cplusplus
}
#endif
#endif
