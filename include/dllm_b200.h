/*
 * dllm_b200.h — C ABI of libdllm_b200.so: the B200-native (sm_100a) quantized-linear /
 * quantize / KV-quant hot path of zetareticula/diffusion-llm-rs.
 *
 * The reference has no FFI of its own (no extern "C", no -sys crate): its
 * seams are Rust traits and free functions.  Every entry point below names the reference
 * item it replaces (file:line under the reference root) — INTEGRATION.md shows the Rust
 * `extern "C"` block and the trait impls a maintainer adds on top.
 *
 * Conventions
 *  - every function returns an int32 status (DLLM_OK == 0).  Reference panics / assert!s
 *    and QuantizationError variants map onto the DLLM_ERR_* codes; the message is
 *    available from dllm_last_error(ctx).  Nothing throws or aborts across the boundary.
 *  - pointers without a `_dev` suffix in the function name are HOST pointers owned by the
 *    caller (the reference's `&[f32]` / `Vec<u8>`); the call copies in, runs the CUDA
 *    kernels and copies out before returning.  `*_dev` entry points take DEVICE pointers,
 *    enqueue on the context's stream and return without synchronising.
 *  - a dllm_ctx binds one CUDA device and one stream; use one ctx per thread (the Rust
 *    wrapper keeps it behind a Mutex to satisfy `Send + Sync`).
 *  - there is NO CPU fallback: without a CUDA device dllm_ctx_create fails with
 *    DLLM_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef DLLM_B200_H
#define DLLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLLM_API __attribute__((visibility("default")))

/* ---- status codes ---- */
enum {
    DLLM_OK = 0,
    /* quantization/src/error.rs:19-40 (QuantizationError) */
    DLLM_ERR_INVALID_PARAMS = 1,       /* InvalidParams; also assert!(1..=8 bits) quantization.rs:39 */
    DLLM_ERR_UNSUPPORTED = 2,          /* UnsupportedOperation */
    DLLM_ERR_SHAPE = 3,                /* ShapeMismatch */
    DLLM_ERR_CALIBRATION_REQUIRED = 4, /* CalibrationRequired, calibrate.rs:73-75 */
    DLLM_ERR_IO = 5,
    DLLM_ERR_SERIALIZATION = 6,
    DLLM_ERR_INVALID_DATA_FORMAT = 7,
    DLLM_ERR_INDEX = 8,                /* Vec index out of bounds panic, prefill-kvquant-rs/lib.rs:133 */
    /* runtime */
    DLLM_ERR_CUDA = 100,
    DLLM_ERR_NO_DEVICE = 101,
    DLLM_ERR_NCCL = 102,
    DLLM_ERR_OOM = 103,
    DLLM_ERR_NULL = 104
};

/* QuantizationType, quantization/src/quantize.rs:62-78 */
enum { DLLM_QT_INT8 = 0, DLLM_QT_INT4 = 1, DLLM_QT_BINARY = 2, DLLM_QT_FLOAT8 = 3 };
/* BetaSchedule, diffuse-llm-rs/src/lib.rs:109-117 */
enum { DLLM_BETA_LINEAR = 0, DLLM_BETA_QUADRATIC = 1, DLLM_BETA_COSINE = 2 };
/* KV quantization scheme: which reference quantizer a cache entry uses */
enum {
    DLLM_KV_TENSOR_B = 0, /* QuantizedKVCacheEntry::new: one scale/zp per tensor, quantization.rs:140-157 */
    DLLM_KV_ROW_D = 1,    /* per-token row, prefill_kv.rs:104-121 */
    DLLM_KV_FIXED_C = 2   /* fixed scale 1/(2^bits-1), prefill-kvquant-rs/lib.rs:39-46 */
};
/* which kernel a quantized linear runs on */
enum {
    DLLM_PATH_AUTO = 0,
    DLLM_PATH_SIMT = 1,  /* f32 CUDA-core dequant-GEMV (exact f32 dequant, f32 accumulate) */
    DLLM_PATH_UMMA = 2,  /* tcgen05 / TMEM path: bf16 operands dequantized on the fly, f32 accumulate */
    DLLM_PATH_GEMV = 3,  /* 1..16 tokens, HBM-bound: bulk-copy ring + int8 mma.sync (u8 codes x signed-digit block-fixed-point
                            activations per aligned block of 128 k, exact int32 sums per quantization-group block), f32 scale and
                            accumulation across blocks */
    DLLM_PATH_I8 = 4     /* int8 denoise mode (explicit only, never chosen by AUTO): per-tensor quantized weights (group_size 0, the
                            reference's own scheme, quantization.rs:38-85), activations quantized per token to symmetric int8
                            (step max|x_row| / 127 on the bf16-rounded row), tcgen05 kind::i8 with exact s32 sums, and
                            dequantize_tensor's `(q - zp) * scale` (quantization.rs:83) composed with the token step in the epilogue:
                            y = (sum x_q q - zp sum x_q) * (scale * step) + b.  Error against the f64 stack: the activation rounding
                            only (<= step / 2 per element), stated in tests/test_gpu_linear.py */
};

typedef struct dllm_ctx dllm_ctx;
typedef struct dllm_qweight dllm_qweight; /* one quantized [K,N] weight resident in HBM */
typedef struct dllm_model dllm_model;     /* stack of quantized linears behind DiffusionModel */
typedef struct dllm_kv dllm_kv;           /* quantized K/V cache entry resident in HBM */

/* ======================= context ======================= */
DLLM_API const char *dllm_version(void);
DLLM_API int32_t dllm_device_count(void);
DLLM_API int32_t dllm_ctx_create(int32_t device, dllm_ctx **out);
/* adopt an existing cudaStream_t (e.g. the caller's framework stream); not owned */
DLLM_API int32_t dllm_ctx_create_on_stream(int32_t device, void *cuda_stream, dllm_ctx **out);
DLLM_API void dllm_ctx_destroy(dllm_ctx *ctx);
DLLM_API int32_t dllm_ctx_sync(dllm_ctx *ctx);
DLLM_API void *dllm_ctx_stream(dllm_ctx *ctx);
DLLM_API const char *dllm_last_error(const dllm_ctx *ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
DLLM_API uint64_t dllm_launch_count(const dllm_ctx *ctx);
/* number of CUDA-graph launches this ctx has made (each replays a whole captured denoise step) */
DLLM_API uint64_t dllm_graph_replay_count(const dllm_ctx *ctx);
/* bytes this context has copied host->device / device->host so far (what the host-pointer entry points stage; the *_dev
 * entry points copy nothing) */
DLLM_API int32_t dllm_copy_bytes(const dllm_ctx *ctx, uint64_t *h2d, uint64_t *d2h);
DLLM_API int32_t dllm_sm_count(const dllm_ctx *ctx);

/* per-launch CUDA-event timing of the dominant (tcgen05 linear) kernel on the context's stream.
 * begin() arms it; end() synchronises and returns the number of bracketed launches, their summed
 * duration, and the algorithmic flops / bytes they covered (SURVEY.md §8d formulas). */
/* self-test: the per-row quantizer divides by a row's scale with a hoisted reciprocal and exact-FMA corrections;
 * compares that against IEEE division on `cases` divisors x 2305 numerators around every code boundary and returns
 * the number of differing quotients (must be 0) */
DLLM_API int32_t dllm_selftest_division(dllm_ctx *ctx, uint64_t cases, uint64_t seed, uint64_t *mismatches);
DLLM_API int32_t dllm_profile_begin(dllm_ctx *ctx);
DLLM_API int32_t dllm_profile_end(dllm_ctx *ctx, uint64_t *n_launches, double *total_ms, double *total_flops,
                                  double *total_bytes);

/* device / pinned-host memory helpers for callers without a CUDA runtime binding */
DLLM_API int32_t dllm_malloc(dllm_ctx *ctx, size_t bytes, void **dptr);
DLLM_API int32_t dllm_free(dllm_ctx *ctx, void *dptr);
DLLM_API int32_t dllm_memcpy_h2d(dllm_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
DLLM_API int32_t dllm_memcpy_d2h(dllm_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
DLLM_API int32_t dllm_host_alloc(size_t bytes, void **hptr);
DLLM_API int32_t dllm_host_free(void *hptr);

/* ======================= quantizer B =======================
 * diffuse_llm_rs::quantization::quantize_tensor(&[f32], bits) -> (Vec<u8>, f32, f32)
 *   diffuse-llm-rs/src/quantization.rs:38-68.   bits outside 1..=8 -> DLLM_ERR_INVALID_PARAMS. */
DLLM_API int32_t dllm_quantize_tensor(dllm_ctx *ctx, const float *x, size_t n, uint8_t bits,
                                      uint8_t *codes, float *scale, float *zero_point);
/* dequantize_tensor(&[u8], scale, zp) -> Vec<f32>, quantization.rs:81-85 */
DLLM_API int32_t dllm_dequantize_tensor(dllm_ctx *ctx, const uint8_t *codes, size_t n, float scale,
                                        float zero_point, float *out);
/* device-resident forms.  params_dev: float[4] = {scale, zero_point, min, max} written by the
 * quantize call and read by the dequantize call (no host round trip).
 * packed != 0: codes are bit-packed (bits in {1,2,4,8}, layout of dllm_pack). */
DLLM_API int32_t dllm_quantize_tensor_dev(dllm_ctx *ctx, const float *x_dev, size_t n, uint8_t bits,
                                          int32_t packed, uint8_t *codes_dev, float *params_dev);
DLLM_API int32_t dllm_dequantize_tensor_dev(dllm_ctx *ctx, const uint8_t *codes_dev, size_t n,
                                            uint8_t bits, int32_t packed, const float *params_dev,
                                            float *out_dev);
/* the code step alone with caller-supplied (scale, zp): quantization.rs:59-65 and
 * AdaptiveQuantizer::quantize quantization.rs:220-234 (its CKMS sketch stays on the host) */
DLLM_API int32_t dllm_quantize_codes(dllm_ctx *ctx, const float *x, size_t n, uint8_t bits,
                                     float scale, float zero_point, uint8_t *codes);
/* QuantizedTensor::compression_ratio, quantization.rs:120-124 (host arithmetic) */
DLLM_API float dllm_compression_ratio(size_t numel, size_t data_len, uint8_t bits);

/* ======================= quantizer A =======================
 * quantization::DefaultQuantizer::{quantize,dequantize} + quant_utils::{quantize,dequantize}
 *   quantization/src/quantize.rs:93-215, types.rs:71-81.  Through the reference's public API
 *   scale == 1.0 and zero_point == 0 (quantize.rs:98-108); both are parameters here. */
DLLM_API int32_t dllm_quantize_a(dllm_ctx *ctx, const float *x, size_t n, int32_t qtype, float scale,
                                 int32_t zero_point, uint8_t *codes);
DLLM_API int32_t dllm_dequantize_a(dllm_ctx *ctx, const uint8_t *codes, size_t n, float scale,
                                   int32_t zero_point, float *out);
/* CalibrationData::update's min/max fold on the device (calibrate.rs:42-46): out = {min, max} */
DLLM_API int32_t dllm_minmax(dllm_ctx *ctx, const float *x, size_t n, float *min_out, float *max_out);
/* CalibrationData::compute_params, calibrate.rs:72-110 (host arithmetic) */
DLLM_API int32_t dllm_calibrate_params(float min, float max, size_t total_samples, uint8_t bits,
                                       int32_t symmetric, float *scale, int32_t *zero_point);

/* ======================= quantizer C =======================
 * prefill_kvquant_rs::kvquant::BitQuantizer, prefill-kvquant-rs/lib.rs:34-53 */
DLLM_API float dllm_bitquantizer_scale(uint8_t bits); /* 1/((1<<bits)-1), lib.rs:105 */
DLLM_API int32_t dllm_quantize_c(dllm_ctx *ctx, const float *x, size_t n, uint8_t bits, float scale,
                                 float zero_point, uint8_t *codes);
/* BitQuantizer::dequantize `x as f32 * scale + zp` — shared by C and D (lib.rs:49-51,
 * diffusion_prefill/src/prefill_kv.rs:62-66) */
DLLM_API int32_t dllm_dequantize_cd(dllm_ctx *ctx, const uint8_t *codes, size_t n, float scale,
                                    float zero_point, float *out);
/* PrefillKVQuant::quantize_vectors, lib.rs:127-146: vector v uses bits[v % nbits] and the
 * quantizer at index bits/2 of the Vec built from cfg_bits (DLLM_ERR_INDEX when out of range) */
DLLM_API int32_t dllm_kvquant_quantize_vectors(dllm_ctx *ctx, const float *embeddings, size_t nvec,
                                               size_t elems_per_vec, const uint8_t *cfg_bits,
                                               size_t ncfg, const uint8_t *bits, size_t nbits,
                                               uint8_t *codes);

/* ======================= quantizer D (per-token rows) =======================
 * KVCache::compress_vector / FusionANN::quantize: diffusion_prefill/src/prefill_kv.rs:104-121,
 * fusion_ann.rs:53-88.  Row r uses bits[r % nbits] (fusion_ann.rs:58). */
DLLM_API int32_t dllm_quantize_d_rows(dllm_ctx *ctx, const float *x, size_t rows, size_t dim,
                                      const uint8_t *bits, size_t nbits, uint8_t *codes,
                                      float *scales, float *zero_points);
DLLM_API int32_t dllm_dequantize_d_rows(dllm_ctx *ctx, const uint8_t *codes, size_t rows, size_t dim,
                                        const float *scales, const float *zero_points, float *out);
/* fused per-token KV quantize, device resident, one bit width for all rows.
 * packed != 0: row r's codes occupy dim*bits/8 bytes (bits in {1,2,4,8}, dim*bits % 8 == 0). */
DLLM_API int32_t dllm_quantize_d_rows_dev(dllm_ctx *ctx, const float *x_dev, size_t rows, size_t dim,
                                          uint8_t bits, int32_t packed, uint8_t *codes_dev,
                                          float *scales_dev, float *zero_points_dev);
DLLM_API int32_t dllm_dequantize_d_rows_dev(dllm_ctx *ctx, const uint8_t *codes_dev, size_t rows,
                                            size_t dim, uint8_t bits, int32_t packed,
                                            const float *scales_dev, const float *zero_points_dev,
                                            float *out_dev);

/* ======================= pack / unpack =======================
 * The reference stores one code per u8 and only *accounts* for packed sizes
 * ((len*bits+7)/8, quantization.rs:122, lib.rs:284-285).  This build defines the layout:
 * element i -> bits [(i*bits)%8 ..) of byte (i*bits)/8, LSB first; bits in {1,2,4,8}. */
DLLM_API size_t dllm_packed_len(size_t n, uint8_t bits);
DLLM_API int32_t dllm_pack(dllm_ctx *ctx, const uint8_t *codes, size_t n, uint8_t bits, uint8_t *packed);
DLLM_API int32_t dllm_unpack(dllm_ctx *ctx, const uint8_t *packed, size_t n, uint8_t bits, uint8_t *codes);
DLLM_API int32_t dllm_pack_dev(dllm_ctx *ctx, const uint8_t *codes_dev, size_t n, uint8_t bits,
                               uint8_t *packed_dev);
DLLM_API int32_t dllm_unpack_dev(dllm_ctx *ctx, const uint8_t *packed_dev, size_t n, uint8_t bits,
                                 uint8_t *codes_dev);

/* ======================= quantized linear =======================
 * The layer op is SimpleDiffusionModel::forward  `x.dot(&W) + &b`
 * (diffuse-llm-rs/src/lib.rs:806-813) with W = dequantize_tensor(codes) (quantization.rs:81-85).
 * W is [K,N] row-major like the reference's Array2 ([input_dim, output_dim], lib.rs:777).
 * group == 0: one (scale, zp) for the whole tensor (reference behaviour);
 * group  > 0: quantizer B per `group` consecutive k of each column (BASELINE.json configs[0];
 *             QuantizationConfig::group_size = 128, quantization/src/types.rs:126); K % group == 0.
 * bits in {2,4,8} are stored bit-packed in a tile-major layout private to the library. */
DLLM_API int32_t dllm_qweight_quantize(dllm_ctx *ctx, const float *w, size_t K, size_t N, uint8_t bits,
                                       size_t group, const float *bias /* [N] or NULL */,
                                       dllm_qweight **out);
DLLM_API int32_t dllm_qweight_quantize_dev(dllm_ctx *ctx, const float *w_dev, size_t K, size_t N,
                                           uint8_t bits, size_t group, const float *bias_dev,
                                           dllm_qweight **out);
/* adopt codes produced elsewhere (one code per u8, [K,N]); scales/zps are [K/group, N] (or [1]) */
DLLM_API int32_t dllm_qweight_from_codes(dllm_ctx *ctx, const uint8_t *codes, const float *scales,
                                         const float *zero_points, size_t K, size_t N, uint8_t bits,
                                         size_t group, const float *bias, dllm_qweight **out);
/* read back in the canonical layout (parity checks / serialisation) */
DLLM_API int32_t dllm_qweight_export(dllm_ctx *ctx, const dllm_qweight *w, uint8_t *codes,
                                     float *scales, float *zero_points);
DLLM_API int32_t dllm_qweight_info(const dllm_qweight *w, size_t *K, size_t *N, uint8_t *bits,
                                   size_t *group, size_t *packed_bytes);
DLLM_API void dllm_qweight_destroy(dllm_qweight *w);

/* Packed-weights container "DLLMQW01" (SURVEY.md 8f-3): the wire / on-disk form of a quantized linear — what the
 * reference's serde derives on QuantizedTensor { data, shape, params } (quantization/src/types.rs:42-47) would carry, with
 * the codes bit-packed the way the reference only accounts for ((len * bits + 7) / 8, quantization.rs:122).  64-byte header
 * (magic, version, bits, K, N, group, scheme, has_bias, codes_bytes), packed codes (dllm_pack layout at the narrowest width
 * in {1,2,4,8} holding `bits`), f32 scales and zero-points [K/group, N] (or one each), optional f32 bias [N], CRC-32.
 * Errors: DLLM_ERR_IO (file), DLLM_ERR_SERIALIZATION (magic / CRC / truncation), DLLM_ERR_INVALID_DATA_FORMAT (header). */
DLLM_API size_t dllm_qweight_serialized_size(const dllm_qweight *w);
DLLM_API int32_t dllm_qweight_serialize(dllm_ctx *ctx, const dllm_qweight *w, uint8_t *buf, size_t cap, size_t *written);
DLLM_API int32_t dllm_qweight_deserialize(dllm_ctx *ctx, const uint8_t *buf, size_t len, dllm_qweight **out);
DLLM_API int32_t dllm_qweight_save(dllm_ctx *ctx, const dllm_qweight *w, const char *path);
DLLM_API int32_t dllm_qweight_load(dllm_ctx *ctx, const char *path, dllm_qweight **out);

/* y[M,N] = x[M,K] · dequant(W) + b.  path: DLLM_PATH_* */
DLLM_API int32_t dllm_qlinear_forward(dllm_ctx *ctx, const dllm_qweight *w, const float *x, size_t M,
                                      float *y, int32_t path);
DLLM_API int32_t dllm_qlinear_forward_dev(dllm_ctx *ctx, const dllm_qweight *w, const float *x_dev,
                                          size_t M, float *y_dev, int32_t path);
/* Exact integer linear (BASELINE.json north_star: "int8 natively ... 0 for int8->int32"): for a weight quantized per
 * tensor (group_size 0: one scale, one integer zero-point — quantize_tensor's own scheme, quantization.rs:38-68) and
 * int8 activations xq[M,K],   y[M,N] (int32) = sum_k xq[m,k] * (q[k,n] - zp)   with no rounding anywhere (tcgen05
 * kind::i8, u8 codes x s8 activations, s32 accumulators).  The float result of dequantize_tensor composed with the
 * matmul is tensor_scale * x_scale * y.  K % 64 == 0, K <= 65536 (no int32 overflow); else DLLM_ERR_UNSUPPORTED. */
DLLM_API int32_t dllm_qlinear_forward_i8(dllm_ctx *ctx, const dllm_qweight *w, const int8_t *xq, size_t M,
                                         int32_t *y);
DLLM_API int32_t dllm_qlinear_forward_i8_dev(dllm_ctx *ctx, const dllm_qweight *w, const int8_t *xq_dev,
                                             size_t M, int32_t *y_dev);
/* `quantization` crate extension named by BASELINE.json north_star ("quantize/dequantize/matmul"):
 * one-shot dequant-matmul from canonical codes. */
DLLM_API int32_t dllm_dequant_matmul(dllm_ctx *ctx, const uint8_t *codes, const float *scales,
                                     const float *zero_points, size_t K, size_t N, uint8_t bits,
                                     size_t group, const float *bias, const float *x, size_t M,
                                     float *y, int32_t path);

/* ======================= model / denoising loop =======================
 * DiffusionModel::forward(x[batch, feat], t[batch]) -> [batch, feat]  (lib.rs:748-772): x is
 * viewed as [batch*feat/hidden, hidden] tokens and sent through the stack of linears; the
 * stack's last N equals its first K (output shape == input shape, lib.rs:759). */
DLLM_API int32_t dllm_model_create(dllm_ctx *ctx, size_t hidden, dllm_qweight *const *layers,
                                   size_t n_layers, size_t num_timesteps, int32_t beta_kind,
                                   float beta_start, float beta_end, dllm_model **out);
DLLM_API void dllm_model_destroy(dllm_model *m);
DLLM_API int32_t dllm_model_forward(dllm_ctx *ctx, dllm_model *m, const float *x, const size_t *t,
                                    size_t batch, size_t feat, float *noise_pred, int32_t path);
DLLM_API int32_t dllm_model_forward_dev(dllm_ctx *ctx, dllm_model *m, const float *x_dev, size_t batch,
                                        size_t feat, float *noise_pred_dev, int32_t path);
/* create_beta_schedule, lib.rs:554-593 (host arithmetic, f32) */
DLLM_API int32_t dllm_beta_schedule(int32_t kind, size_t T, float beta_start, float beta_end, float *betas);
/* p_sample, lib.rs:1152-1215, noise injected (z may be NULL; ignored when t[0]==0).
 * guard_t0 != 0: rows with 1 - alpha_bar_t == 0 keep x_prev = x_t (DESIGN.md "p_sample"). */
DLLM_API int32_t dllm_p_sample(dllm_ctx *ctx, dllm_model *m, const float *x_t, const float *noise_pred,
                               const float *z, const size_t *t, size_t batch, size_t feat,
                               int32_t guard_t0, float *x_prev);
/* DiffuseLLM::add_noise, lib.rs:1100-1137: noisy = x_start * sqrt(alpha_bar_t) + noise * sqrt(1 - alpha_bar_t), t[b] clamped
 * to T-1 per batch row (:1123).  The noise is an input (the reference's own draw when `None` is an unseeded thread_rng,
 * :1107-1109: parity unpinned); the returned pair's second element is that same noise.  _dev: one timestep for all rows. */
DLLM_API int32_t dllm_add_noise(dllm_ctx *ctx, dllm_model *m, const float *x_start, const float *noise, const size_t *t,
                                size_t batch, size_t feat, float *noisy);
DLLM_API int32_t dllm_add_noise_dev(dllm_ctx *ctx, dllm_model *m, const float *x_start_dev, const float *noise_dev,
                                    size_t t, size_t batch, size_t feat, float *noisy_dev);
/* one denoise step on the device: noise_pred = forward(x); x <- p_sample(x, t, noise_pred, z) */
DLLM_API int32_t dllm_denoise_step_dev(dllm_ctx *ctx, dllm_model *m, float *x_dev, const float *z_dev,
                                       size_t t, size_t batch, size_t feat, int32_t guard_t0,
                                       int32_t path);
/* the same step with HOST buffers: x (in/out) and z are copied in, x_prev is copied out (one iteration
 * of the loop body at lib.rs:924-925 as seen by a host caller) */
DLLM_API int32_t dllm_denoise_step(dllm_ctx *ctx, dllm_model *m, float *x, const float *z, size_t t, size_t batch,
                                   size_t feat, int32_t guard_t0, int32_t path);
/* how the last dllm_denoise_step on this ctx spent its time (CUDA events on the compute stream): host->device copy of x,
 * the step itself (forward + p_sample; the noise upload rides a second stream underneath), device->host copy of x_prev */
DLLM_API int32_t dllm_last_step_breakdown(const dllm_ctx *ctx, float *h2d_ms, float *compute_ms, float *d2h_ms);
/* DiffuseLLM::sample without cache, lib.rs:853-927.  x0: initial noise [batch, feat];
 * noises: [num_steps, batch, feat], slice t used at timestep t (slice 0 unused).  Host pointers. */
DLLM_API int32_t dllm_sample(dllm_ctx *ctx, dllm_model *m, const float *x0, const float *noises,
                             size_t batch, size_t feat, size_t num_steps, int32_t guard_t0,
                             int32_t path, float *x_out);
/* ---- seeded loop: the noise comes from "dllm_noise v1", a counter-based N(0,1) generator defined by this
 * build (csrc/noise.cuh: splitmix64 -> Box-Muller with fixed-order f32 polynomials; element i of stream s depends on
 * (seed, s, i) only; the CPU oracle reproduces it bit for bit).  The reference draws from an unseeded thread_rng
 * (lib.rs:875-878, :1201), which no implementation can reproduce.  Timestep t uses stream t; the initial x uses stream
 * num_steps.  Nothing is uploaded per step, and with use_graph != 0 the step is captured once as a CUDA graph and replayed
 * with t held in device memory (one graph launch per step instead of one launch per kernel). */
DLLM_API int32_t dllm_noise_fill(dllm_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t first, size_t n, float *out);
DLLM_API int32_t dllm_noise_fill_dev(dllm_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t first, size_t n, float *out_dev);
DLLM_API int32_t dllm_denoise_step_seeded_dev(dllm_ctx *ctx, dllm_model *m, float *x_dev, uint64_t seed, size_t t,
                                              size_t batch, size_t feat, int32_t guard_t0, int32_t path);
/* p_sample alone on device tensors (x_prev_dev may alias x_t_dev), with supplied noise or with the seeded generator: for callers
 * that compute noise_pred themselves, e.g. the cached branch of the sampling loop (lib.rs:910-921) */
DLLM_API int32_t dllm_p_sample_dev(dllm_ctx *ctx, dllm_model *m, const float *x_t_dev, const float *noise_pred_dev,
                                   const float *z_dev, size_t t, size_t batch, size_t feat, int32_t guard_t0, float *x_prev_dev);
DLLM_API int32_t dllm_p_sample_seeded_dev(dllm_ctx *ctx, dllm_model *m, const float *x_t_dev, const float *noise_pred_dev,
                                          uint64_t seed, size_t t, size_t batch, size_t feat, int32_t guard_t0,
                                          float *x_prev_dev);
/* x0 == NULL: the initial x is drawn from stream num_steps */
DLLM_API int32_t dllm_sample_seeded(dllm_ctx *ctx, dllm_model *m, const float *x0, uint64_t seed, size_t batch, size_t feat,
                                    size_t num_steps, int32_t guard_t0, int32_t path, int32_t use_graph, float *x_out);
/* x_dev: in = x_T, out = the sample; enqueues only */
DLLM_API int32_t dllm_sample_seeded_dev(dllm_ctx *ctx, dllm_model *m, float *x_dev, uint64_t seed, size_t batch, size_t feat,
                                        size_t num_steps, int32_t guard_t0, int32_t path, int32_t use_graph);
/* progressive decode precision, lib.rs:886-897 (host arithmetic) */
DLLM_API uint8_t dllm_progressive_bits(size_t num_steps, size_t t, uint8_t decode_bits,
                                       uint8_t min_decode_bits, int32_t *is_prefill);

/* ======================= KV cache =======================
 * QuantizedKVCacheEntry::new(keys, values, bits) / dequantize_keys / dequantize_values
 * (quantization.rs:129-176) over [layers, seq, hidden] f32, plus the per-token (D) and
 * fixed-scale (C) schemes of the prefill crates.  Codes are bit-packed when bits in {1,2,4,8}. */
DLLM_API int32_t dllm_kv_quantize(dllm_ctx *ctx, const float *keys, const float *values, size_t layers,
                                  size_t seq, size_t hidden, uint8_t bits, int32_t scheme,
                                  dllm_kv **out);
DLLM_API int32_t dllm_kv_quantize_dev(dllm_ctx *ctx, const float *keys_dev, const float *values_dev,
                                      size_t layers, size_t seq, size_t hidden, uint8_t bits,
                                      int32_t scheme, dllm_kv **out);
/* K / V whose token rows are sharded over the ranks of the context's group (dllm_tp_init): this rank's [layers, seq_local,
 * hidden] slice.  ROW_D / FIXED_C need no exchange; TENSOR_B all-reduces the tensor's min / max (2 floats) so that every
 * rank encodes with the whole tensor's scale / zero-point: the concatenated codes equal a single-GPU quantization. */
DLLM_API int32_t dllm_kv_quantize_sharded_dev(dllm_ctx *ctx, const float *keys_dev, const float *values_dev,
                                              size_t layers, size_t seq_local, size_t hidden, uint8_t bits,
                                              int32_t scheme, dllm_kv **out);
/* re-quantize into an existing entry (KVCacheEntry::update, lib.rs:246-276) */
DLLM_API int32_t dllm_kv_update_dev(dllm_ctx *ctx, dllm_kv *kv, const float *keys_dev,
                                    const float *values_dev);
/* Append-only growth.  The reference re-quantizes the WHOLE cache on every update
 * (KVCacheEntry::update, lib.rs:246-276, called from the sampling loop :913-918).  With per-token parameters (ROW_D,
 * prefill_kv.rs:104-121) or a fixed scale (FIXED_C, prefill-kvquant-rs/lib.rs:39-53) a token's codes depend on that token
 * only, so only the NEW tokens are quantized: keys_new / values_new are [layers, t_new, hidden] and land after the
 * seq tokens each layer already holds.  The result is bit-identical to quantizing the concatenated tensor at once.
 * dllm_kv_create: an empty entry with room for `capacity` tokens per layer; per-tensor entries (TENSOR_B) cannot grow
 * (DLLM_ERR_UNSUPPORTED); appending beyond the capacity is DLLM_ERR_INDEX. */
DLLM_API int32_t dllm_kv_create(dllm_ctx *ctx, size_t layers, size_t capacity, size_t hidden, uint8_t bits,
                                int32_t scheme, dllm_kv **out);
DLLM_API int32_t dllm_kv_append(dllm_ctx *ctx, dllm_kv *kv, const float *keys_new, const float *values_new,
                                size_t t_new);
DLLM_API int32_t dllm_kv_append_dev(dllm_ctx *ctx, dllm_kv *kv, const float *keys_new_dev,
                                    const float *values_new_dev, size_t t_new);
DLLM_API size_t dllm_kv_seq_len(const dllm_kv *kv);
DLLM_API int32_t dllm_kv_dequantize(dllm_ctx *ctx, const dllm_kv *kv, float *keys, float *values);
DLLM_API int32_t dllm_kv_dequantize_dev(dllm_ctx *ctx, const dllm_kv *kv, float *keys_dev,
                                        float *values_dev);
/* one-code-per-u8 view + parameters, for parity checks against the reference layout.
 * scales/zps hold 1 entry (B, C) or layers*seq entries (D) per tensor. */
DLLM_API int32_t dllm_kv_export(dllm_ctx *ctx, const dllm_kv *kv, uint8_t *key_codes,
                                uint8_t *value_codes, float *key_scales, float *key_zps,
                                float *value_scales, float *value_zps);
/* KVCacheEntry::memory_usage accounting, lib.rs:279-302: (len*bits+7)/8 per tensor */
DLLM_API size_t dllm_kv_memory_usage(const dllm_kv *kv);
DLLM_API void dllm_kv_destroy(dllm_kv *kv);

/* Phase-aware cache entry, resident in HBM: KVCacheEntry (lib.rs:122-313) — the f32 keys / values [layers, seq, hidden] plus
 * a prefill-precision and a decode-precision quantized copy (0 bits = no copy) and the phase that selects which one
 * get_keys / get_values decode.  All tensors are device pointers, dense [layers, seq, hidden]; nothing crosses PCIe, so the
 * cached branch of the sampling loop (lib.rs:885-921) runs without host copies.
 *   update_dev       KVCacheEntry::update (:246-276): replaces the tensors, rebuilds BOTH copies (creating a missing one)
 *   append_dev       the same update when only t_new tokens per layer are new: ROW_D / FIXED_C entries quantize just those
 *                    (bit-identical to update_dev with the concatenation); TENSOR_B re-quantizes everything, as the reference does
 *   set_phase        transition_phase (:220-238): entering decode creates the decode copy from the f32 tensors if it is missing
 *   set_decode_bits  progressive precision (:899-903): a new width drops the decode copy until the next update / append
 *   get_dev          get_keys / get_values (:176-205) decoded straight into the consumer's buffers (either may be NULL)
 *   info             len (:305), phase, get_current_quant_bits (:211-217), memory_usage (:279-302) */
typedef struct dllm_kvcache dllm_kvcache;
DLLM_API int32_t dllm_kvcache_create(dllm_ctx *ctx, size_t layers, size_t hidden, size_t capacity, uint8_t prefill_bits,
                                     uint8_t decode_bits, int32_t scheme, dllm_kvcache **out);
DLLM_API int32_t dllm_kvcache_update_dev(dllm_ctx *ctx, dllm_kvcache *kc, const float *keys_dev, const float *values_dev,
                                         size_t seq);
DLLM_API int32_t dllm_kvcache_append_dev(dllm_ctx *ctx, dllm_kvcache *kc, const float *keys_new_dev,
                                         const float *values_new_dev, size_t t_new);
DLLM_API int32_t dllm_kvcache_set_phase(dllm_ctx *ctx, dllm_kvcache *kc, int32_t is_prefill);
DLLM_API int32_t dllm_kvcache_set_decode_bits(dllm_ctx *ctx, dllm_kvcache *kc, uint8_t bits);
DLLM_API int32_t dllm_kvcache_get_dev(dllm_ctx *ctx, const dllm_kvcache *kc, float *keys_out_dev, float *values_out_dev);
DLLM_API int32_t dllm_kvcache_info(const dllm_kvcache *kc, size_t *seq_len, int32_t *is_prefill, uint8_t *current_bits,
                                   size_t *memory_usage);
/* the quantized copy of one phase (NULL if absent), for dllm_kv_export / dllm_kv_dequantize; owned by the entry */
DLLM_API const dllm_kv *dllm_kvcache_copy(const dllm_kvcache *kc, int32_t prefill);
DLLM_API void dllm_kvcache_destroy(dllm_kvcache *kc);

/* ======================= multi-GPU (one process per GPU) =======================
 * Tensor parallel linears: column-parallel (split N) followed by row-parallel (split K) with
 * one NCCL collective per pair at the layer boundary.  Rank 0 creates the id,
 * the launcher broadcasts its 128 bytes (e.g. torch.distributed), every rank calls tp_init. */
DLLM_API int32_t dllm_tp_unique_id(uint8_t id_out[128]);
DLLM_API int32_t dllm_tp_init(dllm_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t world);
DLLM_API int32_t dllm_tp_finalize(dllm_ctx *ctx);
/* overlap of the row-parallel all-reduces with the GEMMs: the tokens are cut into `chunks` pieces (0 = default 2; 1 = no
 * overlap, collectives on the compute stream) whose all-reduces run on a second stream under the next piece's GEMMs, which
 * leave `reserve_sms` SMs (-1 = default 8) to the collective.  skip_comm != 0 runs the sharded stack WITHOUT its collectives
 * (wrong results; measurement of the exposed collective time only). */
DLLM_API int32_t dllm_tp_configure(dllm_ctx *ctx, int32_t chunks, int32_t reserve_sms, int32_t skip_comm);
/* This library's own all-reduce over NVLink peer memory (tp.cu: two-shot reduce with peer loads / stores and flag barriers).
 * Every rank of the group allocates an arena of `arena_bytes` and maps all the others' through CUDA IPC (the handles travel
 * over the communicator of dllm_tp_init, so no extra rendezvous is needed); a tensor-parallel tcgen05 stack whose two activation
 * buffers fit into it (2 x tokens x widest shard x 2 bytes) then reduces its row-parallel partial sums in place with that
 * kernel instead of ncclAllReduce.  Collective: every rank calls it, in the same order relative to other collectives.
 * arena_bytes == 0 releases the arena (back to ncclAllReduce).
 * DLLM_ERR_UNSUPPORTED (on every rank alike) if some rank cannot map its peers: the NCCL path stays in use.
 * dllm_tp_p2p_status: this rank's arena (a tensor placed in it is reduced by the kernel when handed to dllm_tp_allreduce_dev),
 * its size (0 = off), all-reduces done by the kernel so far, and whether a barrier ever timed out (a peer died: results
 * after that are invalid). */
DLLM_API int32_t dllm_tp_p2p_enable(dllm_ctx *ctx, size_t arena_bytes);
DLLM_API int32_t dllm_tp_p2p_status(dllm_ctx *ctx, void **arena_dev, size_t *arena_bytes, uint64_t *allreduces,
                                    uint32_t *timed_out);
/* sum-all-reduce of a device f32 buffer over the TP group (row-parallel partial sums) */
DLLM_API int32_t dllm_tp_allreduce_dev(dllm_ctx *ctx, float *buf_dev, size_t n);
/* all-gather of column shards: in [M, N/world] per rank -> out [M, N] */
DLLM_API int32_t dllm_tp_allgather_cols_dev(dllm_ctx *ctx, const float *in_dev, size_t M,
                                            size_t n_local, float *out_dev);
/* model whose layers alternate column-/row-parallel shards held by this rank:
 * parallel[i] in {0 replicated, 1 column (split N), 2 row (split K)} */
DLLM_API int32_t dllm_model_set_parallel(dllm_ctx *ctx, dllm_model *m, const int32_t *parallel,
                                         size_t n_layers);

#ifdef __cplusplus
}
#endif
#endif /* DLLM_B200_H */
