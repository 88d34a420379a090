// api.cu — the C ABI (include/dllm_b200.h): context, host-pointer entry points that stage
// through device buffers, device-pointer entry points that only enqueue, the quantized weight
// object, the layer stack behind DiffusionModel, p_sample / sample, and the KV cache entry.
// There is no CPU path anywhere in this file: every compute call launches CUDA kernels.
#include <math.h>

#include <new>

#include "common.cuh"
#include "kernels.h"
#include "wlayout.cuh"

// every entry point runs on the context's device, whatever the calling thread's current device is
#define CTX_CHECK(ctx)                                                        \
    do {                                                                      \
        if (!(ctx)) return DLLM_ERR_NULL;                                     \
        (ctx)->err[0] = 0;                                                    \
        if (cudaSetDevice((ctx)->device) != cudaSuccess) {                    \
            cudaGetLastError();                                               \
            DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cudaSetDevice(%d) failed", (ctx)->device); \
        }                                                                     \
    } while (0)

#define ARG_CHECK(ctx, cond, code, ...)                  \
    do {                                                 \
        if (!(cond)) DLLM_FAIL(ctx, code, __VA_ARGS__);  \
    } while (0)

static inline bool pack_width_ok(int bits) { return bits == 1 || bits == 2 || bits == 4 || bits == 8; }

// ==========================================================================================
// context
// ==========================================================================================
extern "C" {

const char *dllm_version(void) { return "dllm_b200 0.1.0 (sm_100a)"; }

int32_t dllm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int32_t ctx_init(int32_t device, void *stream, bool adopt, dllm_ctx **out) {
    if (!out) return DLLM_ERR_NULL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return DLLM_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) return DLLM_ERR_NO_DEVICE;
    dllm_ctx *ctx = new (std::nothrow) dllm_ctx();
    if (!ctx) return DLLM_ERR_OOM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DLLM_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return DLLM_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major != 10) {   // kernels are sm_100a only: fail loudly, never fall back
        delete ctx;
        return DLLM_ERR_NO_DEVICE;
    }
    if (adopt) {
        ctx->stream = (cudaStream_t)stream;
        ctx->owns_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return DLLM_ERR_CUDA; }
        ctx->owns_stream = true;
    }
    bool ok = cudaMalloc(&ctx->d_partials, sizeof(float) * 2 * kMaxPartials) == cudaSuccess &&
              cudaMalloc(&ctx->d_ticket, sizeof(unsigned int)) == cudaSuccess &&
              cudaMalloc(&ctx->d_params, sizeof(float) * 16) == cudaSuccess &&
              cudaMallocHost(&ctx->h_params, sizeof(float) * 16) == cudaSuccess &&
              cudaMemset(ctx->d_ticket, 0, sizeof(unsigned int)) == cudaSuccess;
    if (!ok) { dllm_ctx_destroy(ctx); return DLLM_ERR_CUDA; }
    *out = ctx;
    return DLLM_OK;
}

int32_t dllm_ctx_create(int32_t device, dllm_ctx **out) { return ctx_init(device, nullptr, false, out); }
int32_t dllm_ctx_create_on_stream(int32_t device, void *cuda_stream, dllm_ctx **out) {
    return ctx_init(device, cuda_stream, true, out);
}

int32_t dllm_tp_finalize(dllm_ctx *ctx);

void dllm_ctx_destroy(dllm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    dllm_tp_finalize(ctx);
    for (auto &b : ctx->ws) if (b.p) cudaFree(b.p);
    for (auto &b : ctx->act) if (b.p) cudaFree(b.p);
    if (ctx->tp_ws.p) cudaFree(ctx->tp_ws.p);
    for (auto e : ctx->tp_ev) cudaEventDestroy(e);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    for (auto e : ctx->ev_step) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->lin_ws.p) cudaFree(ctx->lin_ws.p);
    if (ctx->lin_flags.p) cudaFree(ctx->lin_flags.p);
    if (ctx->gemv_tickets.p) cudaFree(ctx->gemv_tickets.p);
    if (ctx->d_partials) cudaFree(ctx->d_partials);
    if (ctx->d_ticket) cudaFree(ctx->d_ticket);
    if (ctx->d_params) cudaFree(ctx->d_params);
    if (ctx->h_params) cudaFreeHost(ctx->h_params);
    for (auto e : ctx->prof_ev) cudaEventDestroy(e);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int32_t dllm_ctx_sync(dllm_ctx *ctx) {
    CTX_CHECK(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DLLM_OK;
}
void *dllm_ctx_stream(dllm_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
const char *dllm_last_error(const dllm_ctx *ctx) { return ctx ? ctx->err : "null context"; }
uint64_t dllm_launch_count(const dllm_ctx *ctx) { return ctx ? ctx->launches : 0; }
int32_t dllm_copy_bytes(const dllm_ctx *ctx, uint64_t *h2d, uint64_t *d2h) {
    if (!ctx) return DLLM_ERR_NULL;
    if (h2d) *h2d = ctx->h2d_bytes;
    if (d2h) *d2h = ctx->d2h_bytes;
    return DLLM_OK;
}

uint64_t dllm_graph_replay_count(const dllm_ctx *ctx) { return ctx ? ctx->graph_replays : 0; }
int32_t dllm_sm_count(const dllm_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

int32_t dllm_selftest_division(dllm_ctx *ctx, uint64_t cases, uint64_t seed, uint64_t *mismatches) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, mismatches, DLLM_ERR_NULL, "null out pointer");
    unsigned long long *d = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(unsigned long long)));
    cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream);
    int32_t rc = k_selftest_division(ctx, cases, seed, d);
    unsigned long long h = 0;
    if (rc == DLLM_OK && cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = DLLM_ERR_CUDA;
    if (rc == DLLM_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = DLLM_ERR_CUDA;
    cudaFree(d);
    *mismatches = h;
    return rc;
}

int32_t dllm_profile_begin(dllm_ctx *ctx) {
    CTX_CHECK(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->prof_on = true;
    ctx->prof_n = 0;
    ctx->prof_flops = ctx->prof_bytes = 0.0;
    return DLLM_OK;
}

int32_t dllm_profile_end(dllm_ctx *ctx, uint64_t *n_launches, double *total_ms, double *total_flops,
                         double *total_bytes) {
    CTX_CHECK(ctx);
    ctx->prof_on = false;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0.0;
    for (size_t i = 0; i < ctx->prof_n; ++i) {
        float t = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        ms += t;
    }
    if (n_launches) *n_launches = ctx->prof_n;
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = ctx->prof_flops;
    if (total_bytes) *total_bytes = ctx->prof_bytes;
    return DLLM_OK;
}

int32_t dllm_malloc(dllm_ctx *ctx, size_t bytes, void **dptr) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, dptr, DLLM_ERR_NULL, "null out pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return DLLM_OK;
}
int32_t dllm_free(dllm_ctx *ctx, void *dptr) {
    CTX_CHECK(ctx);
    if (dptr) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); CUDA_TRY(ctx, cudaFree(dptr)); }
    return DLLM_OK;
}
int32_t dllm_memcpy_h2d(dllm_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    CTX_CHECK(ctx);
    if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += bytes;
    return DLLM_OK;
}
int32_t dllm_memcpy_d2h(dllm_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    CTX_CHECK(ctx);
    if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += bytes;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DLLM_OK;
}
int32_t dllm_host_alloc(size_t bytes, void **hptr) {
    if (!hptr) return DLLM_ERR_NULL;
    return cudaMallocHost(hptr, bytes ? bytes : 1) == cudaSuccess ? DLLM_OK : DLLM_ERR_OOM;
}
int32_t dllm_host_free(void *hptr) {
    if (hptr && cudaFreeHost(hptr) != cudaSuccess) return DLLM_ERR_CUDA;
    return DLLM_OK;
}

}  // extern "C"

// staging helpers ---------------------------------------------------------------------------
static int32_t stage_in(dllm_ctx *ctx, int slot, const void *host, size_t bytes, void **dev) {
    DLLM_TRY(ensure_buf(ctx, ctx->ws[slot], bytes ? bytes : 16));
    *dev = ctx->ws[slot].p;
    if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += bytes;
    return DLLM_OK;
}
static int32_t stage_out_buf(dllm_ctx *ctx, int slot, size_t bytes, void **dev) {
    DLLM_TRY(ensure_buf(ctx, ctx->ws[slot], bytes ? bytes : 16));
    *dev = ctx->ws[slot].p;
    return DLLM_OK;
}
static int32_t copy_out(dllm_ctx *ctx, void *host, const void *dev, size_t bytes) {
    if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += bytes;
    return DLLM_OK;
}
static int32_t sync(dllm_ctx *ctx) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DLLM_OK;
}

extern "C" {

// ==========================================================================================
// quantizer B
// ==========================================================================================
int32_t dllm_quantize_tensor_dev(dllm_ctx *ctx, const float *x_dev, size_t n, uint8_t bits, int32_t packed,
                                 uint8_t *codes_dev, float *params_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, bits >= 1 && bits <= 8, DLLM_ERR_INVALID_PARAMS, "Bits must be between 1 and 8 (got %d)", (int)bits);
    ARG_CHECK(ctx, !packed || pack_width_ok(bits), DLLM_ERR_INVALID_PARAMS, "packed codes need bits in {1,2,4,8}");
    ARG_CHECK(ctx, params_dev && (n == 0 || (x_dev && codes_dev)), DLLM_ERR_NULL, "null device pointer");
    DLLM_TRY(k_minmax(ctx, x_dev, n, bits, params_dev));
    return k_encode_b(ctx, x_dev, n, bits, packed ? bits : 0, params_dev, 0.f, 0.f, codes_dev);
}

int32_t dllm_dequantize_tensor_dev(dllm_ctx *ctx, const uint8_t *codes_dev, size_t n, uint8_t bits, int32_t packed,
                                   const float *params_dev, float *out_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, !packed || pack_width_ok(bits), DLLM_ERR_INVALID_PARAMS, "packed codes need bits in {1,2,4,8}");
    ARG_CHECK(ctx, params_dev && (n == 0 || (codes_dev && out_dev)), DLLM_ERR_NULL, "null device pointer");
    return k_decode_ab(ctx, codes_dev, n, packed ? bits : 0, params_dev, 0.f, 0.f, out_dev);
}

int32_t dllm_quantize_tensor(dllm_ctx *ctx, const float *x, size_t n, uint8_t bits, uint8_t *codes, float *scale,
                             float *zero_point) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, bits >= 1 && bits <= 8, DLLM_ERR_INVALID_PARAMS, "Bits must be between 1 and 8 (got %d)", (int)bits);
    ARG_CHECK(ctx, scale && zero_point && (n == 0 || (x && codes)), DLLM_ERR_NULL, "null pointer");
    void *dx, *dc;
    DLLM_TRY(stage_in(ctx, 0, x, n * sizeof(float), &dx));
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    DLLM_TRY(dllm_quantize_tensor_dev(ctx, (const float *)dx, n, bits, 0, (uint8_t *)dc, ctx->d_params));
    DLLM_TRY(copy_out(ctx, codes, dc, n));
    DLLM_TRY(copy_out(ctx, ctx->h_params, ctx->d_params, 4 * sizeof(float)));
    DLLM_TRY(sync(ctx));
    *scale = ctx->h_params[0];
    *zero_point = ctx->h_params[1];
    return DLLM_OK;
}

int32_t dllm_dequantize_tensor(dllm_ctx *ctx, const uint8_t *codes, size_t n, float scale, float zero_point,
                               float *out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, n == 0 || (codes && out), DLLM_ERR_NULL, "null pointer");
    void *dc, *dout;
    DLLM_TRY(stage_in(ctx, 1, codes, n, &dc));
    DLLM_TRY(stage_out_buf(ctx, 0, n * sizeof(float), &dout));
    DLLM_TRY(k_decode_ab(ctx, (const uint8_t *)dc, n, 0, nullptr, scale, zero_point, (float *)dout));
    DLLM_TRY(copy_out(ctx, out, dout, n * sizeof(float)));
    return sync(ctx);
}

int32_t dllm_quantize_codes(dllm_ctx *ctx, const float *x, size_t n, uint8_t bits, float scale, float zero_point,
                            uint8_t *codes) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, bits >= 1 && bits <= 8, DLLM_ERR_INVALID_PARAMS, "Bits must be between 1 and 8 (got %d)", (int)bits);
    ARG_CHECK(ctx, n == 0 || (x && codes), DLLM_ERR_NULL, "null pointer");
    void *dx, *dc;
    DLLM_TRY(stage_in(ctx, 0, x, n * sizeof(float), &dx));
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    DLLM_TRY(k_encode_b(ctx, (const float *)dx, n, bits, 0, nullptr, scale, zero_point, (uint8_t *)dc));
    DLLM_TRY(copy_out(ctx, codes, dc, n));
    return sync(ctx);
}

float dllm_compression_ratio(size_t numel, size_t data_len, uint8_t bits) {
    const size_t original = numel * 4;                         // quantization.rs:121
    const size_t compressed = (data_len * (size_t)bits + 7) / 8;  // :122
    return (float)original / (float)compressed;
}

// ==========================================================================================
// quantizer A
// ==========================================================================================
static bool a_range(int32_t qtype, float *lo, float *hi) {
    switch (qtype) {   // quantization/src/quantize.rs:139-144
        case DLLM_QT_INT8: *lo = -128.f; *hi = 127.f; return true;
        case DLLM_QT_INT4: *lo = -8.f; *hi = 7.f; return true;
        case DLLM_QT_BINARY: *lo = 0.f; *hi = 1.f; return true;
        case DLLM_QT_FLOAT8: *lo = -127.f; *hi = 127.f; return true;
    }
    return false;
}

int32_t dllm_quantize_a(dllm_ctx *ctx, const float *x, size_t n, int32_t qtype, float scale, int32_t zero_point,
                        uint8_t *codes) {
    CTX_CHECK(ctx);
    float lo, hi;
    ARG_CHECK(ctx, a_range(qtype, &lo, &hi), DLLM_ERR_INVALID_PARAMS, "unknown QuantizationType %d", qtype);
    ARG_CHECK(ctx, n == 0 || (x && codes), DLLM_ERR_NULL, "null pointer");
    void *dx, *dc;
    DLLM_TRY(stage_in(ctx, 0, x, n * sizeof(float), &dx));
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    DLLM_TRY(k_encode_a(ctx, (const float *)dx, n, scale, (float)zero_point, lo, hi, (uint8_t *)dc));
    DLLM_TRY(copy_out(ctx, codes, dc, n));
    return sync(ctx);
}

int32_t dllm_dequantize_a(dllm_ctx *ctx, const uint8_t *codes, size_t n, float scale, int32_t zero_point,
                          float *out) {
    return dllm_dequantize_tensor(ctx, codes, n, scale, (float)zero_point, out);  // same form, quantize.rs:179-181
}

int32_t dllm_minmax(dllm_ctx *ctx, const float *x, size_t n, float *min_out, float *max_out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, min_out && max_out && (n == 0 || x), DLLM_ERR_NULL, "null pointer");
    void *dx;
    DLLM_TRY(stage_in(ctx, 0, x, n * sizeof(float), &dx));
    DLLM_TRY(k_minmax(ctx, (const float *)dx, n, 0, ctx->d_params));
    DLLM_TRY(copy_out(ctx, ctx->h_params, ctx->d_params, 4 * sizeof(float)));
    DLLM_TRY(sync(ctx));
    *min_out = ctx->h_params[2];
    *max_out = ctx->h_params[3];
    return DLLM_OK;
}

static int32_t f32_as_i32_sat(float v) {
    if (v != v) return 0;
    if (v <= -2147483648.0f) return INT32_MIN;
    if (v >= 2147483648.0f) return INT32_MAX;
    return (int32_t)v;
}

int32_t dllm_calibrate_params(float min, float max, size_t total_samples, uint8_t bits, int32_t symmetric,
                              float *scale, int32_t *zero_point) {
    if (!scale || !zero_point) return DLLM_ERR_NULL;
    if (total_samples == 0) return DLLM_ERR_CALIBRATION_REQUIRED;        // calibrate.rs:73-75
    if (bits > 31) return DLLM_ERR_INVALID_PARAMS;
    volatile float num_levels = (float)(1u << bits);                     // :77
    volatile float range = max - min;                                    // :78
    if (range <= 1.1920929e-07f) { *scale = 1.0f; *zero_point = 0; return DLLM_OK; }
    if (symmetric) {
        volatile float max_abs = fmaxf(fabsf(max), fabsf(min));          // :91
        volatile float t = max_abs * 2.0f;
        *scale = t / (num_levels - 1.0f);                                // :92
        volatile float h = num_levels / 2.0f;
        *zero_point = f32_as_i32_sat(h - 1.0f);                          // :98
    } else {
        volatile float s = range / (num_levels - 1.0f);                  // :94
        *scale = s;
        volatile float q = -min / s;
        *zero_point = f32_as_i32_sat(roundf(q));                         // :100
    }
    return DLLM_OK;
}

// ==========================================================================================
// quantizers C and D
// ==========================================================================================
float dllm_bitquantizer_scale(uint8_t bits) {
    if (bits > 30) return 0.f;
    return 1.0f / (float)(int32_t)((1u << bits) - 1u);   // prefill-kvquant-rs/lib.rs:105
}

int32_t dllm_quantize_c(dllm_ctx *ctx, const float *x, size_t n, uint8_t bits, float scale, float zero_point,
                        uint8_t *codes) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, bits <= 30, DLLM_ERR_INVALID_PARAMS, "bits %d overflows `1 << bits`", (int)bits);
    ARG_CHECK(ctx, n == 0 || (x && codes), DLLM_ERR_NULL, "null pointer");
    void *dx, *dc;
    DLLM_TRY(stage_in(ctx, 0, x, n * sizeof(float), &dx));
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    DLLM_TRY(k_encode_cd(ctx, (const float *)dx, n, bits, 0, scale, zero_point, (uint8_t *)dc));
    DLLM_TRY(copy_out(ctx, codes, dc, n));
    return sync(ctx);
}

int32_t dllm_dequantize_cd(dllm_ctx *ctx, const uint8_t *codes, size_t n, float scale, float zero_point, float *out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, n == 0 || (codes && out), DLLM_ERR_NULL, "null pointer");
    void *dc, *dout;
    DLLM_TRY(stage_in(ctx, 1, codes, n, &dc));
    DLLM_TRY(stage_out_buf(ctx, 0, n * sizeof(float), &dout));
    DLLM_TRY(k_decode_cd(ctx, (const uint8_t *)dc, n, 0, scale, zero_point, nullptr, nullptr, 0, (float *)dout));
    DLLM_TRY(copy_out(ctx, out, dout, n * sizeof(float)));
    return sync(ctx);
}

int32_t dllm_kvquant_quantize_vectors(dllm_ctx *ctx, const float *embeddings, size_t nvec, size_t elems_per_vec,
                                      const uint8_t *cfg_bits, size_t ncfg, const uint8_t *bits, size_t nbits,
                                      uint8_t *codes) {
    CTX_CHECK(ctx);
    if (nbits == 0 || nvec == 0) return DLLM_OK;   // zip with an empty cycle yields nothing (lib.rs:132)
    ARG_CHECK(ctx, embeddings && cfg_bits && bits && codes, DLLM_ERR_NULL, "null pointer");
    // validate first: the reference panics on the first out-of-range vector (lib.rs:133)
    for (size_t v = 0; v < nvec; ++v) {
        const size_t qi = (size_t)bits[v % nbits] / 2;
        ARG_CHECK(ctx, qi < ncfg, DLLM_ERR_INDEX, "quantizers[%zu] out of bounds (len %zu)", qi, ncfg);
        ARG_CHECK(ctx, bits[v % nbits] <= 30 && cfg_bits[qi] <= 30, DLLM_ERR_INVALID_PARAMS, "bits overflow");
    }
    const size_t n = nvec * elems_per_vec;
    void *dx, *dc;
    DLLM_TRY(stage_in(ctx, 0, embeddings, n * sizeof(float), &dx));
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    if (nbits == 1) {   // one bit width: a single launch over all vectors
        const float s = dllm_bitquantizer_scale(cfg_bits[bits[0] / 2]);
        DLLM_TRY(k_encode_cd(ctx, (const float *)dx, n, bits[0], 0, s, 0.0f, (uint8_t *)dc));
    } else {
        for (size_t v = 0; v < nvec; ++v) {
            const uint8_t b = bits[v % nbits];
            const float s = dllm_bitquantizer_scale(cfg_bits[b / 2]);
            DLLM_TRY(k_encode_cd(ctx, (const float *)dx + v * elems_per_vec, elems_per_vec, b, 0, s, 0.0f,
                                 (uint8_t *)dc + v * elems_per_vec));
        }
    }
    DLLM_TRY(copy_out(ctx, codes, dc, n));
    return sync(ctx);
}

int32_t dllm_quantize_d_rows_dev(dllm_ctx *ctx, const float *x_dev, size_t rows, size_t dim, uint8_t bits,
                                 int32_t packed, uint8_t *codes_dev, float *scales_dev, float *zero_points_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, bits <= 31, DLLM_ERR_INVALID_PARAMS, "bits %d overflows `1u32 << bits`", (int)bits);
    ARG_CHECK(ctx, !packed || pack_width_ok(bits), DLLM_ERR_INVALID_PARAMS, "packed codes need bits in {1,2,4,8}");
    ARG_CHECK(ctx, rows * dim == 0 || (x_dev && codes_dev && scales_dev && zero_points_dev), DLLM_ERR_NULL, "null device pointer");
    return k_quant_d_rows(ctx, x_dev, rows, dim, nullptr, 1, bits, packed ? bits : 0, codes_dev, scales_dev,
                          zero_points_dev);
}

int32_t dllm_dequantize_d_rows_dev(dllm_ctx *ctx, const uint8_t *codes_dev, size_t rows, size_t dim, uint8_t bits,
                                   int32_t packed, const float *scales_dev, const float *zero_points_dev,
                                   float *out_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, !packed || pack_width_ok(bits), DLLM_ERR_INVALID_PARAMS, "packed codes need bits in {1,2,4,8}");
    ARG_CHECK(ctx, rows * dim == 0 || (codes_dev && scales_dev && zero_points_dev && out_dev), DLLM_ERR_NULL, "null device pointer");
    return k_decode_cd(ctx, codes_dev, rows * dim, packed ? bits : 0, 0.f, 0.f, scales_dev, zero_points_dev, dim, out_dev);
}

int32_t dllm_quantize_d_rows(dllm_ctx *ctx, const float *x, size_t rows, size_t dim, const uint8_t *bits,
                             size_t nbits, uint8_t *codes, float *scales, float *zero_points) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, nbits > 0 && bits, DLLM_ERR_INVALID_PARAMS, "empty bits list (`i %% 0` panics, fusion_ann.rs:58)");
    for (size_t i = 0; i < nbits; ++i)
        ARG_CHECK(ctx, bits[i] <= 31, DLLM_ERR_INVALID_PARAMS, "bits %d overflows `1u32 << bits`", (int)bits[i]);
    const size_t n = rows * dim;
    ARG_CHECK(ctx, n == 0 || (x && codes && scales && zero_points), DLLM_ERR_NULL, "null pointer");
    if (rows == 0) return DLLM_OK;
    void *dx, *dc, *dsz, *dtab;
    DLLM_TRY(stage_in(ctx, 0, x, n * sizeof(float), &dx));
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    DLLM_TRY(stage_out_buf(ctx, 2, 2 * rows * sizeof(float), &dsz));
    DLLM_TRY(stage_in(ctx, 3, bits, nbits, &dtab));
    float *ds = (float *)dsz, *dz = ds + rows;
    DLLM_TRY(k_quant_d_rows(ctx, (const float *)dx, rows, dim, nbits > 1 ? (const uint8_t *)dtab : nullptr,
                            (int)nbits, bits[0], 0, (uint8_t *)dc, ds, dz));
    DLLM_TRY(copy_out(ctx, codes, dc, n));
    DLLM_TRY(copy_out(ctx, scales, ds, rows * sizeof(float)));
    DLLM_TRY(copy_out(ctx, zero_points, dz, rows * sizeof(float)));
    return sync(ctx);
}

int32_t dllm_dequantize_d_rows(dllm_ctx *ctx, const uint8_t *codes, size_t rows, size_t dim, const float *scales,
                               const float *zero_points, float *out) {
    CTX_CHECK(ctx);
    const size_t n = rows * dim;
    ARG_CHECK(ctx, n == 0 || (codes && scales && zero_points && out), DLLM_ERR_NULL, "null pointer");
    if (n == 0) return DLLM_OK;
    void *dc, *dout, *dsz;
    DLLM_TRY(stage_in(ctx, 1, codes, n, &dc));
    DLLM_TRY(stage_out_buf(ctx, 2, 2 * rows * sizeof(float), &dsz));
    float *ds = (float *)dsz, *dz = ds + rows;
    CUDA_TRY(ctx, cudaMemcpyAsync(ds, scales, rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dz, zero_points, rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    DLLM_TRY(stage_out_buf(ctx, 0, n * sizeof(float), &dout));
    DLLM_TRY(k_decode_cd(ctx, (const uint8_t *)dc, n, 0, 0.f, 0.f, ds, dz, dim, (float *)dout));
    DLLM_TRY(copy_out(ctx, out, dout, n * sizeof(float)));
    return sync(ctx);
}

// ==========================================================================================
// pack / unpack
// ==========================================================================================
size_t dllm_packed_len(size_t n, uint8_t bits) { return (n * (size_t)bits + 7) / 8; }

int32_t dllm_pack_dev(dllm_ctx *ctx, const uint8_t *codes_dev, size_t n, uint8_t bits, uint8_t *packed_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, pack_width_ok(bits), DLLM_ERR_INVALID_PARAMS, "pack width %d not in {1,2,4,8}", (int)bits);
    return k_pack(ctx, codes_dev, n, bits, packed_dev);
}
int32_t dllm_unpack_dev(dllm_ctx *ctx, const uint8_t *packed_dev, size_t n, uint8_t bits, uint8_t *codes_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, pack_width_ok(bits), DLLM_ERR_INVALID_PARAMS, "pack width %d not in {1,2,4,8}", (int)bits);
    return k_unpack(ctx, packed_dev, n, bits, codes_dev);
}
int32_t dllm_pack(dllm_ctx *ctx, const uint8_t *codes, size_t n, uint8_t bits, uint8_t *packed) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, pack_width_ok(bits), DLLM_ERR_INVALID_PARAMS, "pack width %d not in {1,2,4,8}", (int)bits);
    ARG_CHECK(ctx, n == 0 || (codes && packed), DLLM_ERR_NULL, "null pointer");
    void *dc, *dp;
    DLLM_TRY(stage_in(ctx, 1, codes, n, &dc));
    DLLM_TRY(stage_out_buf(ctx, 2, dllm_packed_len(n, bits), &dp));
    DLLM_TRY(k_pack(ctx, (const uint8_t *)dc, n, bits, (uint8_t *)dp));
    DLLM_TRY(copy_out(ctx, packed, dp, dllm_packed_len(n, bits)));
    return sync(ctx);
}
int32_t dllm_unpack(dllm_ctx *ctx, const uint8_t *packed, size_t n, uint8_t bits, uint8_t *codes) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, pack_width_ok(bits), DLLM_ERR_INVALID_PARAMS, "pack width %d not in {1,2,4,8}", (int)bits);
    ARG_CHECK(ctx, n == 0 || (codes && packed), DLLM_ERR_NULL, "null pointer");
    void *dc, *dp;
    DLLM_TRY(stage_in(ctx, 2, packed, dllm_packed_len(n, bits), &dp));
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    DLLM_TRY(k_unpack(ctx, (const uint8_t *)dp, n, bits, (uint8_t *)dc));
    DLLM_TRY(copy_out(ctx, codes, dc, n));
    return sync(ctx);
}

// ==========================================================================================
// quantized weight
// ==========================================================================================
static int32_t qweight_alloc(dllm_ctx *ctx, size_t K, size_t N, uint8_t bits, size_t group, dllm_qweight **out) {
    ARG_CHECK(ctx, out, DLLM_ERR_NULL, "null out pointer");
    *out = nullptr;
    ARG_CHECK(ctx, bits >= 1 && bits <= 8, DLLM_ERR_INVALID_PARAMS, "Bits must be between 1 and 8 (got %d)", (int)bits);
    ARG_CHECK(ctx, K > 0 && N > 0, DLLM_ERR_SHAPE, "empty weight [%zu, %zu]", K, N);
    ARG_CHECK(ctx, group == 0 || (group % WL_TILE_K == 0 && K % group == 0), DLLM_ERR_SHAPE,
              "group %zu must be a multiple of %d that divides K=%zu", group, WL_TILE_K, K);
    dllm_qweight *w = new (std::nothrow) dllm_qweight();
    if (!w) return DLLM_ERR_OOM;
    w->K = K; w->N = N; w->bits = bits; w->device = ctx->device;
    w->k_blocks = (K + WL_TILE_K - 1) / WL_TILE_K;
    w->n_tiles = (N + WL_TILE_N - 1) / WL_TILE_N;
    w->per_tensor = group == 0;
    w->group = group == 0 ? w->k_blocks * WL_TILE_K : group;
    const int cb = wl_container_bits(bits);
    w->tile_bytes = wl_tile_bytes(cb);
    const size_t G = w->per_tensor ? 1 : K / group;
    const size_t Npad = w->n_tiles * WL_TILE_N;
    cudaError_t e = cudaMalloc(&w->d_packed, w->n_tiles * w->k_blocks * w->tile_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_scales, G * Npad * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&w->d_zps, G * Npad * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&w->d_dqparams, G * Npad * sizeof(uint2));
    if (e == cudaSuccess) e = cudaMalloc(&w->d_gparams, G * Npad * sizeof(uint2));
    if (e != cudaSuccess) {
        cudaGetLastError();
        dllm_qweight_destroy(w);
        DLLM_FAIL(ctx, DLLM_ERR_OOM, "cudaMalloc failed for a [%zu,%zu] %d-bit weight", K, N, (int)bits);
    }
    *out = w;
    return DLLM_OK;
}

static int32_t qweight_set_bias(dllm_ctx *ctx, dllm_qweight *w, const float *bias, bool bias_on_device) {
    if (!bias) return DLLM_OK;
    CUDA_TRY(ctx, cudaMalloc(&w->d_bias, w->N * sizeof(float)));
    CUDA_TRY(ctx, cudaMemcpyAsync(w->d_bias, bias, w->N * sizeof(float),
                                  bias_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    return DLLM_OK;
}

int32_t dllm_qweight_quantize_dev(dllm_ctx *ctx, const float *w_dev, size_t K, size_t N, uint8_t bits, size_t group,
                                  const float *bias_dev, dllm_qweight **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w_dev, DLLM_ERR_NULL, "null weight pointer");
    dllm_qweight *w = nullptr;
    DLLM_TRY(qweight_alloc(ctx, K, N, bits, group, &w));
    int32_t rc;
    if (w->per_tensor) {
        rc = k_minmax(ctx, w_dev, K * N, bits, ctx->d_params);
        if (rc == DLLM_OK) rc = k_wparams_broadcast(ctx, ctx->d_params, N, w->d_scales, w->d_zps);
        if (rc == DLLM_OK) {
            cudaMemcpyAsync(ctx->h_params, ctx->d_params, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            w->tensor_scale = ctx->h_params[0];
            w->tensor_zp = ctx->h_params[1];
        }
    } else {
        rc = k_wparams_grouped(ctx, w_dev, K, N, group, bits, w->d_scales, w->d_zps);
    }
    if (rc == DLLM_OK) rc = k_wpack_from_f32(ctx, w_dev, w);
    if (rc == DLLM_OK) rc = k_wdq_params(ctx, w);
    if (rc == DLLM_OK) rc = qweight_set_bias(ctx, w, bias_dev, true);
    if (rc != DLLM_OK) { dllm_qweight_destroy(w); return rc; }
    *out = w;
    return DLLM_OK;
}

int32_t dllm_qweight_quantize(dllm_ctx *ctx, const float *w_host, size_t K, size_t N, uint8_t bits, size_t group,
                              const float *bias, dllm_qweight **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w_host, DLLM_ERR_NULL, "null weight pointer");
    ARG_CHECK(ctx, K > 0 && N > 0, DLLM_ERR_SHAPE, "empty weight [%zu, %zu]", K, N);
    void *dw, *db = nullptr;
    DLLM_TRY(stage_in(ctx, 0, w_host, K * N * sizeof(float), &dw));
    if (bias) DLLM_TRY(stage_in(ctx, 3, bias, N * sizeof(float), &db));
    DLLM_TRY(dllm_qweight_quantize_dev(ctx, (const float *)dw, K, N, bits, group, (const float *)db, out));
    return sync(ctx);
}

// adopt canonical codes that are already on the device (one per u8, [K,N]); scales / zero-points / bias are host arrays
static int32_t qweight_from_codes_dev(dllm_ctx *ctx, const uint8_t *codes_dev, const float *scales, const float *zero_points,
                                      size_t K, size_t N, uint8_t bits, size_t group, const float *bias, dllm_qweight **out) {
    dllm_qweight *w = nullptr;
    DLLM_TRY(qweight_alloc(ctx, K, N, bits, group, &w));
    const size_t Npad = w->n_tiles * WL_TILE_N;
    int32_t rc = DLLM_OK;
    do {
        if (w->per_tensor) {
            ctx->h_params[0] = scales[0];
            ctx->h_params[1] = zero_points[0];
            w->tensor_scale = scales[0];
            w->tensor_zp = zero_points[0];
            cudaMemcpyAsync(ctx->d_params, ctx->h_params, 2 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
            if ((rc = k_wparams_broadcast(ctx, ctx->d_params, N, w->d_scales, w->d_zps)) != DLLM_OK) break;
        } else {
            const size_t G = K / group;
            cudaMemsetAsync(w->d_scales, 0, G * Npad * sizeof(float), ctx->stream);
            cudaMemsetAsync(w->d_zps, 0, G * Npad * sizeof(float), ctx->stream);
            cudaMemcpy2DAsync(w->d_scales, Npad * sizeof(float), scales, N * sizeof(float), N * sizeof(float), G,
                              cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpy2DAsync(w->d_zps, Npad * sizeof(float), zero_points, N * sizeof(float), N * sizeof(float), G,
                              cudaMemcpyHostToDevice, ctx->stream);
        }
        if ((rc = k_wpack_from_codes(ctx, codes_dev, w)) != DLLM_OK) break;
        if ((rc = k_wdq_params(ctx, w)) != DLLM_OK) break;
        if ((rc = qweight_set_bias(ctx, w, bias, false)) != DLLM_OK) break;
        rc = sync(ctx);
    } while (0);
    if (rc != DLLM_OK) { dllm_qweight_destroy(w); return rc; }
    *out = w;
    return DLLM_OK;
}

int32_t dllm_qweight_from_codes(dllm_ctx *ctx, const uint8_t *codes, const float *scales, const float *zero_points,
                                size_t K, size_t N, uint8_t bits, size_t group, const float *bias,
                                dllm_qweight **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, codes && scales && zero_points && out, DLLM_ERR_NULL, "null pointer");
    ARG_CHECK(ctx, K > 0 && N > 0, DLLM_ERR_SHAPE, "empty weight [%zu, %zu]", K, N);
    void *dc = nullptr;
    DLLM_TRY(stage_in(ctx, 1, codes, K * N, &dc));
    return qweight_from_codes_dev(ctx, (const uint8_t *)dc, scales, zero_points, K, N, bits, group, bias, out);
}

// ---- packed-weights container: the wire / on-disk form of a quantized linear ----
//   header  64 B : "DLLMQW01" | u32 version (1) | u32 bits | u64 K | u64 N | u64 group (0 = one scale / zero-point per
//                  tensor, the reference's scheme) | u32 scheme (0 = quantizer B, `(q - zp) * scale`) | u32 has_bias |
//                  u64 codes_bytes | u64 reserved (0)
//   codes        : the [K,N] row-major codes bit-packed LSB first (dllm_pack's layout) at the narrowest width in {1,2,4,8}
//                  that holds `bits` — QuantizedTensor::data (quantization/src/types.rs:42-47) packed the way its own
//                  accounting assumes ((len * bits + 7) / 8, diffuse-llm-rs/src/quantization.rs:122)
//   scales, zps  : f32 [K/group, N] each (or one value each) — QuantizationParams::{scale, zero_point}
//   bias         : f32 [N] when has_bias
//   crc32   4 B  : IEEE CRC-32 of everything before it
// The tile-major layout the kernels read (wlayout.cuh) is private and never serialised.
namespace {
struct QwHeader {
    char magic[8];
    uint32_t version, bits;
    uint64_t K, N, group;
    uint32_t scheme, has_bias;
    uint64_t codes_bytes, reserved;
};
static_assert(sizeof(QwHeader) == 64, "container header is 64 bytes");

uint32_t crc32_ieee(const uint8_t *p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
int pack_width_for(int bits) { return bits <= 1 ? 1 : (bits <= 2 ? 2 : (bits <= 4 ? 4 : 8)); }
size_t qw_param_count(const dllm_qweight *w) { return w->per_tensor ? 1 : (w->K / w->group) * w->N; }
}  // namespace

size_t dllm_qweight_serialized_size(const dllm_qweight *w) {
    if (!w) return 0;
    const size_t codes = dllm_packed_len(w->K * w->N, (uint8_t)pack_width_for(w->bits));
    return sizeof(QwHeader) + codes + 2 * qw_param_count(w) * sizeof(float) + (w->d_bias ? w->N * sizeof(float) : 0) + 4;
}

int32_t dllm_qweight_serialize(dllm_ctx *ctx, const dllm_qweight *w, uint8_t *buf, size_t cap, size_t *written) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w && buf, DLLM_ERR_NULL, "null pointer");
    const size_t total = dllm_qweight_serialized_size(w);
    ARG_CHECK(ctx, cap >= total, DLLM_ERR_INVALID_PARAMS, "buffer of %zu bytes is too small for %zu", cap, total);
    const int pw = pack_width_for(w->bits);
    const size_t n = w->K * w->N, codes_bytes = dllm_packed_len(n, (uint8_t)pw), np = qw_param_count(w);
    QwHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "DLLMQW01", 8);
    h.version = 1; h.bits = (uint32_t)w->bits; h.K = w->K; h.N = w->N; h.group = w->per_tensor ? 0 : w->group;
    h.scheme = 0; h.has_bias = w->d_bias ? 1u : 0u; h.codes_bytes = codes_bytes;
    memcpy(buf, &h, sizeof(h));
    uint8_t *p = buf + sizeof(h);
    void *dc, *dp;
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    DLLM_TRY(stage_out_buf(ctx, 2, codes_bytes, &dp));
    DLLM_TRY(k_wexport_codes(ctx, w, (uint8_t *)dc));
    DLLM_TRY(k_pack(ctx, (const uint8_t *)dc, n, pw, (uint8_t *)dp));
    DLLM_TRY(copy_out(ctx, p, dp, codes_bytes));
    DLLM_TRY(sync(ctx));
    p += codes_bytes;
    DLLM_TRY(dllm_qweight_export(ctx, w, nullptr, (float *)p, (float *)(p + np * sizeof(float))));
    p += 2 * np * sizeof(float);
    if (w->d_bias) {
        DLLM_TRY(copy_out(ctx, p, w->d_bias, w->N * sizeof(float)));
        DLLM_TRY(sync(ctx));
        p += w->N * sizeof(float);
    }
    const uint32_t crc = crc32_ieee(buf, (size_t)(p - buf));
    memcpy(p, &crc, 4);
    if (written) *written = total;
    return DLLM_OK;
}

int32_t dllm_qweight_deserialize(dllm_ctx *ctx, const uint8_t *buf, size_t len, dllm_qweight **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, buf && out, DLLM_ERR_NULL, "null pointer");
    *out = nullptr;
    ARG_CHECK(ctx, len >= sizeof(QwHeader) + 4, DLLM_ERR_SERIALIZATION, "truncated container (%zu bytes)", len);
    QwHeader h;
    memcpy(&h, buf, sizeof(h));
    ARG_CHECK(ctx, memcmp(h.magic, "DLLMQW01", 8) == 0 && h.version == 1, DLLM_ERR_SERIALIZATION, "not a DLLMQW01 container");
    ARG_CHECK(ctx, h.bits >= 1 && h.bits <= 8 && h.scheme == 0 && h.K > 0 && h.N > 0 && h.K < ((uint64_t)1 << 32) && h.N < ((uint64_t)1 << 32) &&
              (h.group == 0 || h.K % h.group == 0), DLLM_ERR_INVALID_DATA_FORMAT, "inconsistent container header");
    const int pw = pack_width_for((int)h.bits);
    const size_t n = (size_t)h.K * h.N, codes_bytes = dllm_packed_len(n, (uint8_t)pw);
    const size_t np = h.group == 0 ? 1 : (size_t)(h.K / h.group) * h.N;
    const size_t total = sizeof(QwHeader) + codes_bytes + 2 * np * sizeof(float) + (h.has_bias ? h.N * sizeof(float) : 0) + 4;
    ARG_CHECK(ctx, h.codes_bytes == codes_bytes && len == total, DLLM_ERR_INVALID_DATA_FORMAT, "container size %zu does not match its header (%zu)", len, total);
    uint32_t crc;
    memcpy(&crc, buf + total - 4, 4);
    ARG_CHECK(ctx, crc == crc32_ieee(buf, total - 4), DLLM_ERR_SERIALIZATION, "CRC mismatch: the container is corrupt");
    const uint8_t *p = buf + sizeof(QwHeader);
    void *dp, *dc;
    DLLM_TRY(stage_in(ctx, 2, p, codes_bytes, &dp));
    DLLM_TRY(stage_out_buf(ctx, 1, n, &dc));
    DLLM_TRY(k_unpack(ctx, (const uint8_t *)dp, n, pw, (uint8_t *)dc));
    p += codes_bytes;
    // (scales / zps / bias may be unaligned inside the caller's buffer: copy them out)
    std::vector<float> sc(np), zp(np), bias(h.has_bias ? h.N : 0);
    memcpy(sc.data(), p, np * sizeof(float));
    memcpy(zp.data(), p + np * sizeof(float), np * sizeof(float));
    p += 2 * np * sizeof(float);
    if (h.has_bias) memcpy(bias.data(), p, h.N * sizeof(float));
    return qweight_from_codes_dev(ctx, (const uint8_t *)dc, sc.data(), zp.data(), (size_t)h.K, (size_t)h.N, (uint8_t)h.bits,
                                  (size_t)h.group, h.has_bias ? bias.data() : nullptr, out);
}

int32_t dllm_qweight_save(dllm_ctx *ctx, const dllm_qweight *w, const char *path) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w && path, DLLM_ERR_NULL, "null pointer");
    std::vector<uint8_t> buf(dllm_qweight_serialized_size(w));
    size_t written = 0;
    DLLM_TRY(dllm_qweight_serialize(ctx, w, buf.data(), buf.size(), &written));
    FILE *f = fopen(path, "wb");
    if (!f) DLLM_FAIL(ctx, DLLM_ERR_IO, "cannot open %s for writing", path);
    const size_t nw = fwrite(buf.data(), 1, written, f);
    const int rc = fclose(f);
    if (nw != written || rc != 0) DLLM_FAIL(ctx, DLLM_ERR_IO, "short write to %s", path);
    return DLLM_OK;
}

int32_t dllm_qweight_load(dllm_ctx *ctx, const char *path, dllm_qweight **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, path && out, DLLM_ERR_NULL, "null pointer");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) DLLM_FAIL(ctx, DLLM_ERR_IO, "cannot open %s", path);
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); DLLM_FAIL(ctx, DLLM_ERR_IO, "cannot size %s", path); }
    std::vector<uint8_t> buf((size_t)sz);
    const size_t nr = fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    if (nr != buf.size()) DLLM_FAIL(ctx, DLLM_ERR_IO, "short read from %s", path);
    return dllm_qweight_deserialize(ctx, buf.data(), buf.size(), out);
}

int32_t dllm_qweight_export(dllm_ctx *ctx, const dllm_qweight *w, uint8_t *codes, float *scales, float *zero_points) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w, DLLM_ERR_NULL, "null weight");
    const size_t Npad = w->n_tiles * WL_TILE_N;
    if (codes) {
        void *dc;
        DLLM_TRY(stage_out_buf(ctx, 1, w->K * w->N, &dc));
        DLLM_TRY(k_wexport_codes(ctx, w, (uint8_t *)dc));
        DLLM_TRY(copy_out(ctx, codes, dc, w->K * w->N));
    }
    const size_t G = w->per_tensor ? 1 : w->K / w->group;
    if (w->per_tensor) {
        if (scales) CUDA_TRY(ctx, cudaMemcpyAsync(scales, w->d_scales, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        if (zero_points) CUDA_TRY(ctx, cudaMemcpyAsync(zero_points, w->d_zps, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        if (scales) CUDA_TRY(ctx, cudaMemcpy2DAsync(scales, w->N * sizeof(float), w->d_scales, Npad * sizeof(float), w->N * sizeof(float), G, cudaMemcpyDeviceToHost, ctx->stream));
        if (zero_points) CUDA_TRY(ctx, cudaMemcpy2DAsync(zero_points, w->N * sizeof(float), w->d_zps, Npad * sizeof(float), w->N * sizeof(float), G, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return sync(ctx);
}

int32_t dllm_qweight_info(const dllm_qweight *w, size_t *K, size_t *N, uint8_t *bits, size_t *group,
                          size_t *packed_bytes) {
    if (!w) return DLLM_ERR_NULL;
    if (K) *K = w->K;
    if (N) *N = w->N;
    if (bits) *bits = (uint8_t)w->bits;
    if (group) *group = w->per_tensor ? 0 : w->group;
    if (packed_bytes) *packed_bytes = w->n_tiles * w->k_blocks * w->tile_bytes;
    return DLLM_OK;
}

void dllm_qweight_destroy(dllm_qweight *w) {
    if (!w) return;
    cudaSetDevice(w->device);
    if (w->d_packed) cudaFree(w->d_packed);
    if (w->d_scales) cudaFree(w->d_scales);
    if (w->d_zps) cudaFree(w->d_zps);
    if (w->d_dqparams) cudaFree(w->d_dqparams);
    if (w->d_gparams) cudaFree(w->d_gparams);
    if (w->d_bias) cudaFree(w->d_bias);
    delete w;
}

// ==========================================================================================
// quantized linear
// ==========================================================================================
static int resolve_path(const dllm_qweight *w, size_t M, int32_t path) {
    if (path == DLLM_PATH_AUTO) {
        if (k_gemv_supported(w, M)) return DLLM_PATH_GEMV;
        return k_umma_supported(w, M) ? DLLM_PATH_UMMA : DLLM_PATH_SIMT;
    }
    return path;
}

// int8 denoise mode, one linear: per-token symmetric int8 activations (from the bf16 tensor the stack carries), exact integer
// contraction, dequantization + bias in the epilogue
static int32_t i8_linear(dllm_ctx *ctx, const dllm_qweight *w, const void *x_bf16, size_t M, float *y_f32, void *y_bf16) {
    const size_t aux = (M * 4 + 255) & ~(size_t)255;
    DLLM_TRY(ensure_buf(ctx, ctx->act[2], M * w->K));
    DLLM_TRY(ensure_buf(ctx, ctx->lin_flags, 2 * aux));
    int32_t *rowsum = (int32_t *)ctx->lin_flags.p;
    float *rowscale = (float *)((char *)ctx->lin_flags.p + aux);
    DLLM_TRY(k_rowquant_i8(ctx, x_bf16, M, w->K, w->tensor_scale, (int8_t *)ctx->act[2].p, rowscale, rowsum));
    return k_qlinear_umma_i8_deq(ctx, w, (const int8_t *)ctx->act[2].p, rowsum, rowscale, M, y_f32, y_bf16);
}

int32_t dllm_qlinear_forward_dev(dllm_ctx *ctx, const dllm_qweight *w, const float *x_dev, size_t M, float *y_dev,
                                 int32_t path) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w, DLLM_ERR_NULL, "null weight");
    ARG_CHECK(ctx, M == 0 || (x_dev && y_dev), DLLM_ERR_NULL, "null device pointer");
    ARG_CHECK(ctx, path >= DLLM_PATH_AUTO && path <= DLLM_PATH_I8, DLLM_ERR_INVALID_PARAMS, "unknown path %d", path);
    if (M == 0) return DLLM_OK;
    const int p = resolve_path(w, M, path);
    if (p == DLLM_PATH_I8) {
        ARG_CHECK(ctx, k_umma_i8_supported(w, M), DLLM_ERR_UNSUPPORTED,
                  "int8 path needs a per-tensor quantized weight (group_size 0) with K %% 64 == 0 and K <= 65536");
        DLLM_TRY(ensure_buf(ctx, ctx->act[0], M * w->K * 2));
        DLLM_TRY(k_f32_to_bf16(ctx, x_dev, M * w->K, ctx->act[0].p));
        return i8_linear(ctx, w, ctx->act[0].p, M, y_dev, nullptr);
    }
    if (p == DLLM_PATH_SIMT) return k_qlinear_simt(ctx, w, x_dev, M, y_dev);
    if (p == DLLM_PATH_GEMV) {
        ARG_CHECK(ctx, k_gemv_supported(w, M), DLLM_ERR_UNSUPPORTED, "GEMV path needs 1 <= M <= 16 (got %zu)", M);
        return k_qlinear_gemv(ctx, w, x_dev, M, y_dev);
    }
    ARG_CHECK(ctx, k_umma_supported(w, M), DLLM_ERR_UNSUPPORTED, "tcgen05 path does not support this shape");
    DLLM_TRY(ensure_buf(ctx, ctx->act[0], M * w->K * 2));
    DLLM_TRY(k_f32_to_bf16(ctx, x_dev, M * w->K, ctx->act[0].p));
    // timing experiments only (scripts/dense_probe.py): write bf16 into y, as the layers inside a stack do
    static const bool probe_bf16 = getenv("DLLM_PROBE_BF16_OUT") != nullptr;
    if (probe_bf16) return k_qlinear_umma(ctx, w, ctx->act[0].p, M, nullptr, y_dev);
    return k_qlinear_umma(ctx, w, ctx->act[0].p, M, y_dev, nullptr);
}

int32_t dllm_qlinear_forward(dllm_ctx *ctx, const dllm_qweight *w, const float *x, size_t M, float *y, int32_t path) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w, DLLM_ERR_NULL, "null weight");
    ARG_CHECK(ctx, M == 0 || (x && y), DLLM_ERR_NULL, "null pointer");
    if (M == 0) return DLLM_OK;
    void *dx, *dy;
    DLLM_TRY(stage_in(ctx, 4, x, M * w->K * sizeof(float), &dx));
    DLLM_TRY(stage_out_buf(ctx, 5, M * w->N * sizeof(float), &dy));
    DLLM_TRY(dllm_qlinear_forward_dev(ctx, w, (const float *)dx, M, (float *)dy, path));
    DLLM_TRY(copy_out(ctx, y, dy, M * w->N * sizeof(float)));
    return sync(ctx);
}

int32_t dllm_qlinear_forward_i8_dev(dllm_ctx *ctx, const dllm_qweight *w, const int8_t *xq_dev, size_t M, int32_t *y_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w, DLLM_ERR_NULL, "null weight");
    ARG_CHECK(ctx, M == 0 || (xq_dev && y_dev), DLLM_ERR_NULL, "null device pointer");
    if (M == 0) return DLLM_OK;
    ARG_CHECK(ctx, k_umma_i8_supported(w, M), DLLM_ERR_UNSUPPORTED,
              "int8 path needs a per-tensor quantized weight (group_size 0) with K %% 64 == 0 and K <= 65536");
    return k_qlinear_umma_i8(ctx, w, xq_dev, M, y_dev);
}

int32_t dllm_qlinear_forward_i8(dllm_ctx *ctx, const dllm_qweight *w, const int8_t *xq, size_t M, int32_t *y) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, w, DLLM_ERR_NULL, "null weight");
    ARG_CHECK(ctx, M == 0 || (xq && y), DLLM_ERR_NULL, "null pointer");
    if (M == 0) return DLLM_OK;
    void *dx, *dy;
    DLLM_TRY(stage_in(ctx, 4, xq, M * w->K, &dx));
    DLLM_TRY(stage_out_buf(ctx, 5, M * w->N * sizeof(int32_t), &dy));
    DLLM_TRY(dllm_qlinear_forward_i8_dev(ctx, w, (const int8_t *)dx, M, (int32_t *)dy));
    DLLM_TRY(copy_out(ctx, y, dy, M * w->N * sizeof(int32_t)));
    return sync(ctx);
}

int32_t dllm_dequant_matmul(dllm_ctx *ctx, const uint8_t *codes, const float *scales, const float *zero_points,
                            size_t K, size_t N, uint8_t bits, size_t group, const float *bias, const float *x,
                            size_t M, float *y, int32_t path) {
    dllm_qweight *w = nullptr;
    DLLM_TRY(dllm_qweight_from_codes(ctx, codes, scales, zero_points, K, N, bits, group, bias, &w));
    int32_t rc = dllm_qlinear_forward(ctx, w, x, M, y, path);
    dllm_qweight_destroy(w);
    return rc;
}

}  // extern "C"

// ==========================================================================================
// model: stack of quantized linears + schedule (diffuse-llm-rs/src/lib.rs:748-813, 554-593)
// ==========================================================================================
struct dllm_model {
    size_t hidden = 0;
    std::vector<dllm_qweight *> layers;
    std::vector<int> parallel;   // 0 replicated, 1 column-parallel, 2 row-parallel
    size_t T = 0;
    std::vector<float> betas, alpha_bars;
    float *d_coef_table[2] = {nullptr, nullptr};   // [T][4] for guard_t0 = 0 / 1 (c1, c2, std, degenerate)
    float *d_noise_table = nullptr;                // [T][2] {sqrt(alpha_bar_t), sqrt(1 - alpha_bar_t)} (add_noise)
    // seeded loop: {int t; int pad; u64 seed} on the device, and one captured denoise step (CUDA graph) that reads it
    void *d_state = nullptr;
    cudaGraphExec_t step_graph = nullptr;
    float *graph_x = nullptr;
    size_t graph_batch = 0, graph_feat = 0, graph_launches = 0;
    int graph_path = -1, graph_guard = -1;
    int *d_rowmap = nullptr;     // per-row timestep for the host-pointer p_sample
    std::vector<int> h_rowmap;   // its host staging
    size_t rowmap_cap = 0;
    int device = 0;
};

static int32_t beta_schedule_host(int32_t kind, size_t T, float beta_start, float beta_end, float *betas) {
    if (T == 0 || !betas) return DLLM_ERR_INVALID_PARAMS;
    const float PI_F = 3.14159265358979323846f;
    for (size_t t = 0; t < T; ++t) {
        // volatile temporaries: one f32 rounding per reference operation, whatever the host compiler flags
        switch (kind) {
            case DLLM_BETA_LINEAR: {   // lib.rs:560-563
                volatile float d = beta_end - beta_start;
                volatile float p = d * (float)t;
                volatile float q = p / (float)(T - 1);
                betas[t] = beta_start + q;
                break;
            }
            case DLLM_BETA_QUADRATIC: {   // lib.rs:569-573
                volatile float tn = (float)t / (float)(T - 1);
                volatile float d = beta_end - beta_start;
                volatile float p = d * tn;
                volatile float q = p * tn;
                betas[t] = beta_start + q;
                break;
            }
            case DLLM_BETA_COSINE: {   // lib.rs:578-587
                const float s = 0.008f;
                volatile float tn = (float)t / (float)T;
                volatile float a = tn + s;
                volatile float b = a / (1.0f + s);
                volatile float c = b * PI_F;
                volatile float d = c / 2.0f;
                volatile float ct = cosf(d);
                volatile float ft = ct * ct;
                volatile float a0 = s / (1.0f + s);
                volatile float c0 = a0 * PI_F;
                volatile float d0 = c0 / 2.0f;
                volatile float cz = cosf(d0);
                volatile float f0 = cz * cz;
                volatile float r = ft / f0;
                betas[t] = fminf(1.0f - r, 0.999f);
                break;
            }
            default: return DLLM_ERR_INVALID_PARAMS;
        }
    }
    return DLLM_OK;
}

// coefficients of p_sample for timestep t: lib.rs:1160-1192, 1208-1209 (alphas read as alpha_t)
static void p_sample_coeffs_host(const dllm_model *m, size_t t, int guard_t0, float out[4]) {
    const size_t ti = t < m->T - 1 ? t : m->T - 1;
    const float ab_t = m->alpha_bars[ti];
    const float beta_t = m->betas[ti];
    volatile float alpha_t = 1.0f - beta_t;
    const float ab_prev = ti > 0 ? m->alpha_bars[ti - 1] : 1.0f;
    volatile float one_m_abt = 1.0f - ab_t;
    volatile float one_m_abp = 1.0f - ab_prev;
    volatile float n1 = sqrtf(ab_prev) * beta_t;
    volatile float n2 = sqrtf(alpha_t) * one_m_abp;
    volatile float c1 = n1 / one_m_abt, c2 = n2 / one_m_abt;
    volatile float ratio = one_m_abp / one_m_abt;
    volatile float var = ratio * beta_t;
    out[0] = c1; out[1] = c2; out[2] = sqrtf(var);
    out[3] = (guard_t0 && one_m_abt == 0.0f) ? 1.0f : 0.0f;
}

// tensor-parallel hooks (tp.cu)
int32_t tp_allreduce(dllm_ctx *ctx, float *buf, size_t n);
int32_t tp_allreduce_bf16(dllm_ctx *ctx, void *buf, size_t n);
int32_t tp_allreduce_on(dllm_ctx *ctx, void *buf, size_t n, bool bf16, cudaStream_t stream);
int32_t tp_allreduce_minmax(dllm_ctx *ctx, float *params_dev);
int32_t tp_allgather_cols(dllm_ctx *ctx, const float *in, size_t M, size_t n_local, float *out);
bool tp_p2p_regions(const dllm_ctx *ctx, size_t bytes_each, char **b0, char **b1, char **recv);
void tp_p2p_peer_ptrs(const dllm_ctx *ctx, const void *local, void **out8);
int32_t tp_reduce_gather(dllm_ctx *ctx, const void *recv, void *dst, size_t rows, size_t N, cudaStream_t stream, int signal);

// The two ping-pong activation buffers of a tensor-parallel tcgen05 stack: the halves of the peer-to-peer arena when it is
// enabled and large enough (dllm_tp_p2p_enable: the row-parallel partial sums are then reduced in place by this library's
// own NVLink kernel), else the context's ordinary buffers (ncclAllReduce).  Same sizes on every rank => same offsets.
static int32_t tp_act_bufs(dllm_ctx *ctx, size_t bytes_each, char **b0, char **b1, char **recv) {
    *recv = nullptr;
    if (tp_p2p_regions(ctx, bytes_each, b0, b1, recv)) return DLLM_OK;
    DLLM_TRY(ensure_buf(ctx, ctx->act[0], bytes_each));
    DLLM_TRY(ensure_buf(ctx, ctx->act[1], bytes_each));
    *b0 = (char *)ctx->act[0].p;
    *b1 = (char *)ctx->act[1].p;
    return DLLM_OK;
}

// A row-parallel layer that is not the stack's last one, with the reduce-scatter fused into the dense kernel's epilogue and the
// all-gather done by tp_reduce_gather: possible when the arena has a receive region and the shape suits the CTA-pair kernel.
// (DLLM_TP_FUSED_RS=0 switches it off: experiments.)
static bool tp_fused_rs_ok(const dllm_ctx *ctx, const dllm_qweight *w, size_t tokens, const char *recv) {
    const char *sw = getenv("DLLM_TP_FUSED_RS");                  // read per call: the multi-GPU check toggles it
    const bool off = sw && atoi(sw) == 0;
    return !off && recv && !ctx->tp_skip_comm && k_umma_rs_supported(ctx, w, tokens, ctx->tp_world);
}

// x [tokens, K shard] -> recv buffers of all ranks (fused epilogue); then, on `stream`, recv -> dst [tokens, N] on every rank
// gate_next: the reduce / gather kernel ends with per-source signals instead of a barrier and arms the NEXT dense kernel to gate its
// activation loads on them (serial placement only: the consumer must follow on the same stream)
static int32_t tp_row_linear_fused(dllm_ctx *ctx, const dllm_qweight *w, const void *x_bf16, size_t tokens, char *recv, void *dst,
                                   cudaStream_t reduce_stream, cudaEvent_t gemm_done, int gate_next = 0) {
    UmmaRs rs;
    tp_p2p_peer_ptrs(ctx, recv, rs.recv);
    rs.world = ctx->tp_world; rs.rank = ctx->tp_rank; rs.rows = tokens / (size_t)ctx->tp_world;
    DLLM_TRY(k_qlinear_umma_rs(ctx, w, x_bf16, tokens, &rs));
    if (reduce_stream != ctx->stream) {
        CUDA_TRY(ctx, cudaEventRecord(gemm_done, ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(reduce_stream, gemm_done, 0));
    }
    return tp_reduce_gather(ctx, recv, dst, rs.rows, w->N, reduce_stream, gate_next);
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// The tcgen05 stack under tensor parallelism with the collectives overlapped (one exchange per column->row
// pair, at the layer boundary).  The tokens are cut into `chunks` pieces; a segment of layers ending in a row-parallel one
// is run chunk by chunk, and each chunk's all-reduce (in place, on the bf16 tensor the next linear reads) goes to the
// communication stream as soon as its GEMM is done — it runs under the NEXT chunk's GEMMs, which leave `sm_reserve` SMs
// free for the collective's CTAs.  The first GEMM that reads a chunk waits for that chunk's all-reduce only.
static int32_t forward_tp_overlapped(dllm_ctx *ctx, dllm_model *m, const float *x_dev, size_t tokens, float *out_dev, int chunks) {
    const size_t L = m->layers.size();
    size_t maxw = m->hidden;
    for (auto *w : m->layers) { if (w->N > maxw) maxw = w->N; if (w->K > maxw) maxw = w->K; }
    if (!ctx->comm_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    while (ctx->tp_ev.size() < (size_t)(2 * chunks)) {
        cudaEvent_t e;
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->tp_ev.push_back(e);
    }
    // chunk boundaries on multiples of 256 tokens (the dense kernel's widest tile)
    std::vector<size_t> c0(chunks + 1, tokens);
    const size_t per = ((tokens + chunks - 1) / chunks + 255) / 256 * 256;
    for (int c = 0; c <= chunks; ++c) c0[c] = (size_t)c * per < tokens ? (size_t)c * per : tokens;
    // every chunk owns one region of each ping-pong buffer ([its tokens, width] dense inside): chunks never overlap,
    // whatever the layers' widths are
    const size_t region = per * maxw * 2;
    char *buf0, *buf1, *recv0;
    DLLM_TRY(tp_act_bufs(ctx, region * chunks, &buf0, &buf1, &recv0));
    for (int c = 0; c < chunks; ++c)
        if (c0[c + 1] > c0[c])
            DLLM_TRY(k_f32_to_bf16(ctx, x_dev + c0[c] * m->layers[0]->K, (c0[c + 1] - c0[c]) * m->layers[0]->K, buf0 + c * region));
    // SMs left to the collective's CTAs (NCCL's, or 16 blocks of the library's own reduce / gather kernel)
    const int reserve = ctx->sm_reserve >= 0 ? ctx->sm_reserve : env_int("DLLM_TP_RESERVE_SMS", 8);
    ctx->sm_limit = ctx->sm_count - reserve;
    std::vector<char> pending(chunks, 0);          // chunk c's input is still being all-reduced on the comm stream
    char *cur = buf0, *nxt = buf1;
    int32_t rc = DLLM_OK;
    size_t l0 = 0;
    while (l0 < L && rc == DLLM_OK) {
        size_t l1 = l0;                            // segment [l0, l1]: up to and including the next row-parallel layer
        while (l1 + 1 < L && m->parallel[l1] != 2) ++l1;
        const bool reduce = m->parallel[l1] == 2;
        for (int c = 0; c < chunks && rc == DLLM_OK; ++c) {
            const size_t t0 = c0[c], tn = c0[c + 1] - c0[c];
            if (tn == 0) continue;
            if (pending[c]) { rc = cudaStreamWaitEvent(ctx->stream, ctx->tp_ev[2 * c + 1], 0) == cudaSuccess ? DLLM_OK : DLLM_ERR_CUDA; pending[c] = 0; }
            char *a = cur + c * region, *b = nxt + c * region;
            bool fused = false;
            for (size_t l = l0; l <= l1 && rc == DLLM_OK; ++l) {
                const dllm_qweight *w = m->layers[l];
                const bool last = l + 1 == L;
                if (l == l1 && reduce && !last && tp_fused_rs_ok(ctx, w, tn, recv0)) {
                    // reduce-scatter inside the GEMM's epilogue; the all-gather half runs on the communication stream
                    rc = tp_row_linear_fused(ctx, w, a, tn, recv0 + c * region, b, ctx->comm_stream, ctx->tp_ev[2 * c]);
                    if (rc == DLLM_OK && cudaEventRecord(ctx->tp_ev[2 * c + 1], ctx->comm_stream) != cudaSuccess) rc = DLLM_ERR_CUDA;
                    pending[c] = 1;
                    fused = true;
                } else {
                    rc = k_qlinear_umma(ctx, w, a, tn, last ? out_dev + t0 * w->N : nullptr, last ? nullptr : b);
                }
                char *t = a; a = b; b = t;
            }
            if (rc == DLLM_OK && reduce && !fused) {
                const dllm_qweight *w = m->layers[l1];
                const bool last = l1 + 1 == L;
                void *buf = last ? (void *)(out_dev + t0 * w->N) : (void *)a;      // `a` is the last layer's output after the swap
                if (cudaEventRecord(ctx->tp_ev[2 * c], ctx->stream) != cudaSuccess ||
                    cudaStreamWaitEvent(ctx->comm_stream, ctx->tp_ev[2 * c], 0) != cudaSuccess) { rc = DLLM_ERR_CUDA; break; }
                rc = tp_allreduce_on(ctx, buf, tn * w->N, !last, ctx->comm_stream);
                if (rc == DLLM_OK && cudaEventRecord(ctx->tp_ev[2 * c + 1], ctx->comm_stream) != cudaSuccess) rc = DLLM_ERR_CUDA;
                pending[c] = 1;
            }
        }
        if ((l1 - l0 + 1) % 2 == 1) { char *t = cur; cur = nxt; nxt = t; }     // where the segment's output lives
        l0 = l1 + 1;
    }
    ctx->sm_limit = 0;
    for (int c = 0; c < chunks; ++c)
        if (pending[c] && cudaStreamWaitEvent(ctx->stream, ctx->tp_ev[2 * c + 1], 0) != cudaSuccess && rc == DLLM_OK) rc = DLLM_ERR_CUDA;
    if (rc == DLLM_ERR_CUDA && !ctx->err[0]) DLLM_SET_ERR(ctx, "CUDA error in the overlapped tensor-parallel forward: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

// The widths that flow through the stack must chain: DiffusionModel::forward returns the input's shape (lib.rs:759),
// and every intermediate buffer is sized from the layers' own K / N.  `world` ranks hold COLUMN (N split) / ROW (K split)
// shards; a COLUMN shard feeds a following ROW layer directly, anything else sees the gathered / reduced full width.
static int32_t validate_chain(dllm_ctx *ctx, const dllm_model *m, const int *parallel, size_t world) {
    const size_t L = m->layers.size();
    size_t width = m->hidden;          // width of the activation entering layer l
    bool shard_in = false;             // ... which is this rank's column shard of the previous layer
    for (size_t l = 0; l < L; ++l) {
        const dllm_qweight *w = m->layers[l];
        const int par = parallel ? parallel[l] : 0;
        if (par == 2 && !shard_in) DLLM_FAIL(ctx, DLLM_ERR_SHAPE, "layer %zu is row-parallel but does not follow a column-parallel layer", l);
        if (par != 2 && shard_in) DLLM_FAIL(ctx, DLLM_ERR_SHAPE, "layer %zu follows an un-gathered column shard but is not row-parallel", l);
        if (w->K != width) DLLM_FAIL(ctx, DLLM_ERR_SHAPE, "layer %zu expects K=%zu but receives width %zu", l, w->K, width);
        const bool feeds_row = par == 1 && l + 1 < L && (parallel ? parallel[l + 1] : 0) == 2;
        shard_in = feeds_row;
        width = (par == 1 && !feeds_row) ? w->N * world : w->N;
    }
    if (width != m->hidden) DLLM_FAIL(ctx, DLLM_ERR_SHAPE, "the stack's output width %zu != hidden %zu (output shape must equal input shape, lib.rs:759)", width, m->hidden);
    return DLLM_OK;
}

static int32_t model_forward_tokens(dllm_ctx *ctx, dllm_model *m, const float *x_dev, size_t tokens, float *out_dev,
                                    int32_t path) {
    const size_t L = m->layers.size();
    DLLM_TRY(validate_chain(ctx, m, m->parallel.data(), (size_t)(ctx->tp_world > 1 ? ctx->tp_world : 1)));
    // widest activation of the stack
    size_t maxw = m->hidden;
    for (auto *w : m->layers) { if (w->N > maxw) maxw = w->N; if (w->K > maxw) maxw = w->K; }
    // few tokens: every layer is HBM-bound -> the GEMV kernel (f32 activations at the boundaries)
    bool use_gemv = path == DLLM_PATH_AUTO || path == DLLM_PATH_GEMV;
    for (auto *w : m->layers) use_gemv = use_gemv && k_gemv_supported(w, tokens);
    if (path == DLLM_PATH_GEMV && !use_gemv) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "GEMV path needs 1 <= tokens <= 16");
    bool all_umma = path != DLLM_PATH_SIMT && !use_gemv;
    for (auto *w : m->layers) all_umma = all_umma && k_umma_supported(w, tokens);
    bool any_parallel = false;
    for (int p : m->parallel) any_parallel = any_parallel || p != 0;
    if (path == DLLM_PATH_UMMA && !all_umma) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "tcgen05 path does not support this stack");
    if (path == DLLM_PATH_I8) {
        // int8 denoise mode: bf16 activations between the layers, each quantized per token in front of its integer linear
        for (auto *w : m->layers)
            if (!k_umma_i8_supported(w, tokens))
                DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "int8 mode needs per-tensor quantized weights (group_size 0) with K %% 64 == 0");
        if (any_parallel && ctx->tp_world > 1) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "int8 mode runs replicated stacks only");
        DLLM_TRY(ensure_buf(ctx, ctx->act[0], tokens * maxw * 2));
        DLLM_TRY(ensure_buf(ctx, ctx->act[1], tokens * maxw * 2));
        void *cur = ctx->act[0].p, *nxt = ctx->act[1].p;
        DLLM_TRY(k_f32_to_bf16(ctx, x_dev, tokens * m->layers[0]->K, cur));
        for (size_t l = 0; l < L; ++l) {
            const bool last = l + 1 == L;
            DLLM_TRY(i8_linear(ctx, m->layers[l], cur, tokens, last ? out_dev : nullptr, last ? nullptr : nxt));
            void *t = cur; cur = nxt; nxt = t;
        }
        return DLLM_OK;
    }

    // tensor-parallel stacks whose column-parallel layers all feed a row-parallel one need no all-gather
    bool no_gather = true;
    for (size_t l = 0; l < L; ++l)
        if (m->parallel[l] == 1 && (l + 1 == L || m->parallel[l + 1] != 2)) no_gather = false;
    if (all_umma && any_parallel && no_gather && ctx->tp_world > 1) {
        // default placement (measured on the 7B-class stack at TP2, ms per step): with the peer-to-peer arena the row layers push
        // their partial sums from the GEMM epilogue and only the all-gather half is left at the boundary — one launch per
        // layer, 47.1 (token chunks + overlap: 50.9, chunked GEMMs cost more than the overlap hides); with NCCL the overlap
        // pays: 2 chunks 50.9 against 53.2
        //   the same holds at TP4 once every rank starts at its own token slice (32.8 against 37.0 ms)
        //   (from 4 ranks on the exchange outweighs the GEMMs and two token chunks — the gather half of one chunk under the GEMMs
        //   of the other — win again: 33.7 against 34.0 ms at TP4, 26.1 against 29.5 at TP8)
        const int chunks = ctx->tp_chunks > 0 ? ctx->tp_chunks : env_int("DLLM_TP_CHUNKS", ctx->p2p_arena && ctx->tp_world <= 2 ? 1 : 2);
        if (chunks > 1 && tokens >= (size_t)chunks * 512) return forward_tp_overlapped(ctx, m, x_dev, tokens, out_dev, chunks);
    }
    if (all_umma && (!any_parallel || no_gather)) {
        // bf16 activations between layers; the last layer writes f32.  Row-parallel layers leave partial sums:
        // one NCCL all-reduce at the layer boundary, on the bf16 tensor the next linear reads (f32 for the last).
        char *buf0, *buf1, *recv0 = nullptr;
        if (any_parallel) {
            DLLM_TRY(tp_act_bufs(ctx, tokens * maxw * 2, &buf0, &buf1, &recv0));
        } else {
            DLLM_TRY(ensure_buf(ctx, ctx->act[0], tokens * maxw * 2));
            DLLM_TRY(ensure_buf(ctx, ctx->act[1], tokens * maxw * 2));
            buf0 = (char *)ctx->act[0].p; buf1 = (char *)ctx->act[1].p;
        }
        DLLM_TRY(k_f32_to_bf16(ctx, x_dev, tokens * m->layers[0]->K, buf0));
        void *cur = buf0, *nxt = buf1;
        bool comm_pending = false;         // a reduce / gather kernel on the communication stream has not been joined yet
        for (size_t l = 0; l < L; ++l) {
            const bool last = l + 1 == L;
            if (m->parallel[l] == 2 && !last && tp_fused_rs_ok(ctx, m->layers[l], tokens, recv0)) {
                // the all-gather half can run under the next GEMM when that one is a dense bf16-output kernel on whole 256-token
                // tiles per slice (DLLM_TP_GATED=0 keeps the closing barrier: experiments)
                // DLLM_TP_GATED: 0 (default) = the reduce / gather kernel ends with a barrier; 1 = it ends with arrival counters and the
                // next GEMM gates its activation loads on them, so the peers' rows still in flight overlap its first tiles; 2 = the
                // kernel runs on the communication stream UNDER that GEMM, which leaves it 16 SMs and gates its own slice too.
                // Measured (7B-class, ms per step): TP2 47.7 / 46.8 / 49.8 on one box and 48.8 / 49.5 / 49.8 on another, TP4
                // 33.1 / 34.1 / 34.1 — within the box-to-box spread at TP2 and behind the plain barrier at TP4 (the GEMM on 132 SMs
                // and the programmatic-launch overlap that has to be given up cost what the overlap gains), so the barrier stays.
                const char *gsw = getenv("DLLM_TP_GATED");
                int gate_next = gsw ? atoi(gsw) : 0;
                if (gate_next < 0 || gate_next > 2 ||
                    !(l + 2 < L && m->parallel[l + 1] == 1 && k_umma_gate_supported(ctx, m->layers[l + 1], tokens, ctx->tp_world, cur)))
                    gate_next = 0;
                if (gate_next == 2) {
                    if (!ctx->comm_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
                    while (ctx->tp_ev.size() < 2) {
                        cudaEvent_t e;
                        CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                        ctx->tp_ev.push_back(e);
                    }
                    DLLM_TRY(tp_row_linear_fused(ctx, m->layers[l], cur, tokens, recv0, nxt, ctx->comm_stream, ctx->tp_ev[0], 2));
                    CUDA_TRY(ctx, cudaEventRecord(ctx->tp_ev[1], ctx->comm_stream));
                    comm_pending = true;
                    ctx->sm_limit = ctx->sm_count - 16;            // for the gated GEMM that follows (reset right after it)
                } else {
                    DLLM_TRY(tp_row_linear_fused(ctx, m->layers[l], cur, tokens, recv0, nxt, ctx->stream, nullptr, gate_next));
                }
                void *t = cur; cur = nxt; nxt = t;
                continue;
            }
            {
                const bool under_gather = ctx->sm_limit != 0;      // this GEMM runs next to the reduce / gather kernel it is gated on
                const int32_t rc_l = k_qlinear_umma(ctx, m->layers[l], cur, tokens, last ? out_dev : nullptr, last ? nullptr : nxt);
                ctx->sm_limit = 0;
                if (under_gather) ctx->no_pdl_once = true;         // the next GEMM must not squat on the SMs left to that kernel
                DLLM_TRY(rc_l);
            }
            if (m->parallel[l] == 2) {
                if (last) DLLM_TRY(tp_allreduce(ctx, out_dev, tokens * m->layers[l]->N));
                else DLLM_TRY(tp_allreduce_bf16(ctx, nxt, tokens * m->layers[l]->N));
            }
            void *t = cur; cur = nxt; nxt = t;
        }
        if (comm_pending) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->tp_ev[1], 0));
        return DLLM_OK;
    }
    // General path: f32 activations at the layer boundaries (SIMT layers, and tensor-parallel stacks
    // where the NCCL collective sits at the boundary).  Width bookkeeping:
    //   COLUMN (1): this rank holds W[:, N/p] -> y is the [tokens, N/p] shard; a following ROW layer
    //               consumes it directly (its K shard), anything else needs the all-gather;
    //   ROW    (2): input is the K shard, y is a partial sum over K -> all-reduce -> full [tokens, N].
    const size_t wmul = (size_t)(ctx->tp_world > 1 ? ctx->tp_world : 1);
    DLLM_TRY(ensure_buf(ctx, ctx->act[0], tokens * maxw * 4 * wmul));
    DLLM_TRY(ensure_buf(ctx, ctx->act[1], tokens * maxw * 4 * wmul));
    DLLM_TRY(ensure_buf(ctx, ctx->act[2], tokens * maxw * 2));
    float *bufs[2] = {(float *)ctx->act[0].p, (float *)ctx->act[1].p};
    const float *cur = x_dev;
    int next_buf = 0;
    for (size_t l = 0; l < L; ++l) {
        const bool last = l + 1 == L;
        dllm_qweight *w = m->layers[l];
        const int par = m->parallel[l];
        const bool gather = par == 1 && (last || m->parallel[l + 1] != 2);
        float *dst = (last && !gather) ? out_dev : bufs[next_buf];
        if (use_gemv) {
            DLLM_TRY(k_qlinear_gemv(ctx, w, cur, tokens, dst));
        } else if (path != DLLM_PATH_SIMT && k_umma_supported(w, tokens)) {
            DLLM_TRY(k_f32_to_bf16(ctx, cur, tokens * w->K, ctx->act[2].p));
            DLLM_TRY(k_qlinear_umma(ctx, w, ctx->act[2].p, tokens, dst, nullptr));
        } else {
            DLLM_TRY(k_qlinear_simt(ctx, w, cur, tokens, dst));
        }
        if (par == 2) {
            DLLM_TRY(tp_allreduce(ctx, dst, tokens * w->N));
        } else if (gather) {
            float *full = last ? out_dev : bufs[next_buf ^ 1];
            DLLM_TRY(tp_allgather_cols(ctx, dst, tokens, w->N, full));
            dst = full;
        }
        cur = dst;
        next_buf = (dst == bufs[0]) ? 1 : 0;
    }
    return DLLM_OK;
}

extern "C" {

int32_t dllm_beta_schedule(int32_t kind, size_t T, float beta_start, float beta_end, float *betas) {
    return beta_schedule_host(kind, T, beta_start, beta_end, betas);
}

uint8_t dllm_progressive_bits(size_t num_steps, size_t t, uint8_t decode_bits, uint8_t min_decode_bits,
                              int32_t *is_prefill) {
    if (is_prefill) *is_prefill = t > num_steps / 2;                       // lib.rs:886
    if (t > num_steps) return 0;       // `num_steps - t` underflows: the reference panics (debug) — 0 bits is "no valid width"
    volatile float progress = (float)(num_steps - t) / (float)(num_steps / 2);   // :895
    volatile float a = (float)decode_bits * (1.0f - progress);
    volatile float b = (float)min_decode_bits * progress;
    volatile float target = a + b;                                         // :896-897
    float v = target;
    if (v != v || v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

int32_t dllm_model_create(dllm_ctx *ctx, size_t hidden, dllm_qweight *const *layers, size_t n_layers,
                          size_t num_timesteps, int32_t beta_kind, float beta_start, float beta_end,
                          dllm_model **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, out && layers, DLLM_ERR_NULL, "null pointer");
    *out = nullptr;
    ARG_CHECK(ctx, n_layers > 0 && hidden > 0, DLLM_ERR_INVALID_PARAMS, "empty model");
    ARG_CHECK(ctx, num_timesteps > 0, DLLM_ERR_INVALID_PARAMS, "num_timesteps must be > 0 (lib.rs:547 panics)");
    for (size_t l = 0; l < n_layers; ++l) ARG_CHECK(ctx, layers[l], DLLM_ERR_NULL, "layer %zu is null", l);
    ARG_CHECK(ctx, layers[0]->K == hidden, DLLM_ERR_SHAPE, "first layer K=%zu != hidden=%zu", layers[0]->K, hidden);
    dllm_model *m = new (std::nothrow) dllm_model();
    if (!m) return DLLM_ERR_OOM;
    m->hidden = hidden;
    m->layers.assign(layers, layers + n_layers);
    m->parallel.assign(n_layers, 0);
    m->T = num_timesteps;
    m->device = ctx->device;
    if (ctx->tp_world <= 1) {   // shards of a tensor-parallel group are validated with their plan (set_parallel / forward)
        int32_t vrc = validate_chain(ctx, m, nullptr, 1);
        if (vrc != DLLM_OK) { delete m; return vrc; }
    }
    m->betas.resize(num_timesteps);
    int32_t rc = beta_schedule_host(beta_kind, num_timesteps, beta_start, beta_end, m->betas.data());
    if (rc != DLLM_OK) { delete m; DLLM_FAIL(ctx, rc, "unknown beta schedule %d", beta_kind); }
    m->alpha_bars.resize(num_timesteps);
    m->alpha_bars[0] = 1.0f;                                              // lib.rs:1162-1165
    for (size_t i = 1; i < num_timesteps; ++i) {
        volatile float a = 1.0f - m->betas[i - 1];
        volatile float p = m->alpha_bars[i - 1] * a;
        m->alpha_bars[i] = p;
    }
    // coefficient tables for every timestep, resident on the device: no host work per step
    std::vector<float> tab(num_timesteps * 4);
    for (int g = 0; g < 2; ++g) {
        for (size_t t = 0; t < num_timesteps; ++t) p_sample_coeffs_host(m, t, g, tab.data() + 4 * t);
        if (cudaMalloc(&m->d_coef_table[g], tab.size() * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(m->d_coef_table[g], tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError();
            dllm_model_destroy(m);
            DLLM_FAIL(ctx, DLLM_ERR_CUDA, "coefficient table upload failed");
        }
    }
    {   // add_noise's two factors per timestep (lib.rs:1131-1132)
        std::vector<float> nt(num_timesteps * 2);
        for (size_t t = 0; t < num_timesteps; ++t) {
            volatile float om = 1.0f - m->alpha_bars[t];
            nt[2 * t] = sqrtf(m->alpha_bars[t]);
            nt[2 * t + 1] = sqrtf(om);
        }
        if (cudaMalloc(&m->d_noise_table, nt.size() * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(m->d_noise_table, nt.data(), nt.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError();
            dllm_model_destroy(m);
            DLLM_FAIL(ctx, DLLM_ERR_CUDA, "noise table upload failed");
        }
    }
    *out = m;
    return DLLM_OK;
}

void dllm_model_destroy(dllm_model *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    for (int g = 0; g < 2; ++g) if (m->d_coef_table[g]) cudaFree(m->d_coef_table[g]);
    if (m->d_noise_table) cudaFree(m->d_noise_table);
    if (m->d_state) cudaFree(m->d_state);
    if (m->step_graph) cudaGraphExecDestroy(m->step_graph);
    if (m->d_rowmap) cudaFree(m->d_rowmap);
    delete m;   // layers are owned by the caller
}

int32_t dllm_model_set_parallel(dllm_ctx *ctx, dllm_model *m, const int32_t *parallel, size_t n_layers) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m && parallel, DLLM_ERR_NULL, "null pointer");
    ARG_CHECK(ctx, n_layers == m->layers.size(), DLLM_ERR_SHAPE, "expected %zu entries", m->layers.size());
    std::vector<int> plan(n_layers);
    for (size_t l = 0; l < n_layers; ++l) {
        ARG_CHECK(ctx, parallel[l] >= 0 && parallel[l] <= 2, DLLM_ERR_INVALID_PARAMS, "parallel[%zu]=%d", l, parallel[l]);
        plan[l] = parallel[l];
    }
    DLLM_TRY(validate_chain(ctx, m, plan.data(), (size_t)(ctx->tp_world > 1 ? ctx->tp_world : 1)));
    m->parallel = plan;
    return DLLM_OK;
}

int32_t dllm_model_forward_dev(dllm_ctx *ctx, dllm_model *m, const float *x_dev, size_t batch, size_t feat,
                               float *noise_pred_dev, int32_t path) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    ARG_CHECK(ctx, feat % m->hidden == 0, DLLM_ERR_SHAPE, "feature dim %zu is not a multiple of hidden %zu", feat, m->hidden);
    if (batch * feat == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_dev && noise_pred_dev, DLLM_ERR_NULL, "null device pointer");
    return model_forward_tokens(ctx, m, x_dev, batch * (feat / m->hidden), noise_pred_dev, path);
}

int32_t dllm_model_forward(dllm_ctx *ctx, dllm_model *m, const float *x, const size_t *t, size_t batch, size_t feat,
                           float *noise_pred, int32_t path) {
    CTX_CHECK(ctx);
    (void)t;   // SimpleDiffusionModel::forward ignores the timestep (lib.rs:806 `_t`)
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    const size_t n = batch * feat;
    ARG_CHECK(ctx, n == 0 || (x && noise_pred), DLLM_ERR_NULL, "null pointer");
    if (n == 0) return DLLM_OK;
    void *dx, *dy;
    DLLM_TRY(stage_in(ctx, 4, x, n * sizeof(float), &dx));
    DLLM_TRY(stage_out_buf(ctx, 5, n * sizeof(float), &dy));
    DLLM_TRY(dllm_model_forward_dev(ctx, m, (const float *)dx, batch, feat, (float *)dy, path));
    DLLM_TRY(copy_out(ctx, noise_pred, dy, n * sizeof(float)));
    return sync(ctx);
}

static int32_t upload_rowmap(dllm_ctx *ctx, dllm_model *m, const size_t *t, size_t batch) {
    if (batch > m->rowmap_cap) {
        if (m->d_rowmap) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(m->d_rowmap); m->d_rowmap = nullptr; }
        CUDA_TRY(ctx, cudaMalloc(&m->d_rowmap, batch * sizeof(int)));
        m->rowmap_cap = batch;
    }
    // staged through the model's own host vector: it outlives the copy (every caller synchronises before it returns,
    // and a ctx is used by one thread at a time), so no extra stream synchronisation is needed here
    m->h_rowmap.resize(batch);
    for (size_t b = 0; b < batch; ++b) m->h_rowmap[b] = (int)(t[b] < m->T - 1 ? t[b] : m->T - 1);   // lib.rs:1174 clamp
    CUDA_TRY(ctx, cudaMemcpyAsync(m->d_rowmap, m->h_rowmap.data(), batch * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += batch * sizeof(int);
    return DLLM_OK;
}

int32_t dllm_p_sample(dllm_ctx *ctx, dllm_model *m, const float *x_t, const float *noise_pred, const float *z,
                      const size_t *t, size_t batch, size_t feat, int32_t guard_t0, float *x_prev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    const size_t n = batch * feat;
    if (n == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_t && noise_pred && t && x_prev, DLLM_ERR_NULL, "null pointer");
    const bool add_noise = z != nullptr && t[0] > 0;      // lib.rs:1199-1205
    void *dx, *dp, *dz = nullptr, *dout;
    DLLM_TRY(stage_in(ctx, 4, x_t, n * sizeof(float), &dx));
    DLLM_TRY(stage_in(ctx, 5, noise_pred, n * sizeof(float), &dp));
    if (add_noise) DLLM_TRY(stage_in(ctx, 6, z, n * sizeof(float), &dz));
    DLLM_TRY(stage_out_buf(ctx, 7, n * sizeof(float), &dout));
    DLLM_TRY(upload_rowmap(ctx, m, t, batch));
    DLLM_TRY(k_p_sample(ctx, (const float *)dx, (const float *)dp, (const float *)dz, m->d_coef_table[guard_t0 ? 1 : 0],
                        m->d_rowmap, 0, batch, feat, (float *)dout));
    DLLM_TRY(copy_out(ctx, x_prev, dout, n * sizeof(float)));
    return sync(ctx);
}

int32_t dllm_add_noise_dev(dllm_ctx *ctx, dllm_model *m, const float *x_start_dev, const float *noise_dev, size_t t,
                           size_t batch, size_t feat, float *noisy_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    if (batch * feat == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_start_dev && noise_dev && noisy_dev, DLLM_ERR_NULL, "null device pointer");
    const int row = (int)(t < m->T - 1 ? t : m->T - 1);                       // lib.rs:1123
    return k_add_noise(ctx, x_start_dev, noise_dev, m->d_noise_table, nullptr, row, batch, feat, noisy_dev);
}

int32_t dllm_add_noise(dllm_ctx *ctx, dllm_model *m, const float *x_start, const float *noise, const size_t *t,
                       size_t batch, size_t feat, float *noisy) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    const size_t n = batch * feat;
    if (n == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_start && noise && t && noisy, DLLM_ERR_NULL, "null pointer (the noise is an input: the reference's own draw is an unseeded thread_rng, lib.rs:1107-1109)");
    void *dx, *dn, *dout;
    DLLM_TRY(stage_in(ctx, 4, x_start, n * sizeof(float), &dx));
    DLLM_TRY(stage_in(ctx, 5, noise, n * sizeof(float), &dn));
    DLLM_TRY(stage_out_buf(ctx, 7, n * sizeof(float), &dout));
    DLLM_TRY(upload_rowmap(ctx, m, t, batch));
    DLLM_TRY(k_add_noise(ctx, (const float *)dx, (const float *)dn, m->d_noise_table, m->d_rowmap, 0, batch, feat, (float *)dout));
    DLLM_TRY(copy_out(ctx, noisy, dout, n * sizeof(float)));
    return sync(ctx);
}

int32_t dllm_denoise_step_dev(dllm_ctx *ctx, dllm_model *m, float *x_dev, const float *z_dev, size_t t, size_t batch,
                              size_t feat, int32_t guard_t0, int32_t path) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    const size_t n = batch * feat;
    if (n == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_dev, DLLM_ERR_NULL, "null device pointer");
    void *dpred;
    DLLM_TRY(stage_out_buf(ctx, 7, n * sizeof(float), &dpred));
    DLLM_TRY(dllm_model_forward_dev(ctx, m, x_dev, batch, feat, (float *)dpred, path));   // lib.rs:924
    // p_sample in place: every element is read once before it is written (lib.rs:925)
    const int row = (int)(t < m->T - 1 ? t : m->T - 1);
    return k_p_sample(ctx, x_dev, (const float *)dpred, t > 0 ? z_dev : nullptr, m->d_coef_table[guard_t0 ? 1 : 0],
                      nullptr, row, batch, feat, x_dev);
}

int32_t dllm_denoise_step(dllm_ctx *ctx, dllm_model *m, float *x, const float *z, size_t t, size_t batch, size_t feat,
                          int32_t guard_t0, int32_t path) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    const size_t n = batch * feat;
    if (n == 0) return DLLM_OK;
    ARG_CHECK(ctx, x, DLLM_ERR_NULL, "null pointer");
    if (!ctx->copy_stream) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
        for (auto &e : ctx->ev_step) CUDA_TRY(ctx, cudaEventCreate(&e));
    }
    const bool noise = z && t > 0;
    void *dx = nullptr, *dz = nullptr, *dpred = nullptr;
    // buffers first (growing one synchronises), then the copies: x on the compute stream — the forward needs it — and the
    // noise on the copy stream, under the forward pass; p_sample waits for it
    DLLM_TRY(ensure_buf(ctx, ctx->ws[4], n * sizeof(float)));
    if (noise) DLLM_TRY(ensure_buf(ctx, ctx->ws[6], n * sizeof(float)));
    DLLM_TRY(stage_out_buf(ctx, 7, n * sizeof(float), &dpred));
    dx = ctx->ws[4].p;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_step[0], ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dx, x, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += n * sizeof(float);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_step[1], ctx->stream));
    if (noise) {
        dz = ctx->ws[6].p;
        CUDA_TRY(ctx, cudaMemcpyAsync(dz, z, n * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
        ctx->h2d_bytes += n * sizeof(float);
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    }
    DLLM_TRY(dllm_model_forward_dev(ctx, m, (const float *)dx, batch, feat, (float *)dpred, path));   // lib.rs:924
    if (noise) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    const int row = (int)(t < m->T - 1 ? t : m->T - 1);
    DLLM_TRY(k_p_sample(ctx, (const float *)dx, (const float *)dpred, (const float *)dz, m->d_coef_table[guard_t0 ? 1 : 0],
                        nullptr, row, batch, feat, (float *)dx));                                       // lib.rs:925
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_step[2], ctx->stream));
    DLLM_TRY(copy_out(ctx, x, dx, n * sizeof(float)));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_step[3], ctx->stream));
    DLLM_TRY(sync(ctx));
    for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&ctx->step_ms[i], ctx->ev_step[i], ctx->ev_step[i + 1]);
    return DLLM_OK;
}

int32_t dllm_last_step_breakdown(const dllm_ctx *ctx, float *h2d_ms, float *compute_ms, float *d2h_ms) {
    if (!ctx) return DLLM_ERR_NULL;
    if (h2d_ms) *h2d_ms = ctx->step_ms[0];
    if (compute_ms) *compute_ms = ctx->step_ms[1];
    if (d2h_ms) *d2h_ms = ctx->step_ms[2];
    return DLLM_OK;
}

// ---- seeded loop: noise from the counter-based generator ("dllm_noise v1", csrc/noise.cuh), nothing uploaded per step ----
int32_t dllm_noise_fill_dev(dllm_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t first, size_t n, float *out_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, n == 0 || out_dev, DLLM_ERR_NULL, "null device pointer");
    return k_noise_fill(ctx, seed, stream, first, n, out_dev);
}

int32_t dllm_noise_fill(dllm_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t first, size_t n, float *out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, n == 0 || out, DLLM_ERR_NULL, "null pointer");
    if (n == 0) return DLLM_OK;
    void *d;
    DLLM_TRY(stage_out_buf(ctx, 6, n * sizeof(float), &d));
    DLLM_TRY(k_noise_fill(ctx, seed, stream, first, n, (float *)d));
    DLLM_TRY(copy_out(ctx, out, d, n * sizeof(float)));
    return sync(ctx);
}

// forward + p_sample with in-kernel noise; `state` != null: t and seed are read on the device (graph replay)
static int32_t step_seeded(dllm_ctx *ctx, dllm_model *m, float *x_dev, const void *state, size_t t, uint64_t seed, size_t batch,
                           size_t feat, int32_t guard_t0, int32_t path) {
    const size_t n = batch * feat;
    void *dpred;
    DLLM_TRY(stage_out_buf(ctx, 7, n * sizeof(float), &dpred));
    DLLM_TRY(dllm_model_forward_dev(ctx, m, x_dev, batch, feat, (float *)dpred, path));   // lib.rs:924
    return k_p_sample_seeded(ctx, x_dev, (const float *)dpred, m->d_coef_table[guard_t0 ? 1 : 0], state, (int)t, seed,
                             (int)m->T, n, x_dev);                                        // lib.rs:925, in place
}

int32_t dllm_denoise_step_seeded_dev(dllm_ctx *ctx, dllm_model *m, float *x_dev, uint64_t seed, size_t t, size_t batch,
                                     size_t feat, int32_t guard_t0, int32_t path) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    if (batch * feat == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_dev, DLLM_ERR_NULL, "null device pointer");
    ARG_CHECK(ctx, t <= 0x7fffffffu, DLLM_ERR_INVALID_PARAMS, "timestep out of range");
    return step_seeded(ctx, m, x_dev, nullptr, t, seed, batch, feat, guard_t0, path);
}

// p_sample alone on device tensors (lib.rs:1152-1215) — for callers that produce noise_pred themselves (the cached branch of
// the sampling loop, lib.rs:910-921).  x_prev_dev may alias x_t_dev.  z_dev == NULL or t == 0: no noise term.
int32_t dllm_p_sample_dev(dllm_ctx *ctx, dllm_model *m, const float *x_t_dev, const float *noise_pred_dev, const float *z_dev,
                          size_t t, size_t batch, size_t feat, int32_t guard_t0, float *x_prev_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    if (batch * feat == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_t_dev && noise_pred_dev && x_prev_dev, DLLM_ERR_NULL, "null device pointer");
    const int row = (int)(t < m->T - 1 ? t : m->T - 1);
    return k_p_sample(ctx, x_t_dev, noise_pred_dev, t > 0 ? z_dev : nullptr, m->d_coef_table[guard_t0 ? 1 : 0], nullptr, row, batch,
                      feat, x_prev_dev);
}

// the same with the noise of timestep t drawn inside the kernel from the counter-based generator (stream t of `seed`)
int32_t dllm_p_sample_seeded_dev(dllm_ctx *ctx, dllm_model *m, const float *x_t_dev, const float *noise_pred_dev, uint64_t seed,
                                 size_t t, size_t batch, size_t feat, int32_t guard_t0, float *x_prev_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    if (batch * feat == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_t_dev && noise_pred_dev && x_prev_dev, DLLM_ERR_NULL, "null device pointer");
    ARG_CHECK(ctx, t <= 0x7fffffffu, DLLM_ERR_INVALID_PARAMS, "timestep out of range");
    return k_p_sample_seeded(ctx, x_t_dev, noise_pred_dev, m->d_coef_table[guard_t0 ? 1 : 0], nullptr, (int)t, seed, (int)m->T,
                             batch * feat, x_prev_dev);
}

int32_t dllm_sample_seeded_dev(dllm_ctx *ctx, dllm_model *m, float *x_dev, uint64_t seed, size_t batch, size_t feat,
                               size_t num_steps, int32_t guard_t0, int32_t path, int32_t use_graph) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    if (batch * feat == 0 || num_steps == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_dev, DLLM_ERR_NULL, "null device pointer");
    ARG_CHECK(ctx, num_steps <= 0x7fffffffu, DLLM_ERR_INVALID_PARAMS, "num_steps out of range");
    // the first step runs eagerly: it sizes every workspace and sets the kernels' attributes (nothing may allocate while a
    // stream is being captured)
    size_t t = num_steps - 1;                                           // lib.rs:881 `(0..num_steps).rev()`
    DLLM_TRY(step_seeded(ctx, m, x_dev, nullptr, t, seed, batch, feat, guard_t0, path));
    // graphs: one captured step replayed with t in device memory.  Not under tensor parallelism (NCCL calls stay eager) and
    // not while launches are being bracketed with profiling events.
    const bool graph = use_graph && ctx->tp_world <= 1 && !ctx->prof_on && num_steps >= 3;
    if (!graph) {
        while (t-- > 0) DLLM_TRY(step_seeded(ctx, m, x_dev, nullptr, t, seed, batch, feat, guard_t0, path));
        return DLLM_OK;
    }
    if (!m->d_state) CUDA_TRY(ctx, cudaMalloc(&m->d_state, 16));
    struct { int t; int pad; unsigned long long seed; } st = {(int)(num_steps - 2), 0, (unsigned long long)seed};
    CUDA_TRY(ctx, cudaMemcpyAsync(m->d_state, &st, sizeof(st), cudaMemcpyHostToDevice, ctx->stream));   // pageable source: staged before the call returns
    const bool cached = m->step_graph && m->graph_x == x_dev && m->graph_batch == batch && m->graph_feat == feat &&
                        m->graph_path == path && m->graph_guard == (guard_t0 ? 1 : 0);
    if (!cached) {
        if (m->step_graph) { cudaGraphExecDestroy(m->step_graph); m->step_graph = nullptr; }
        const uint64_t l0 = ctx->launches;
        CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int32_t rc = step_seeded(ctx, m, x_dev, m->d_state, 0, 0, batch, feat, guard_t0, path);
        if (rc == DLLM_OK) rc = k_sample_state_step(ctx, m->d_state);
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        if (rc != DLLM_OK || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            if (rc != DLLM_OK) return rc;
            DLLM_FAIL(ctx, DLLM_ERR_CUDA, "stream capture of the denoise step failed: %s", cudaGetErrorString(e));
        }
        e = cudaGraphInstantiate(&m->step_graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { m->step_graph = nullptr; DLLM_FAIL(ctx, DLLM_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
        m->graph_launches = (size_t)(ctx->launches - l0);
        ctx->launches = l0;                                             // captured, not launched
        m->graph_x = x_dev; m->graph_batch = batch; m->graph_feat = feat; m->graph_path = path; m->graph_guard = guard_t0 ? 1 : 0;
    }
    for (size_t s = 0; s + 1 < num_steps; ++s) {
        CUDA_TRY(ctx, cudaGraphLaunch(m->step_graph, ctx->stream));
        ctx->launches += m->graph_launches;
        ctx->graph_replays++;
    }
    return DLLM_OK;
}

int32_t dllm_sample_seeded(dllm_ctx *ctx, dllm_model *m, const float *x0, uint64_t seed, size_t batch, size_t feat,
                           size_t num_steps, int32_t guard_t0, int32_t path, int32_t use_graph, float *x_out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    const size_t n = batch * feat;
    if (n == 0) return DLLM_OK;
    ARG_CHECK(ctx, x_out, DLLM_ERR_NULL, "null pointer");
    void *dx;
    if (x0) {
        DLLM_TRY(stage_in(ctx, 4, x0, n * sizeof(float), &dx));
    } else {                                                            // lib.rs:875-878: x ~ N(0,1), here stream `num_steps`
        DLLM_TRY(stage_out_buf(ctx, 4, n * sizeof(float), &dx));
        DLLM_TRY(k_noise_fill(ctx, seed, (unsigned long long)num_steps, 0, n, (float *)dx));
    }
    DLLM_TRY(dllm_sample_seeded_dev(ctx, m, (float *)dx, seed, batch, feat, num_steps, guard_t0, path, use_graph));
    DLLM_TRY(copy_out(ctx, x_out, dx, n * sizeof(float)));
    return sync(ctx);
}

int32_t dllm_sample(dllm_ctx *ctx, dllm_model *m, const float *x0, const float *noises, size_t batch, size_t feat,
                    size_t num_steps, int32_t guard_t0, int32_t path, float *x_out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, m, DLLM_ERR_NULL, "null model");
    const size_t n = batch * feat;
    if (n == 0) return DLLM_OK;
    ARG_CHECK(ctx, x0 && x_out, DLLM_ERR_NULL, "null pointer");
    void *dx, *dz = nullptr;
    DLLM_TRY(stage_in(ctx, 4, x0, n * sizeof(float), &dx));          // lib.rs:875-878 (noise injected)
    if (noises) DLLM_TRY(stage_out_buf(ctx, 6, n * sizeof(float), &dz));
    for (size_t t = num_steps; t-- > 0;) {                           // lib.rs:881 `(0..num_steps).rev()`
        if (noises && t > 0)
            CUDA_TRY(ctx, cudaMemcpyAsync(dz, noises + t * n, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d_bytes += n * sizeof(float);
        DLLM_TRY(dllm_denoise_step_dev(ctx, m, (float *)dx, noises ? (const float *)dz : nullptr, t, batch, feat,
                                       guard_t0, path));
    }
    DLLM_TRY(copy_out(ctx, x_out, dx, n * sizeof(float)));
    return sync(ctx);
}

}  // extern "C"

// ==========================================================================================
// KV cache entry
// ==========================================================================================
struct dllm_kv {
    size_t L = 0, S = 0, H = 0;
    size_t cap = 0;                   // tokens per layer the buffers have room for: layer l's rows start at l * cap (cap == S: contiguous)
    int bits = 8, scheme = DLLM_KV_TENSOR_B;
    bool packed = false;
    size_t code_bytes = 0;            // per tensor
    uint8_t *d_codes[2] = {nullptr, nullptr};
    float *d_params[2] = {nullptr, nullptr};   // B: {scale, zp, min, max}
    float *d_rows[2] = {nullptr, nullptr};     // D: [rows] scales then [rows] zps
    float c_scale = 0.f;
    bool sharded = false;             // TENSOR_B: this entry holds one rank's token rows; min / max are all-reduced over the group
    int device = 0;
};

static size_t kv_row_bytes(const dllm_kv *kv) { return kv->packed ? kv->H * kv->bits / 8 : kv->H; }

// quantize tokens [s0, s0 + t) of every layer from src [L, t, H] (schemes with per-row or fixed parameters: a row's codes
// depend on that row only, so tokens already in the cache never change)
static int32_t kv_quantize_rows(dllm_ctx *ctx, dllm_kv *kv, int which, const float *src_dev, size_t s0, size_t t,
                                size_t src_layer_rows = 0) {
    const int pack = kv->packed ? kv->bits : 0;
    const size_t rb = kv_row_bytes(kv), rows_cap = kv->L * kv->cap;
    if (src_layer_rows == 0) src_layer_rows = t;          // rows between two layers of the source (dense by default)
    const bool contiguous = kv->cap == t && s0 == 0 && src_layer_rows == t;
    const size_t n_l = contiguous ? kv->L : 1, loops = contiguous ? 1 : kv->L;
    for (size_t l = 0; l < loops; ++l) {
        const float *src = src_dev + l * src_layer_rows * kv->H;
        const size_t r0 = l * kv->cap + s0;
        if (kv->scheme == DLLM_KV_ROW_D)
            DLLM_TRY(k_quant_d_rows(ctx, src, n_l * t, kv->H, nullptr, 1, kv->bits, pack, kv->d_codes[which] + r0 * rb,
                                    kv->d_rows[which] + r0, kv->d_rows[which] + rows_cap + r0));
        else
            DLLM_TRY(k_encode_cd(ctx, src, n_l * t * kv->H, kv->bits, pack, kv->c_scale, 0.0f, kv->d_codes[which] + r0 * rb));
    }
    return DLLM_OK;
}

static int32_t kv_quantize_one(dllm_ctx *ctx, dllm_kv *kv, int which, const float *src_dev) {
    const size_t rows = kv->L * kv->S, n = rows * kv->H;
    const int pack = kv->packed ? kv->bits : 0;
    switch (kv->scheme) {
        case DLLM_KV_TENSOR_B:
            if (kv->cap != kv->S) DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "per-tensor KV entries have no spare capacity");
            DLLM_TRY(k_minmax(ctx, src_dev, n, kv->bits, kv->d_params[which]));     // (an empty shard leaves {+inf, -inf}: neutral)
            if (kv->sharded && ctx->tp_world > 1) {
                DLLM_TRY(tp_allreduce_minmax(ctx, kv->d_params[which]));
                DLLM_TRY(k_params_from_minmax(ctx, kv->bits, kv->d_params[which]));
            }
            return k_encode_b(ctx, src_dev, n, kv->bits, pack, kv->d_params[which], 0.f, 0.f, kv->d_codes[which]);
        case DLLM_KV_ROW_D:
        case DLLM_KV_FIXED_C:
            return kv_quantize_rows(ctx, kv, which, src_dev, 0, kv->S);
    }
    return DLLM_ERR_INVALID_PARAMS;
}

// dst [L, S, H] contiguous
static int32_t kv_dequantize_one(dllm_ctx *ctx, const dllm_kv *kv, int which, float *dst_dev) {
    const int pack = kv->packed ? kv->bits : 0;
    const size_t rb = kv_row_bytes(kv), rows_cap = kv->L * kv->cap;
    if (kv->scheme == DLLM_KV_TENSOR_B)
        return k_decode_ab(ctx, kv->d_codes[which], kv->L * kv->S * kv->H, pack, kv->d_params[which], 0.f, 0.f, dst_dev);
    const bool contiguous = kv->cap == kv->S;
    const size_t n_l = contiguous ? kv->L : 1, loops = contiguous ? 1 : kv->L;
    for (size_t l = 0; l < loops; ++l) {
        const size_t r0 = l * kv->cap;
        float *dst = dst_dev + l * kv->S * kv->H;
        if (kv->scheme == DLLM_KV_ROW_D)
            DLLM_TRY(k_decode_cd(ctx, kv->d_codes[which] + r0 * rb, n_l * kv->S * kv->H, pack, 0.f, 0.f, kv->d_rows[which] + r0,
                                 kv->d_rows[which] + rows_cap + r0, kv->H, dst));
        else
            DLLM_TRY(k_decode_cd(ctx, kv->d_codes[which] + r0 * rb, n_l * kv->S * kv->H, pack, kv->c_scale, 0.0f, nullptr, nullptr, 0, dst));
    }
    return DLLM_OK;
}

extern "C" {

void dllm_kv_destroy(dllm_kv *kv) {
    if (!kv) return;
    cudaSetDevice(kv->device);
    for (int i = 0; i < 2; ++i) {
        if (kv->d_codes[i]) cudaFree(kv->d_codes[i]);
        if (kv->d_params[i]) cudaFree(kv->d_params[i]);
        if (kv->d_rows[i]) cudaFree(kv->d_rows[i]);
    }
    delete kv;
}

int32_t dllm_kv_update_dev(dllm_ctx *ctx, dllm_kv *kv, const float *keys_dev, const float *values_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kv, DLLM_ERR_NULL, "null kv entry");
    if (kv->L * kv->S * kv->H == 0) return DLLM_OK;
    ARG_CHECK(ctx, keys_dev && values_dev, DLLM_ERR_NULL, "null device pointer");
    DLLM_TRY(kv_quantize_one(ctx, kv, 0, keys_dev));
    return kv_quantize_one(ctx, kv, 1, values_dev);
}

// an entry with room for `capacity` tokens per layer, holding `seq` of them (not yet quantized)
static int32_t kv_alloc(dllm_ctx *ctx, size_t layers, size_t seq, size_t capacity, size_t hidden, uint8_t bits, int32_t scheme,
                        dllm_kv **out) {
    ARG_CHECK(ctx, scheme >= DLLM_KV_TENSOR_B && scheme <= DLLM_KV_FIXED_C, DLLM_ERR_INVALID_PARAMS, "unknown scheme %d", scheme);
    if (scheme == DLLM_KV_TENSOR_B)
        ARG_CHECK(ctx, bits >= 1 && bits <= 8, DLLM_ERR_INVALID_PARAMS, "Bits must be between 1 and 8 (got %d)", (int)bits);
    ARG_CHECK(ctx, bits <= 30, DLLM_ERR_INVALID_PARAMS, "bits %d overflows", (int)bits);
    dllm_kv *kv = new (std::nothrow) dllm_kv();
    if (!kv) return DLLM_ERR_OOM;
    kv->L = layers; kv->S = seq; kv->cap = capacity; kv->H = hidden; kv->bits = bits; kv->scheme = scheme; kv->device = ctx->device;
    const size_t rows = layers * capacity, n = rows * hidden;
    kv->packed = pack_width_ok(bits) && bits != 8 && hidden % 8 == 0 && (hidden * bits / 8) % 4 == 0 &&
                 (scheme != DLLM_KV_ROW_D || hidden <= 16384);
    kv->code_bytes = kv->packed ? n * bits / 8 : n;
    kv->c_scale = dllm_bitquantizer_scale(bits);
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaMalloc(&kv->d_codes[i], kv->code_bytes ? kv->code_bytes : 16);
        if (e == cudaSuccess) e = cudaMalloc(&kv->d_params[i], 4 * sizeof(float));
        if (e == cudaSuccess && scheme == DLLM_KV_ROW_D) e = cudaMalloc(&kv->d_rows[i], (2 * rows + 4) * sizeof(float));
    }
    if (e != cudaSuccess) { cudaGetLastError(); dllm_kv_destroy(kv); DLLM_FAIL(ctx, DLLM_ERR_OOM, "cudaMalloc failed for the KV entry"); }
    *out = kv;
    return DLLM_OK;
}

int32_t dllm_kv_quantize_dev(dllm_ctx *ctx, const float *keys_dev, const float *values_dev, size_t layers, size_t seq,
                             size_t hidden, uint8_t bits, int32_t scheme, dllm_kv **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, out, DLLM_ERR_NULL, "null out pointer");
    *out = nullptr;
    dllm_kv *kv = nullptr;
    DLLM_TRY(kv_alloc(ctx, layers, seq, seq, hidden, bits, scheme, &kv));
    int32_t rc = dllm_kv_update_dev(ctx, kv, keys_dev, values_dev);
    if (rc != DLLM_OK) { dllm_kv_destroy(kv); return rc; }
    *out = kv;
    return DLLM_OK;
}

// This rank's slice [layers, seq_local, hidden] of K and V whose token rows are sharded over the ranks of the context's group
// (dllm_tp_init): per-token (ROW_D) and fixed-scale (FIXED_C) entries need nothing from the other ranks; a per-tensor entry
// (TENSOR_B: QuantizedKVCacheEntry::new quantizes each tensor with ONE scale / zero-point, quantization.rs:140-157) takes the
// min / max over all ranks first — two floats per tensor on NVLink — so every rank encodes with the
// parameters of the WHOLE tensor and the concatenated codes are bit-identical to a single-GPU quantization.
int32_t dllm_kv_quantize_sharded_dev(dllm_ctx *ctx, const float *keys_dev, const float *values_dev, size_t layers, size_t seq_local,
                                     size_t hidden, uint8_t bits, int32_t scheme, dllm_kv **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, out, DLLM_ERR_NULL, "null out pointer");
    *out = nullptr;
    if (scheme != DLLM_KV_TENSOR_B || ctx->tp_world <= 1)
        return dllm_kv_quantize_dev(ctx, keys_dev, values_dev, layers, seq_local, hidden, bits, scheme, out);
    dllm_kv *kv = nullptr;
    DLLM_TRY(kv_alloc(ctx, layers, seq_local, seq_local, hidden, bits, scheme, &kv));
    kv->sharded = true;
    // every rank takes part in the min / max exchange, also one whose shard is empty
    int32_t rc = DLLM_OK;
    if (layers * seq_local * hidden == 0) {
        for (int i = 0; i < 2 && rc == DLLM_OK; ++i) rc = kv_quantize_one(ctx, kv, i, nullptr);
    } else if (!keys_dev || !values_dev) {
        rc = DLLM_ERR_NULL;
    } else {
        rc = dllm_kv_update_dev(ctx, kv, keys_dev, values_dev);
    }
    if (rc != DLLM_OK) { dllm_kv_destroy(kv); return rc; }
    *out = kv;
    return DLLM_OK;
}

int32_t dllm_kv_create(dllm_ctx *ctx, size_t layers, size_t capacity, size_t hidden, uint8_t bits, int32_t scheme, dllm_kv **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, out, DLLM_ERR_NULL, "null out pointer");
    *out = nullptr;
    ARG_CHECK(ctx, scheme == DLLM_KV_ROW_D || scheme == DLLM_KV_FIXED_C, DLLM_ERR_UNSUPPORTED,
              "only per-token (ROW_D) and fixed-scale (FIXED_C) entries can grow: a per-tensor scale changes every code");
    return kv_alloc(ctx, layers, 0, capacity, hidden, bits, scheme, out);
}

int32_t dllm_kv_append_dev(dllm_ctx *ctx, dllm_kv *kv, const float *keys_new_dev, const float *values_new_dev, size_t t_new) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kv, DLLM_ERR_NULL, "null kv entry");
    ARG_CHECK(ctx, kv->scheme == DLLM_KV_ROW_D || kv->scheme == DLLM_KV_FIXED_C, DLLM_ERR_UNSUPPORTED, "per-tensor KV entries cannot grow");
    ARG_CHECK(ctx, kv->S + t_new <= kv->cap, DLLM_ERR_INDEX, "KV entry is full: %zu + %zu tokens > capacity %zu", kv->S, t_new, kv->cap);
    if (t_new == 0 || kv->L * kv->H == 0) return DLLM_OK;
    ARG_CHECK(ctx, keys_new_dev && values_new_dev, DLLM_ERR_NULL, "null device pointer");
    DLLM_TRY(kv_quantize_rows(ctx, kv, 0, keys_new_dev, kv->S, t_new));
    DLLM_TRY(kv_quantize_rows(ctx, kv, 1, values_new_dev, kv->S, t_new));
    kv->S += t_new;
    return DLLM_OK;
}

int32_t dllm_kv_append(dllm_ctx *ctx, dllm_kv *kv, const float *keys_new, const float *values_new, size_t t_new) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kv, DLLM_ERR_NULL, "null kv entry");
    const size_t n = kv->L * t_new * kv->H;
    ARG_CHECK(ctx, n == 0 || (keys_new && values_new), DLLM_ERR_NULL, "null pointer");
    void *dk, *dv;
    DLLM_TRY(stage_in(ctx, 4, keys_new, n * sizeof(float), &dk));
    DLLM_TRY(stage_in(ctx, 5, values_new, n * sizeof(float), &dv));
    DLLM_TRY(dllm_kv_append_dev(ctx, kv, (const float *)dk, (const float *)dv, t_new));
    return sync(ctx);
}

size_t dllm_kv_seq_len(const dllm_kv *kv) { return kv ? kv->S : 0; }

int32_t dllm_kv_quantize(dllm_ctx *ctx, const float *keys, const float *values, size_t layers, size_t seq,
                         size_t hidden, uint8_t bits, int32_t scheme, dllm_kv **out) {
    CTX_CHECK(ctx);
    const size_t n = layers * seq * hidden;
    ARG_CHECK(ctx, n == 0 || (keys && values), DLLM_ERR_NULL, "null pointer");
    void *dk, *dv;
    DLLM_TRY(stage_in(ctx, 4, keys, n * sizeof(float), &dk));
    DLLM_TRY(stage_in(ctx, 5, values, n * sizeof(float), &dv));
    DLLM_TRY(dllm_kv_quantize_dev(ctx, (const float *)dk, (const float *)dv, layers, seq, hidden, bits, scheme, out));
    return sync(ctx);
}

int32_t dllm_kv_dequantize_dev(dllm_ctx *ctx, const dllm_kv *kv, float *keys_dev, float *values_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kv, DLLM_ERR_NULL, "null kv entry");
    if (kv->L * kv->S * kv->H == 0) return DLLM_OK;
    if (keys_dev) DLLM_TRY(kv_dequantize_one(ctx, kv, 0, keys_dev));
    if (values_dev) DLLM_TRY(kv_dequantize_one(ctx, kv, 1, values_dev));
    return DLLM_OK;
}

int32_t dllm_kv_dequantize(dllm_ctx *ctx, const dllm_kv *kv, float *keys, float *values) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kv, DLLM_ERR_NULL, "null kv entry");
    const size_t n = kv->L * kv->S * kv->H;
    if (n == 0) return DLLM_OK;
    void *dk, *dv;
    DLLM_TRY(stage_out_buf(ctx, 4, n * sizeof(float), &dk));
    DLLM_TRY(stage_out_buf(ctx, 5, n * sizeof(float), &dv));
    DLLM_TRY(dllm_kv_dequantize_dev(ctx, kv, keys ? (float *)dk : nullptr, values ? (float *)dv : nullptr));
    if (keys) DLLM_TRY(copy_out(ctx, keys, dk, n * sizeof(float)));
    if (values) DLLM_TRY(copy_out(ctx, values, dv, n * sizeof(float)));
    return sync(ctx);
}

int32_t dllm_kv_export(dllm_ctx *ctx, const dllm_kv *kv, uint8_t *key_codes, uint8_t *value_codes, float *key_scales,
                       float *key_zps, float *value_scales, float *value_zps) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kv, DLLM_ERR_NULL, "null kv entry");
    const size_t rows = kv->L * kv->S, n = rows * kv->H;
    uint8_t *codes[2] = {key_codes, value_codes};
    float *sc[2] = {key_scales, value_scales}, *zp[2] = {key_zps, value_zps};
    const size_t rb = kv_row_bytes(kv), rows_cap = kv->L * kv->cap;
    const bool contiguous = kv->cap == kv->S;
    const size_t n_l = contiguous ? kv->L : 1, loops = contiguous ? 1 : kv->L;      // layers per copy, copies
    for (int i = 0; i < 2; ++i) {
        if (codes[i] && n) {
            void *du = nullptr;
            if (kv->packed) DLLM_TRY(stage_out_buf(ctx, 1, n, &du));
            for (size_t l = 0; l < loops; ++l) {
                const uint8_t *src = kv->d_codes[i] + l * kv->cap * rb;
                const size_t cnt = n_l * kv->S * kv->H, off = l * kv->S * kv->H;
                if (kv->packed) {
                    DLLM_TRY(k_unpack(ctx, src, cnt, kv->bits, (uint8_t *)du + off));
                } else {
                    DLLM_TRY(copy_out(ctx, codes[i] + off, src, cnt));
                }
            }
            if (kv->packed) {
                DLLM_TRY(copy_out(ctx, codes[i], du, n));
                DLLM_TRY(sync(ctx));
            }
        }
        if (kv->scheme == DLLM_KV_TENSOR_B) {
            DLLM_TRY(copy_out(ctx, ctx->h_params, kv->d_params[i], 4 * sizeof(float)));
            DLLM_TRY(sync(ctx));
            if (sc[i]) sc[i][0] = ctx->h_params[0];
            if (zp[i]) zp[i][0] = ctx->h_params[1];
        } else if (kv->scheme == DLLM_KV_ROW_D) {
            for (size_t l = 0; l < loops; ++l) {
                const size_t cnt = n_l * kv->S, off = l * kv->S, r0 = l * kv->cap;
                if (sc[i] && cnt) DLLM_TRY(copy_out(ctx, sc[i] + off, kv->d_rows[i] + r0, cnt * sizeof(float)));
                if (zp[i] && cnt) DLLM_TRY(copy_out(ctx, zp[i] + off, kv->d_rows[i] + rows_cap + r0, cnt * sizeof(float)));
            }
        } else {
            if (sc[i]) sc[i][0] = kv->c_scale;
            if (zp[i]) zp[i][0] = 0.0f;
        }
    }
    return sync(ctx);
}

size_t dllm_kv_memory_usage(const dllm_kv *kv) {
    if (!kv) return 0;
    const size_t len = kv->L * kv->S * kv->H;                 // quantized.keys.data.len()
    return 2 * ((len * (size_t)kv->bits + 7) / 8);           // lib.rs:284-285 for keys + values
}

}  // extern "C"

// ==========================================================================================
// Phase-aware KV cache entry, resident in HBM
// ==========================================================================================
// KVCacheEntry (diffuse-llm-rs/src/lib.rs:122-313): the f32 keys / values plus up to two quantized copies — one at the
// prefill precision, one at the decode precision — and the phase that says which copy get_keys / get_values decode.
// Everything lives on the device: `update` takes device tensors, the getters decode straight into the consumer's buffer,
// so the cached branch of the sampling loop (lib.rs:885-921) moves nothing over PCIe.
// Layout: f32 rows of layer l at [l * capacity, l * capacity + seq) (so tokens can be appended without moving the others);
// the quantized copies are dllm_kv entries (per tensor: contiguous; per token / fixed scale: the same capacity layout).
struct dllm_kvcache {
    size_t L = 0, H = 0, cap = 0, S = 0;
    int scheme = DLLM_KV_TENSOR_B;
    int bits[2] = {8, 4};              // prefill, decode
    bool is_prefill = true;            // lib.rs:166 "Start in prefill phase by default"
    float *d_f32[2] = {nullptr, nullptr};
    dllm_kv *q[2] = {nullptr, nullptr};
    int device = 0;
};

// the quantized copy `which` (0 prefill, 1 decode) rebuilt from dense device tensors [L, S, H]
static int32_t kvc_requantize(dllm_ctx *ctx, dllm_kvcache *kc, int which, const float *k_dense, const float *v_dense) {
    if (kc->bits[which] <= 0) return DLLM_OK;
    if (!kc->q[which]) {
        // room for the whole capacity; a per-tensor entry is kept contiguous at its current length (cap == S)
        DLLM_TRY(kv_alloc(ctx, kc->L, kc->cap, kc->cap, kc->H, (uint8_t)kc->bits[which], kc->scheme, &kc->q[which]));
        kc->q[which]->S = 0;
    }
    dllm_kv *kv = kc->q[which];
    kv->S = kc->S;
    if (kc->scheme == DLLM_KV_TENSOR_B) kv->cap = kc->S;
    if (kc->L * kc->S * kc->H == 0) return DLLM_OK;
    DLLM_TRY(kv_quantize_one(ctx, kv, 0, k_dense));
    return kv_quantize_one(ctx, kv, 1, v_dense);
}

// dense [L, S, H] <-> the capacity layout
static int32_t kvc_copy_rows(dllm_ctx *ctx, float *dst, size_t dst_layer_rows, const float *src, size_t src_layer_rows,
                             size_t rows, size_t L, size_t H) {
    if (rows * L * H == 0) return DLLM_OK;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, dst_layer_rows * H * sizeof(float), src, src_layer_rows * H * sizeof(float),
                                    rows * H * sizeof(float), L, cudaMemcpyDeviceToDevice, ctx->stream));
    return DLLM_OK;
}

// rebuild copy `which` from the cache's own f32 tensors (transition_phase, lib.rs:228-235)
static int32_t kvc_requantize_from_store(dllm_ctx *ctx, dllm_kvcache *kc, int which) {
    if (kc->bits[which] <= 0) return DLLM_OK;
    if (kc->scheme == DLLM_KV_TENSOR_B && kc->cap != kc->S && kc->S > 0) {
        // one scale per tensor needs the tensor dense: gather the rows first
        const size_t n = kc->L * kc->S * kc->H;
        DLLM_TRY(ensure_buf(ctx, ctx->ws[4], n * sizeof(float)));
        DLLM_TRY(ensure_buf(ctx, ctx->ws[5], n * sizeof(float)));
        DLLM_TRY(kvc_copy_rows(ctx, (float *)ctx->ws[4].p, kc->S, kc->d_f32[0], kc->cap, kc->S, kc->L, kc->H));
        DLLM_TRY(kvc_copy_rows(ctx, (float *)ctx->ws[5].p, kc->S, kc->d_f32[1], kc->cap, kc->S, kc->L, kc->H));
        return kvc_requantize(ctx, kc, which, (const float *)ctx->ws[4].p, (const float *)ctx->ws[5].p);
    }
    if (kc->scheme == DLLM_KV_TENSOR_B) return kvc_requantize(ctx, kc, which, kc->d_f32[0], kc->d_f32[1]);
    if (!kc->q[which]) {
        DLLM_TRY(kv_alloc(ctx, kc->L, 0, kc->cap, kc->H, (uint8_t)kc->bits[which], kc->scheme, &kc->q[which]));
    }
    dllm_kv *kv = kc->q[which];
    kv->S = kc->S;
    if (kc->L * kc->S * kc->H == 0) return DLLM_OK;
    DLLM_TRY(kv_quantize_rows(ctx, kv, 0, kc->d_f32[0], 0, kc->S, kc->cap));
    return kv_quantize_rows(ctx, kv, 1, kc->d_f32[1], 0, kc->S, kc->cap);
}

extern "C" {

void dllm_kvcache_destroy(dllm_kvcache *kc) {
    if (!kc) return;
    cudaSetDevice(kc->device);
    for (int i = 0; i < 2; ++i) {
        if (kc->d_f32[i]) cudaFree(kc->d_f32[i]);
        if (kc->q[i]) dllm_kv_destroy(kc->q[i]);
    }
    delete kc;
}

int32_t dllm_kvcache_create(dllm_ctx *ctx, size_t layers, size_t hidden, size_t capacity, uint8_t prefill_bits, uint8_t decode_bits,
                            int32_t scheme, dllm_kvcache **out) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, out, DLLM_ERR_NULL, "null out pointer");
    *out = nullptr;
    ARG_CHECK(ctx, scheme >= DLLM_KV_TENSOR_B && scheme <= DLLM_KV_FIXED_C, DLLM_ERR_INVALID_PARAMS, "unknown scheme %d", scheme);
    ARG_CHECK(ctx, prefill_bits <= 8 && decode_bits <= 8, DLLM_ERR_INVALID_PARAMS, "Bits must be between 1 and 8 (0 = no quantized copy)");
    dllm_kvcache *kc = new (std::nothrow) dllm_kvcache();
    if (!kc) return DLLM_ERR_OOM;
    kc->L = layers; kc->H = hidden; kc->cap = capacity; kc->scheme = scheme; kc->device = ctx->device;
    kc->bits[0] = prefill_bits; kc->bits[1] = decode_bits;
    const size_t bytes = layers * capacity * hidden * sizeof(float);
    for (int i = 0; i < 2; ++i)
        if (cudaMalloc(&kc->d_f32[i], bytes ? bytes : 16) != cudaSuccess) {
            cudaGetLastError();
            dllm_kvcache_destroy(kc);
            DLLM_FAIL(ctx, DLLM_ERR_OOM, "cudaMalloc failed for the KV cache (%zu bytes per tensor)", bytes);
        }
    *out = kc;
    return DLLM_OK;
}

// KVCacheEntry::update (lib.rs:246-276): the tensors are replaced and BOTH quantized copies are rebuilt from them
int32_t dllm_kvcache_update_dev(dllm_ctx *ctx, dllm_kvcache *kc, const float *keys_dev, const float *values_dev, size_t seq) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kc, DLLM_ERR_NULL, "null cache entry");
    ARG_CHECK(ctx, seq <= kc->cap, DLLM_ERR_INDEX, "%zu tokens > capacity %zu", seq, kc->cap);
    if (!keys_dev && !values_dev && seq == kc->S) {
        // the model handed the cache's own tensors back (SimpleDiffusionModel::update_kv_cache, lib.rs:826-835): the copies are
        // rebuilt from them — which is what re-creates a decode copy dropped by a precision change
        for (int which = 0; which < 2; ++which) DLLM_TRY(kvc_requantize_from_store(ctx, kc, which));
        return DLLM_OK;
    }
    ARG_CHECK(ctx, kc->L * seq * kc->H == 0 || (keys_dev && values_dev), DLLM_ERR_NULL, "null device pointer");
    kc->S = seq;
    DLLM_TRY(kvc_copy_rows(ctx, kc->d_f32[0], kc->cap, keys_dev, seq, seq, kc->L, kc->H));
    DLLM_TRY(kvc_copy_rows(ctx, kc->d_f32[1], kc->cap, values_dev, seq, seq, kc->L, kc->H));
    for (int which = 0; which < 2; ++which) {
        if (kc->bits[which] <= 0) continue;
        if (kc->scheme == DLLM_KV_TENSOR_B) {
            DLLM_TRY(kvc_requantize(ctx, kc, which, keys_dev, values_dev));
        } else {
            if (!kc->q[which]) DLLM_TRY(kv_alloc(ctx, kc->L, 0, kc->cap, kc->H, (uint8_t)kc->bits[which], kc->scheme, &kc->q[which]));
            kc->q[which]->S = seq;
            if (kc->L * seq * kc->H != 0) {
                DLLM_TRY(kv_quantize_rows(ctx, kc->q[which], 0, keys_dev, 0, seq));
                DLLM_TRY(kv_quantize_rows(ctx, kc->q[which], 1, values_dev, 0, seq));
            }
        }
    }
    return DLLM_OK;
}

// The same update when only t_new tokens per layer are new (per-token / fixed-scale entries: a token's codes depend on that
// token alone, so the result is bit-identical to dllm_kvcache_update_dev with the concatenated tensors; a per-tensor entry
// re-quantizes everything, like the reference does, because its one scale changes)
int32_t dllm_kvcache_append_dev(dllm_ctx *ctx, dllm_kvcache *kc, const float *keys_new_dev, const float *values_new_dev, size_t t_new) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kc, DLLM_ERR_NULL, "null cache entry");
    ARG_CHECK(ctx, kc->S + t_new <= kc->cap, DLLM_ERR_INDEX, "KV cache is full: %zu + %zu tokens > capacity %zu", kc->S, t_new, kc->cap);
    if (t_new == 0) return DLLM_OK;
    ARG_CHECK(ctx, kc->L * kc->H == 0 || (keys_new_dev && values_new_dev), DLLM_ERR_NULL, "null device pointer");
    const size_t s0 = kc->S;
    DLLM_TRY(kvc_copy_rows(ctx, kc->d_f32[0] + s0 * kc->H, kc->cap, keys_new_dev, t_new, t_new, kc->L, kc->H));
    DLLM_TRY(kvc_copy_rows(ctx, kc->d_f32[1] + s0 * kc->H, kc->cap, values_new_dev, t_new, t_new, kc->L, kc->H));
    kc->S = s0 + t_new;
    for (int which = 0; which < 2; ++which) {
        if (kc->bits[which] <= 0) continue;
        if (kc->scheme == DLLM_KV_TENSOR_B || !kc->q[which]) {
            DLLM_TRY(kvc_requantize_from_store(ctx, kc, which));
        } else {
            dllm_kv *kv = kc->q[which];
            DLLM_TRY(kv_quantize_rows(ctx, kv, 0, keys_new_dev, s0, t_new));
            DLLM_TRY(kv_quantize_rows(ctx, kv, 1, values_new_dev, s0, t_new));
            kv->S = kc->S;
        }
    }
    return DLLM_OK;
}

// transition_phase (lib.rs:220-238): entering the decode phase creates the decode copy if it is missing
int32_t dllm_kvcache_set_phase(dllm_ctx *ctx, dllm_kvcache *kc, int32_t is_prefill) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kc, DLLM_ERR_NULL, "null cache entry");
    const bool p = is_prefill != 0;
    if (kc->is_prefill == p) return DLLM_OK;
    kc->is_prefill = p;
    if (!p && kc->bits[1] > 0 && !kc->q[1]) DLLM_TRY(kvc_requantize_from_store(ctx, kc, 1));
    return DLLM_OK;
}

// progressive precision (lib.rs:899-903): a new decode width drops the decode copy; the next update re-creates it
int32_t dllm_kvcache_set_decode_bits(dllm_ctx *ctx, dllm_kvcache *kc, uint8_t bits) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kc, DLLM_ERR_NULL, "null cache entry");
    ARG_CHECK(ctx, bits <= 8, DLLM_ERR_INVALID_PARAMS, "Bits must be between 1 and 8 (0 = no quantized copy)");
    if ((int)bits == kc->bits[1]) return DLLM_OK;
    kc->bits[1] = bits;
    if (kc->q[1]) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // its buffers may still be read by enqueued work
        dllm_kv_destroy(kc->q[1]);
        kc->q[1] = nullptr;
    }
    return DLLM_OK;
}

// get_keys / get_values (lib.rs:176-205): the active phase's copy decoded into the consumer's dense [L, S, H] buffers
// (either may be NULL); without a quantized copy, the f32 tensors themselves
int32_t dllm_kvcache_get_dev(dllm_ctx *ctx, const dllm_kvcache *kc, float *keys_out_dev, float *values_out_dev) {
    CTX_CHECK(ctx);
    ARG_CHECK(ctx, kc, DLLM_ERR_NULL, "null cache entry");
    if (kc->L * kc->S * kc->H == 0) return DLLM_OK;
    const dllm_kv *q = kc->is_prefill ? kc->q[0] : kc->q[1];
    if (q) return dllm_kv_dequantize_dev(ctx, q, keys_out_dev, values_out_dev);
    if (keys_out_dev) DLLM_TRY(kvc_copy_rows(ctx, keys_out_dev, kc->S, kc->d_f32[0], kc->cap, kc->S, kc->L, kc->H));
    if (values_out_dev) DLLM_TRY(kvc_copy_rows(ctx, values_out_dev, kc->S, kc->d_f32[1], kc->cap, kc->S, kc->L, kc->H));
    return DLLM_OK;
}

int32_t dllm_kvcache_info(const dllm_kvcache *kc, size_t *seq_len, int32_t *is_prefill, uint8_t *current_bits, size_t *memory_usage) {
    if (!kc) return DLLM_ERR_NULL;
    if (seq_len) *seq_len = kc->S;
    if (is_prefill) *is_prefill = kc->is_prefill ? 1 : 0;
    if (current_bits) *current_bits = (uint8_t)(kc->is_prefill ? kc->bits[0] : kc->bits[1]);      // lib.rs:211-217
    if (memory_usage) {
        size_t total = 0;                                                                       // lib.rs:279-302
        for (int i = 0; i < 2; ++i) if (kc->q[i]) total += dllm_kv_memory_usage(kc->q[i]);
        *memory_usage = total ? total : 2 * kc->L * kc->S * kc->H * sizeof(float);
    }
    return DLLM_OK;
}

// the quantized copy of one phase (NULL if absent) for dllm_kv_export / parity checks; owned by the cache entry
const dllm_kv *dllm_kvcache_copy(const dllm_kvcache *kc, int32_t prefill) {
    return kc ? kc->q[prefill ? 0 : 1] : nullptr;
}

}  // extern "C"
