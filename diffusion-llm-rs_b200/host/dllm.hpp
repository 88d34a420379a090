// dllm.hpp — C++17 host-side mirror of the reference's Rust interfaces for the hot path, written
// above the C ABI (include/dllm_b200.h).  The reference is compiled Rust and no Rust toolchain
// exists in this image, so this header plays the role of the `dllm-b200` wrapper crate shown in
// same names, argument meaning and error behaviour, one-to-one.
//
//   dllm::quantization::{quantize_tensor, dequantize_tensor, QuantizedTensor, QuantizedKVCacheEntry}
//        <- diffuse_llm_rs::quantization            (diffuse-llm-rs/src/quantization.rs)
//   dllm::quant::{QuantizationType, QuantizationParams, QuantizedTensor, Quantizer, DefaultQuantizer,
//                 quant_utils::{quantize, dequantize}, dequant_matmul}
//        <- the `quantization` crate                (quantization/src/{quantize,types}.rs)
//   dllm::kvquant::{Quantizer, BitQuantizer, CompressedVector, KVCache}
//        <- prefill_kvquant_rs::kvquant, diffusion_prefill::prefill_kv
//   dllm::diffuse_llm::{DiffusionModel, QuantizedDiffusionModel, DiffuseLLM}
//        <- diffuse_llm_rs::diffuse_llm             (diffuse-llm-rs/src/lib.rs:748-955)
//
// Errors: Rust `Result::Err(QuantizationError::X)` and panics both surface as dllm::Error (with the
// DLLM_ERR_* code); there is no CPU fallback — constructing a Context without an sm_100 GPU throws.
#pragma once

#include <cstdint>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/dllm_b200.h"

namespace dllm {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error("[dllm status " + std::to_string(c) + "] " + m), code(c) {}
};

// One device + one stream.  `Send + Sync` in the reference == the mutex here.
class Context {
public:
    explicit Context(int device = 0) {
        int32_t rc = dllm_ctx_create(device, &h_);
        if (rc != DLLM_OK) throw Error(rc, "dllm_ctx_create failed (no sm_100 CUDA device?) - there is no CPU fallback");
    }
    ~Context() { dllm_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    dllm_ctx *raw() const { return h_; }
    std::mutex &mutex() { return mu_; }
    void check(int32_t rc) const {
        if (rc != DLLM_OK) throw Error(rc, dllm_last_error(h_));
    }
    uint64_t launches() const { return dllm_launch_count(h_); }

private:
    dllm_ctx *h_ = nullptr;
    std::mutex mu_;
};

// ------------------------------------------------------------------------------------------------
namespace quantization {  // diffuse-llm-rs/src/quantization.rs

// quantize_tensor(data: &[f32], bits: u8) -> (Vec<u8>, f32, f32)   :38-68
inline std::tuple<std::vector<uint8_t>, float, float> quantize_tensor(Context &ctx, const std::vector<float> &data, uint8_t bits) {
    std::lock_guard<std::mutex> lk(ctx.mutex());
    std::vector<uint8_t> codes(data.size());
    float scale = 0.f, zp = 0.f;
    ctx.check(dllm_quantize_tensor(ctx.raw(), data.data(), data.size(), bits, codes.data(), &scale, &zp));
    return {std::move(codes), scale, zp};
}

// dequantize_tensor(data: &[u8], scale, zero_point) -> Vec<f32>   :81-85
inline std::vector<float> dequantize_tensor(Context &ctx, const std::vector<uint8_t> &data, float scale, float zero_point) {
    std::lock_guard<std::mutex> lk(ctx.mutex());
    std::vector<float> out(data.size());
    ctx.check(dllm_dequantize_tensor(ctx.raw(), data.data(), data.size(), scale, zero_point, out.data()));
    return out;
}

struct QuantizedTensor {  // :89-125
    std::vector<uint8_t> data;
    std::vector<size_t> shape;
    float scale = 0.f, zero_point = 0.f;
    uint8_t bits = 8;
    std::vector<float> dequantize(Context &ctx) const { return dequantize_tensor(ctx, data, scale, zero_point); }
    float compression_ratio() const {
        size_t numel = 1;
        for (size_t d : shape) numel *= d;
        return dllm_compression_ratio(numel, data.size(), bits);
    }
};

// QuantizedKVCacheEntry::new(keys, values, bits) over [layers, seq, hidden]   :129-176
class QuantizedKVCacheEntry {
public:
    QuantizedKVCacheEntry(Context &ctx, const std::vector<float> &keys, const std::vector<float> &values, size_t layers,
                          size_t seq, size_t hidden, uint8_t bits, int32_t scheme = DLLM_KV_TENSOR_B)
        : seq_len(seq), ctx_(ctx), layers_(layers), hidden_(hidden) {
        if (keys.size() != layers * seq * hidden || values.size() != keys.size()) throw Error(DLLM_ERR_SHAPE, "shape mismatch");
        std::lock_guard<std::mutex> lk(ctx.mutex());
        ctx.check(dllm_kv_quantize(ctx.raw(), keys.data(), values.data(), layers, seq, hidden, bits, scheme, &h_));
    }
    ~QuantizedKVCacheEntry() { dllm_kv_destroy(h_); }
    std::vector<float> dequantize_keys() { return deq(true); }
    std::vector<float> dequantize_values() { return deq(false); }
    size_t memory_usage() const { return dllm_kv_memory_usage(h_); }
    size_t seq_len;

private:
    std::vector<float> deq(bool keys) {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<float> out(layers_ * seq_len * hidden_);
        ctx_.check(dllm_kv_dequantize(ctx_.raw(), h_, keys ? out.data() : nullptr, keys ? nullptr : out.data()));
        return out;
    }
    Context &ctx_;
    dllm_kv *h_ = nullptr;
    size_t layers_, hidden_;
};

}  // namespace quantization

// ------------------------------------------------------------------------------------------------
namespace quant {  // the `quantization` crate

enum class QuantizationType : int32_t { Int8 = DLLM_QT_INT8, Int4 = DLLM_QT_INT4, Binary = DLLM_QT_BINARY, Float8 = DLLM_QT_FLOAT8 };
inline uint8_t bits_of(QuantizationType t) {  // quantize.rs:70-77
    switch (t) { case QuantizationType::Int8: return 8; case QuantizationType::Int4: return 4;
                 case QuantizationType::Binary: return 1; default: return 8; }
}

struct QuantizationParams {  // types.rs:21-40
    uint8_t bits = 8;
    float scale = 1.0f;
    int32_t zero_point = 0;
    bool symmetric = true;
    std::optional<size_t> axis;
};

struct QuantizedTensor {  // types.rs:42-81
    std::vector<uint8_t> data;
    std::vector<size_t> shape;
    QuantizationParams params;
};

struct Quantizer {  // trait Quantizer, quantize.rs:81-90
    virtual ~Quantizer() = default;
    virtual QuantizedTensor quantize(const std::vector<float> &data, const std::vector<size_t> &shape, QuantizationType qtype) = 0;
    virtual std::vector<float> dequantize(const QuantizedTensor &t) = 0;
    virtual const QuantizationParams &get_params() const = 0;
};

class DefaultQuantizer : public Quantizer {  // quantize.rs:93-185
public:
    DefaultQuantizer(Context &ctx, uint8_t bits, bool symmetric, std::optional<size_t> axis = std::nullopt) : ctx_(ctx) {
        params_.bits = bits; params_.scale = 1.0f; params_.zero_point = 0; params_.symmetric = symmetric; params_.axis = axis;  // :98-108
    }
    DefaultQuantizer(Context &ctx, const QuantizationParams &p) : ctx_(ctx), params_(p) {}   // what calibration feeds
    QuantizedTensor quantize(const std::vector<float> &data, const std::vector<size_t> &shape, QuantizationType qtype) override {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        QuantizedTensor t{std::vector<uint8_t>(data.size()), shape, params_};
        ctx_.check(dllm_quantize_a(ctx_.raw(), data.data(), data.size(), (int32_t)qtype, params_.scale, params_.zero_point, t.data.data()));
        return t;
    }
    std::vector<float> dequantize(const QuantizedTensor &t) override {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<float> out(t.data.size());
        ctx_.check(dllm_dequantize_a(ctx_.raw(), t.data.data(), t.data.size(), t.params.scale, t.params.zero_point, out.data()));
        return out;
    }
    const QuantizationParams &get_params() const override { return params_; }

private:
    Context &ctx_;
    QuantizationParams params_;
};

namespace quant_utils {  // quantize.rs:188-215
inline QuantizedTensor quantize(Context &ctx, const std::vector<float> &data, const std::vector<size_t> &shape,
                                QuantizationType qtype, bool symmetric, std::optional<size_t> axis = std::nullopt) {
    return DefaultQuantizer(ctx, bits_of(qtype), symmetric, axis).quantize(data, shape, qtype);
}
inline std::vector<float> dequantize(Context &ctx, const QuantizedTensor &t) {
    return DefaultQuantizer(ctx, t.params.bits, t.params.symmetric, t.params.axis).dequantize(t);
}
}  // namespace quant_utils

// CalibrationData::compute_params, calibrate.rs:72-110
inline QuantizationParams calibrate_params(float min, float max, size_t total_samples, uint8_t bits, bool symmetric) {
    QuantizationParams p;
    p.bits = bits; p.symmetric = symmetric;
    int32_t rc = dllm_calibrate_params(min, max, total_samples, bits, symmetric ? 1 : 0, &p.scale, &p.zero_point);
    if (rc != DLLM_OK) throw Error(rc, rc == DLLM_ERR_CALIBRATION_REQUIRED ? "Calibration data is required" : "invalid parameters");
    return p;
}

// extension named by BASELINE.json north_star: y[M,N] = x[M,K] . dequant(codes[K,N]) + bias
inline std::vector<float> dequant_matmul(Context &ctx, const std::vector<uint8_t> &codes, const std::vector<float> &scales,
                                         const std::vector<float> &zero_points, size_t K, size_t N, uint8_t bits, size_t group,
                                         const std::vector<float> &x, size_t M, int32_t path = DLLM_PATH_AUTO) {
    std::lock_guard<std::mutex> lk(ctx.mutex());
    std::vector<float> y(M * N);
    ctx.check(dllm_dequant_matmul(ctx.raw(), codes.data(), scales.data(), zero_points.data(), K, N, bits, group, nullptr,
                                  x.data(), M, y.data(), path));
    return y;
}

}  // namespace quant

// ------------------------------------------------------------------------------------------------
namespace kvquant {  // prefill-kvquant-rs/lib.rs, diffusion_prefill/src/prefill_kv.rs

struct Quantizer {  // trait Quantizer: Send + Sync   lib.rs:29-32
    virtual ~Quantizer() = default;
    virtual std::vector<uint8_t> quantize(const std::vector<float> &input, uint8_t bits) = 0;
    virtual std::vector<float> dequantize(const std::vector<uint8_t> &input, uint8_t bits) = 0;
};

class BitQuantizer : public Quantizer {  // lib.rs:34-53 / prefill_kv.rs:48-67
public:
    BitQuantizer(Context &ctx, float scale, float zero_point) : ctx_(ctx), scale(scale), zero_point(zero_point) {}
    std::vector<uint8_t> quantize(const std::vector<float> &input, uint8_t bits) override {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<uint8_t> out(input.size());
        ctx_.check(dllm_quantize_c(ctx_.raw(), input.data(), input.size(), bits, scale, zero_point, out.data()));
        return out;
    }
    std::vector<float> dequantize(const std::vector<uint8_t> &input, uint8_t) override {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<float> out(input.size());
        ctx_.check(dllm_dequantize_cd(ctx_.raw(), input.data(), input.size(), scale, zero_point, out.data()));
        return out;
    }

private:
    Context &ctx_;
public:
    float scale, zero_point;
};

struct CompressedVector {  // prefill_kv.rs:25-33
    std::string id;
    std::vector<uint8_t> data;
    uint8_t bits = 0;
    std::vector<size_t> original_shape;
    float quant_scale = 0.f, quant_zero_point = 0.f;
};

// KVCache::compress_vector / decompress_vector (prefill_kv.rs:104-132) and FusionANN::quantize (fusion_ann.rs:53-63)
class KVCache {
public:
    explicit KVCache(Context &ctx) : ctx_(ctx) {}
    CompressedVector compress_vector(const std::string &id, const std::vector<float> &v, uint8_t bits) {
        return compress_batch(v, 1, v.size(), {bits}, id)[0];
    }
    std::vector<CompressedVector> compress_batch(const std::vector<float> &rows, size_t n_rows, size_t dim,
                                                 const std::vector<uint8_t> &bits, const std::string &id0 = "") {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<uint8_t> codes(n_rows * dim);
        std::vector<float> s(n_rows), z(n_rows);
        ctx_.check(dllm_quantize_d_rows(ctx_.raw(), rows.data(), n_rows, dim, bits.data(), bits.size(), codes.data(), s.data(), z.data()));
        std::vector<CompressedVector> out(n_rows);
        for (size_t r = 0; r < n_rows; ++r) {
            out[r].id = id0.empty() ? std::to_string(r) : id0;
            out[r].data.assign(codes.begin() + r * dim, codes.begin() + (r + 1) * dim);
            out[r].bits = bits[r % bits.size()];
            out[r].original_shape = {dim};
            out[r].quant_scale = s[r];
            out[r].quant_zero_point = z[r];
        }
        return out;
    }
    std::vector<float> decompress_vector(const CompressedVector &v) {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<float> out(v.data.size());
        ctx_.check(dllm_dequantize_d_rows(ctx_.raw(), v.data.data(), 1, v.data.size(), &v.quant_scale, &v.quant_zero_point, out.data()));
        return out;
    }

private:
    Context &ctx_;
};

}  // namespace kvquant

// ------------------------------------------------------------------------------------------------
namespace diffuse_llm {  // diffuse-llm-rs/src/lib.rs

struct DiffusionModel {  // trait DiffusionModel: Send + Sync   :748-772
    virtual ~DiffusionModel() = default;
    virtual std::vector<float> forward(const std::vector<float> &x, const std::vector<size_t> &t, size_t batch, size_t feat) = 0;
    virtual std::vector<float> forward_with_cache(const std::vector<float> &x, const std::vector<size_t> &t, size_t batch,
                                                  size_t feat, const std::vector<float> &keys, const std::vector<float> &values) = 0;
};

class QWeight {
public:
    QWeight(Context &ctx, const std::vector<float> &w, size_t K, size_t N, uint8_t bits, size_t group, const float *bias = nullptr) {
        std::lock_guard<std::mutex> lk(ctx.mutex());
        ctx.check(dllm_qweight_quantize(ctx.raw(), w.data(), K, N, bits, group, bias, &h_));
    }
    ~QWeight() { dllm_qweight_destroy(h_); }
    QWeight(const QWeight &) = delete;
    dllm_qweight *raw() const { return h_; }
    // y[M,N] = x[M,K] . dequant(W) + b   (lib.rs:806-813)
    std::vector<float> forward(Context &ctx, const std::vector<float> &x, size_t M, size_t N, int32_t path = DLLM_PATH_AUTO) const {
        std::vector<float> y(M * N);
        std::lock_guard<std::mutex> lk(ctx.mutex());
        ctx.check(dllm_qlinear_forward(ctx.raw(), h_, x.data(), M, y.data(), path));
        return y;
    }
    // exact: y[M,N] = sum_k xq[m,k] * (q[k,n] - zp) for a per-tensor quantized weight (group 0); no rounding anywhere
    std::vector<int32_t> forward_i8(Context &ctx, const std::vector<int8_t> &xq, size_t M, size_t N) const {
        std::vector<int32_t> y(M * N);
        std::lock_guard<std::mutex> lk(ctx.mutex());
        ctx.check(dllm_qlinear_forward_i8(ctx.raw(), h_, xq.data(), M, y.data()));
        return y;
    }

private:
    dllm_qweight *h_ = nullptr;
};

// A stack of quantized linears, each the reference's `x.dot(&W) + &b` (lib.rs:806-813)
class QuantizedDiffusionModel : public DiffusionModel {
public:
    QuantizedDiffusionModel(Context &ctx, std::vector<std::shared_ptr<QWeight>> layers, size_t hidden, size_t num_timesteps = 1000,
                            int32_t beta_kind = DLLM_BETA_LINEAR, float beta_start = 1e-4f, float beta_end = 0.02f,
                            int32_t path = DLLM_PATH_AUTO)
        : ctx_(ctx), layers_(std::move(layers)), path_(path) {
        std::vector<dllm_qweight *> raw;
        for (auto &l : layers_) raw.push_back(l->raw());
        std::lock_guard<std::mutex> lk(ctx.mutex());
        ctx.check(dllm_model_create(ctx.raw(), hidden, raw.data(), raw.size(), num_timesteps, beta_kind, beta_start, beta_end, &h_));
    }
    ~QuantizedDiffusionModel() override { dllm_model_destroy(h_); }
    std::vector<float> forward(const std::vector<float> &x, const std::vector<size_t> &t, size_t batch, size_t feat) override {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<float> out(x.size());
        ctx_.check(dllm_model_forward(ctx_.raw(), h_, x.data(), t.data(), batch, feat, out.data(), path_));
        return out;
    }
    std::vector<float> forward_with_cache(const std::vector<float> &x, const std::vector<size_t> &t, size_t batch, size_t feat,
                                          const std::vector<float> &, const std::vector<float> &) override {
        return forward(x, t, batch, feat);   // lib.rs:815-824: the cache is ignored
    }
    dllm_model *raw() const { return h_; }
    int32_t path() const { return path_; }

private:
    Context &ctx_;
    std::vector<std::shared_ptr<QWeight>> layers_;
    dllm_model *h_ = nullptr;
    int32_t path_;
};

// KVCacheEntry (lib.rs:122-313), resident in HBM: f32 keys / values [layers, seq, hidden] + a prefill- and a decode-precision
// quantized copy.  The host-vector methods stage through device memory; the *_dev methods take device pointers (what a GPU-resident
// sampling loop uses: the cached branch of DiffuseLLM::sample, :885-921, then moves nothing over PCIe).
class KVCacheEntry {
public:
    KVCacheEntry(Context &ctx, size_t layers, size_t hidden, size_t capacity, uint8_t prefill_bits, uint8_t decode_bits,
                 int32_t scheme = DLLM_KV_TENSOR_B)
        : ctx_(ctx), layers_(layers), hidden_(hidden) {
        std::lock_guard<std::mutex> lk(ctx.mutex());
        ctx.check(dllm_kvcache_create(ctx.raw(), layers, hidden, capacity, prefill_bits, decode_bits, scheme, &h_));
    }
    ~KVCacheEntry() { dllm_kvcache_destroy(h_); }
    KVCacheEntry(const KVCacheEntry &) = delete;
    // update(new_keys, new_values), :246-276: both quantized copies are rebuilt
    void update(const std::vector<float> &keys, const std::vector<float> &values, size_t seq) {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        const size_t bytes = layers_ * seq * hidden_ * sizeof(float);
        void *dk = nullptr, *dv = nullptr;
        ctx_.check(dllm_malloc(ctx_.raw(), bytes ? bytes : 4, &dk));
        ctx_.check(dllm_malloc(ctx_.raw(), bytes ? bytes : 4, &dv));
        int32_t rc = dllm_memcpy_h2d(ctx_.raw(), dk, keys.data(), bytes);
        if (rc == DLLM_OK) rc = dllm_memcpy_h2d(ctx_.raw(), dv, values.data(), bytes);
        if (rc == DLLM_OK) rc = dllm_kvcache_update_dev(ctx_.raw(), h_, (const float *)dk, (const float *)dv, seq);
        if (rc == DLLM_OK) rc = dllm_ctx_sync(ctx_.raw());
        dllm_free(ctx_.raw(), dk);
        dllm_free(ctx_.raw(), dv);
        ctx_.check(rc);
    }
    void update_dev(const float *keys_dev, const float *values_dev, size_t seq) {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        ctx_.check(dllm_kvcache_update_dev(ctx_.raw(), h_, keys_dev, values_dev, seq));
    }
    void append_dev(const float *keys_new_dev, const float *values_new_dev, size_t t_new) {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        ctx_.check(dllm_kvcache_append_dev(ctx_.raw(), h_, keys_new_dev, values_new_dev, t_new));
    }
    void get_dev(float *keys_out_dev, float *values_out_dev) const {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        ctx_.check(dllm_kvcache_get_dev(ctx_.raw(), h_, keys_out_dev, values_out_dev));
    }
    std::vector<float> get_keys() const { return fetch(true); }        // :176-189
    std::vector<float> get_values() const { return fetch(false); }     // :192-205
    void set_phase(bool is_prefill) { transition_phase(is_prefill); }  // :207-209
    void transition_phase(bool is_prefill) {                           // :220-238
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        ctx_.check(dllm_kvcache_set_phase(ctx_.raw(), h_, is_prefill ? 1 : 0));
    }
    void set_decode_bits(uint8_t bits) {                               // :899-903
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        ctx_.check(dllm_kvcache_set_decode_bits(ctx_.raw(), h_, bits));
    }
    uint8_t get_current_quant_bits() const { uint8_t b = 0; dllm_kvcache_info(h_, nullptr, nullptr, &b, nullptr); return b; }
    size_t memory_usage() const { size_t m = 0; dllm_kvcache_info(h_, nullptr, nullptr, nullptr, &m); return m; }
    size_t len() const { size_t s = 0; dllm_kvcache_info(h_, &s, nullptr, nullptr, nullptr); return s; }
    bool is_empty() const { return len() == 0; }
    dllm_kvcache *raw() const { return h_; }

private:
    std::vector<float> fetch(bool keys) const {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        const size_t n = layers_ * len_unlocked() * hidden_;
        std::vector<float> out(n);
        if (n == 0) return out;
        void *d = nullptr;
        ctx_.check(dllm_malloc(ctx_.raw(), n * sizeof(float), &d));
        int32_t rc = dllm_kvcache_get_dev(ctx_.raw(), h_, keys ? (float *)d : nullptr, keys ? nullptr : (float *)d);
        if (rc == DLLM_OK) rc = dllm_memcpy_d2h(ctx_.raw(), out.data(), d, n * sizeof(float));
        if (rc == DLLM_OK) rc = dllm_ctx_sync(ctx_.raw());
        dllm_free(ctx_.raw(), d);
        ctx_.check(rc);
        return out;
    }
    size_t len_unlocked() const { size_t s = 0; dllm_kvcache_info(h_, &s, nullptr, nullptr, nullptr); return s; }
    Context &ctx_;
    size_t layers_, hidden_;
    dllm_kvcache *h_ = nullptr;
};

// DiffuseLLM::sample without cache (lib.rs:853-927) and p_sample (:1152-1215), noise injected
class DiffuseLLM {
public:
    explicit DiffuseLLM(Context &ctx) : ctx_(ctx) {}
    std::vector<float> p_sample(QuantizedDiffusionModel &m, const std::vector<float> &x_t, const std::vector<size_t> &t,
                                const std::vector<float> &noise_pred, const std::vector<float> *noise, size_t batch, size_t feat,
                                bool guard_t0 = true) {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<float> out(x_t.size());
        ctx_.check(dllm_p_sample(ctx_.raw(), m.raw(), x_t.data(), noise_pred.data(), noise ? noise->data() : nullptr, t.data(),
                                 batch, feat, guard_t0 ? 1 : 0, out.data()));
        return out;
    }
    // add_noise (lib.rs:1100-1137): returns the noisy input; the noise is the caller's (the second element of the reference's pair)
    std::vector<float> add_noise(QuantizedDiffusionModel &m, const std::vector<float> &x_start, const std::vector<size_t> &t,
                                 const std::vector<float> &noise, size_t batch, size_t feat) {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<float> out(x_start.size());
        ctx_.check(dllm_add_noise(ctx_.raw(), m.raw(), x_start.data(), noise.data(), t.data(), batch, feat, out.data()));
        return out;
    }
    std::vector<float> sample(QuantizedDiffusionModel &m, const std::vector<float> &x0, const std::vector<float> &noises,
                              size_t batch, size_t feat, size_t num_steps, bool guard_t0 = true) {
        std::lock_guard<std::mutex> lk(ctx_.mutex());
        std::vector<float> out(x0.size());
        ctx_.check(dllm_sample(ctx_.raw(), m.raw(), x0.data(), noises.empty() ? nullptr : noises.data(), batch, feat, num_steps,
                               guard_t0 ? 1 : 0, m.path(), out.data()));
        return out;
    }

private:
    Context &ctx_;
};

}  // namespace diffuse_llm
}  // namespace dllm
