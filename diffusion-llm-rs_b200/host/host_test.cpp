// host_test.cpp — exercises the C++ host mirror (dllm.hpp) the way the reference's own unit tests
// exercise the Rust API: same inputs, same assertions, plus the exact codes the reference's source
// produces for them (tests/golden/reference_kat.json).  Needs a B200; exits non-zero on any failure.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "dllm.hpp"

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

static bool feq(float a, float b) { return std::memcmp(&a, &b, 4) == 0; }

int main() {
    using namespace dllm;
    try {
        Context ctx(0);
        // diffuse-llm-rs/src/quantization.rs:254-265  test_quantized_tensor
        {
            auto [q, scale, zp] = quantization::quantize_tensor(ctx, {1.f, 2.f, 3.f, 4.f}, 4);
            CHECK((q == std::vector<uint8_t>{5, 10, 15, 15}));
            CHECK(feq(scale, 0.2f) && zp == 0.f);
            quantization::QuantizedTensor qt{q, {2, 2}, scale, zp, 4};
            CHECK(qt.dequantize(ctx).size() == 4);
            CHECK(qt.compression_ratio() > 4.0f && qt.compression_ratio() == 8.0f);
        }
        // quantization.rs:242-252  test_quantization (codes as the reference source computes them)
        {
            auto [q, scale, zp] = quantization::quantize_tensor(ctx, {1.f, 2.f, 3.f, 4.f, 5.f}, 4);
            CHECK((q == std::vector<uint8_t>{4, 7, 11, 15, 15}));
            CHECK(feq(scale, 0.26666668f) && zp == 0.f);
        }
        // bits outside 1..=8: the reference asserts (quantization.rs:39)
        {
            bool threw = false;
            try { quantization::quantize_tensor(ctx, {1.f}, 9); } catch (const Error &e) { threw = e.code == DLLM_ERR_INVALID_PARAMS; }
            CHECK(threw);
        }
        // quantization/src/lib.rs:61-79  roundtrip Int8 through quant_utils
        {
            std::vector<float> data{-1.f, 0.f, 1.f, 2.f, 3.f, 4.f};
            auto t = quant::quant_utils::quantize(ctx, data, {2, 3}, quant::QuantizationType::Int8, false);
            CHECK((t.data == std::vector<uint8_t>{0, 0, 1, 2, 3, 4}));
            CHECK(t.params.scale == 1.0f && t.params.zero_point == 0);
            auto d = quant::quant_utils::dequantize(ctx, t);
            CHECK(d.size() == 6 && d[0] == 0.f && d[5] == 4.f);
        }
        // quantization/src/calibrate.rs:123-132
        {
            auto p = quant::calibrate_params(1.0f, 6.0f, 6, 8, false);
            CHECK(feq(p.scale, 0.019607844f) && p.zero_point == -51);
            bool threw = false;
            try { quant::calibrate_params(0.f, 1.f, 0, 8, false); } catch (const Error &e) { threw = e.code == DLLM_ERR_CALIBRATION_REQUIRED; }
            CHECK(threw);
        }
        // diffusion_prefill/src/prefill_kv.rs:147-160
        {
            kvquant::KVCache cache(ctx);
            std::vector<float> v{0.1f, 0.5f, 1.0f, 0.0f};
            auto c = cache.compress_vector("test", v, 4);
            CHECK((c.data == std::vector<uint8_t>{1, 7, 14, 0}));
            CHECK(feq(c.quant_scale, 0.06666667f) && c.quant_zero_point == 0.f);
            auto d = cache.decompress_vector(c);
            for (size_t i = 0; i < v.size(); ++i) CHECK(std::fabs(v[i] - d[i]) < 0.1f);
        }
        // diffusion_prefill/src/fusion_ann.rs:144-165
        {
            kvquant::KVCache cache(ctx);
            std::vector<float> rows{0.1f, 0.2f, 0.3f, 0.4f, 0.5f, 0.6f, 0.7f, 0.8f, 0.8f, 0.7f, 0.6f, 0.5f, 0.4f, 0.3f, 0.2f, 0.1f};
            auto out = cache.compress_batch(rows, 2, 8, {4, 8});
            CHECK(out.size() == 2 && out[0].bits == 4 && out[1].bits == 8);
        }
        // BitQuantizer of prefill-kvquant-rs (fixed scale 1/(2^bits-1))
        {
            kvquant::BitQuantizer bq(ctx, dllm_bitquantizer_scale(4), 0.0f);
            auto q = bq.quantize({0.0f, 0.5f, 1.0f, 2.0f, -1.0f}, 4);
            CHECK((q == std::vector<uint8_t>{0, 7, 14, 15, 0}));   // 1.0 / (1/15 as f32) = 14.999999 -> truncates to 14
        }
        // the layer interface: one quantized linear (SimpleDiffusionModel op) + a 3-step sample loop
        {
            const size_t H = 128, batch = 2, seq = 4, feat = H * seq;
            std::vector<float> w(H * H);
            for (size_t i = 0; i < w.size(); ++i) w[i] = 0.02f * std::sin(0.37f * (float)i);
            auto layer = std::make_shared<diffuse_llm::QWeight>(ctx, w, H, H, 8, 128);
            diffuse_llm::QuantizedDiffusionModel model(ctx, {layer}, H, 50, DLLM_BETA_LINEAR, 1e-4f, 0.02f, DLLM_PATH_SIMT);
            std::vector<float> x(batch * feat);
            for (size_t i = 0; i < x.size(); ++i) x[i] = std::cos(0.11f * (float)i);
            auto y = model.forward(x, {0, 0}, batch, feat);
            CHECK(y.size() == x.size());
            // against the unquantized layer in double: 8-bit error stays small
            double maxerr = 0;
            for (size_t tok = 0; tok < batch * seq; ++tok)
                for (size_t n = 0; n < H; ++n) {
                    double acc = 0;
                    for (size_t k = 0; k < H; ++k) acc += (double)x[tok * H + k] * (double)w[k * H + n];
                    maxerr = std::fmax(maxerr, std::fabs(acc - (double)y[tok * H + n]));
                }
            CHECK(maxerr < 5e-3);
            diffuse_llm::DiffuseLLM llm(ctx);
            std::vector<float> noises(3 * x.size(), 0.25f);
            auto out = llm.sample(model, x, noises, batch, feat, 3);
            CHECK(out.size() == x.size());
            for (float v : out) CHECK(std::isfinite(v));
            std::printf("launches=%llu\n", (unsigned long long)ctx.launches());
        }
        // the int8 denoise mode through the same interface: per-tensor codes (group 0, the reference's scheme), DLLM_PATH_I8
        {
            const size_t H = 256, batch = 2, seq = 8, feat = H * seq;
            std::vector<float> w(H * H);
            for (size_t i = 0; i < w.size(); ++i) w[i] = 0.05f * std::sin(0.37f * (float)i);
            auto layer = std::make_shared<diffuse_llm::QWeight>(ctx, w, H, H, 8, 0);
            diffuse_llm::QuantizedDiffusionModel model(ctx, {layer}, H, 50, DLLM_BETA_LINEAR, 1e-4f, 0.02f, DLLM_PATH_I8);
            std::vector<float> x(batch * feat);
            for (size_t i = 0; i < x.size(); ++i) x[i] = std::cos(0.11f * (float)i);
            auto y = model.forward(x, {0, 0}, batch, feat);
            double num = 0, den = 0;
            for (size_t tok = 0; tok < batch * seq; ++tok)
                for (size_t n = 0; n < H; ++n) {
                    double acc = 0;
                    for (size_t k = 0; k < H; ++k) acc += (double)x[tok * H + k] * (double)w[k * H + n];
                    num += (acc - (double)y[tok * H + n]) * (acc - (double)y[tok * H + n]);
                    den += acc * acc;
                }
            CHECK(std::sqrt(num / den) < 2e-2);         // 8-bit weights + int8 activations: <= 1e-2 each (stated bound)
        }
        // diffuse-llm-rs/src/lib.rs:122-313  KVCacheEntry, resident in HBM: phases, progressive decode width, accounting
        {
            const size_t L = 2, S = 4, H = 64;
            std::vector<float> k(L * S * H), v(L * S * H);
            for (size_t i = 0; i < k.size(); ++i) { k[i] = std::sin(0.37f * (float)i); v[i] = 2.f * std::cos(0.11f * (float)i); }
            diffuse_llm::KVCacheEntry e(ctx, L, H, 16, 8, 4);
            CHECK(e.is_empty() && e.get_current_quant_bits() == 8);
            e.update(k, v, S);
            CHECK(e.len() == S);
            CHECK(e.memory_usage() == 2 * ((k.size() * 8 + 7) / 8) + 2 * ((k.size() * 4 + 7) / 8));   // :279-302
            // prefill phase decodes the 8-bit copy: equal to QuantizedKVCacheEntry::new(keys, values, 8).dequantize_keys()
            quantization::QuantizedKVCacheEntry q8(ctx, k, v, L, S, H, 8);
            auto k8 = e.get_keys(), r8 = q8.dequantize_keys();
            CHECK(k8.size() == r8.size() && std::memcmp(k8.data(), r8.data(), k8.size() * 4) == 0);
            e.set_phase(false);
            CHECK(e.get_current_quant_bits() == 4);
            quantization::QuantizedKVCacheEntry q4(ctx, k, v, L, S, H, 4);
            auto v4 = e.get_values(), rv4 = q4.dequantize_values();
            CHECK(std::memcmp(v4.data(), rv4.data(), v4.size() * 4) == 0);
            e.set_decode_bits(2);                       // :899-903: the decode copy is dropped; the getters return the f32 tensors
            auto kf = e.get_keys();
            CHECK(std::memcmp(kf.data(), k.data(), k.size() * 4) == 0);
            e.update(k, v, S);                          // ... and re-created at the new width by the next update
            quantization::QuantizedKVCacheEntry q2(ctx, k, v, L, S, H, 2);
            auto k2 = e.get_keys(), r2 = q2.dequantize_keys();
            CHECK(std::memcmp(k2.data(), r2.data(), k2.size() * 4) == 0);
        }
    } catch (const dllm::Error &e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "HOST_TEST_FAILED (%d)\n" : "HOST_TEST_OK\n", failures);
    return failures ? 1 : 0;
}
