"""GPU parity (bit-exact) of the quantize / dequantize / pack / KV kernels against the oracle,
through the C ABI.  Small and ragged sizes against the oracle; full sizes through
size-independent properties and an independent torch restatement."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F = np.float32
KAT = {k["name"]: k for k in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.json")))}


@pytest.fixture(scope="module")
def ctx():
    import dllm_b200
    c = dllm_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def O():
    from oracle import pyoracle
    return pyoracle


def beq(a, b):
    """bit-for-bit f32 equality; NaNs must sit at the same places (their payload bits are not
    part of the reference's semantics: x86 and the GPU produce different default NaNs)."""
    a, b = np.asarray(a, F), np.asarray(b, F)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb]))


def edge_cases(rng):
    yield rng.standard_normal(1000).astype(F)
    yield (rng.standard_normal(4097) * 0.02).astype(F)
    yield rng.random(333).astype(F)
    yield np.array([], F)
    yield np.array([3.25], F)
    yield np.full(17, -2.5, F)
    yield np.array([0.0, -0.0, 0.0], F)
    yield np.array([1.0, np.nan, -1.0, 2.0, 0.5], F)
    yield np.array([np.inf, 1.0, -1.0], F)
    yield np.array([-np.inf, 1.0, np.inf], F)
    yield np.array([1e38, -1e38, 0.5], F)
    yield np.array([1e-45, 0.0, 2e-45], F)
    yield (rng.standard_normal(64) * 1e6).astype(F)
    yield rng.standard_normal((1 << 20) + 3).astype(F)


# ---------------------------------------------------------------- quantizer B
def test_b_reference_kat(ctx):
    for name in ("B_test_quantized_tensor", "B_test_quantization"):
        k = KAT[name]
        codes, s, z = ctx.quantize_tensor(k["input"], k["bits"])
        assert codes.tolist() == k["codes"]
        assert beq(s, F(k["scale"])) and beq(z, F(k["zero_point"]))
        assert beq(ctx.dequantize_tensor(codes, s, z), np.array(k["dequant"], F))


@pytest.mark.parametrize("bits", [1, 2, 3, 4, 5, 6, 7, 8])
def test_b_bit_exact(ctx, O, bits):
    rng = np.random.default_rng(100 + bits)
    for x in edge_cases(rng):
        c0, s0, z0 = O.quantize_tensor(x, bits)
        c1, s1, z1 = ctx.quantize_tensor(x, bits)
        assert beq(s0, s1) and beq(z0, z1), (x[:6], s0, s1, z0, z1)
        assert np.array_equal(c0, c1)
        assert beq(O.dequantize_tensor(c0, s0, z0), ctx.dequantize_tensor(c1, s1, z1))


def test_b_bad_bits_is_invalid_params(ctx):
    import dllm_b200
    for bad in (0, 9, 200):
        with pytest.raises(dllm_b200.InvalidParams):
            ctx.quantize_tensor([1.0, 2.0], bad)


def test_b_code_step_with_given_params(ctx, O):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(5001).astype(F)
    for bits, s, z in ((4, 0.37, 7.0), (8, 0.011, 128.0), (2, 1.5, 1.0)):
        assert np.array_equal(ctx.quantize_codes(x, bits, s, z), O.quantize_codes_b(x, bits, s, z))


@pytest.mark.parametrize("bits", [1, 2, 4, 8])
@pytest.mark.parametrize("n", [1, 3, 4, 5, 8, 33, 1023, 4096, 100003])
def test_b_packed_device_path(ctx, O, bits, n):
    """quantize on the device straight into packed codes; dequantize from packed."""
    rng = np.random.default_rng(n + bits)
    x = rng.standard_normal(n).astype(F)
    dx = ctx.malloc(n * 4 + 16)
    dc = ctx.malloc(n + 16)
    dp = ctx.malloc(16)
    dout = ctx.malloc(n * 4 + 16)
    ctx.h2d(dx, x)
    ctx.quantize_tensor_dev(dx, n, bits, True, dc, dp)
    ctx.dequantize_tensor_dev(dc, n, bits, True, dp, dout)
    ctx.sync()
    packed = ctx.d2h(dc, (O.packed_len(n, bits),), np.uint8)
    params = ctx.d2h(dp, (4,), F)
    codes, s, z = O.quantize_tensor(x, bits)
    assert beq(params[0], s) and beq(params[1], z)
    assert beq(params[2], x.min()) and beq(params[3], x.max())
    assert np.array_equal(packed, O.pack(codes, bits))
    assert beq(ctx.d2h(dout, (n,), F), O.dequantize_tensor(codes, s, z))
    for p in (dx, dc, dp, dout):
        ctx.free(p)


# ---------------------------------------------------------------- quantizer A
def test_a_reference_kat(ctx):
    import dllm_b200.quant as Q
    for name in ("A_roundtrip_int8", "A_example_basic"):
        k = KAT[name]
        t = Q.quant_utils.quantize(np.array(k["input"], F), Q.QuantizationType.Int8, False, None, ctx)
        assert t.data.tolist() == k["codes"]
        assert t.params.scale == 1.0 and t.params.zero_point == 0        # quantize.rs:98-108
        assert beq(Q.quant_utils.dequantize(t, ctx).ravel(), np.array(k["dequant"], F))


@pytest.mark.parametrize("qtype", [0, 1, 2, 3])
def test_a_bit_exact(ctx, O, qtype):
    rng = np.random.default_rng(50 + qtype)
    for x in edge_cases(rng):
        for scale, zp in ((1.0, 0), (0.05, 3), (0.5, -51)):
            xs = (x * F(10)).astype(F)
            c0 = O.quantize_a(xs, qtype, scale, zp)
            c1 = ctx.quantize_a(xs, qtype, scale, zp)
            assert np.array_equal(c0, c1)
            assert beq(O.dequantize_a(c0, scale, zp), ctx.dequantize_a(c1, scale, zp))


def test_calibration_on_gpu(ctx, O):
    import dllm_b200.quant as Q
    cal = Q.CalibrationData(10, False, ctx)
    cal.update(np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], F))           # calibrate.rs:123-132
    p = cal.compute_params(8, False)
    assert beq(p.scale, F(0.019607844)) and p.zero_point == -51
    assert cal.total_samples == 6 and sum(cal.histogram) == 6
    rng = np.random.default_rng(3)
    x = rng.standard_normal(100001).astype(F)
    mn, mx = ctx.minmax(x)
    assert beq(mn, x.min()) and beq(mx, x.max())
    with pytest.raises(Q.QuantizationError):
        Q.CalibrationData(10, False, ctx).compute_params(8, False)      # CalibrationRequired


# ---------------------------------------------------------------- quantizer C
@pytest.mark.parametrize("bits", [1, 2, 4, 6, 8, 16])
def test_c_bit_exact(ctx, O, bits):
    rng = np.random.default_rng(70 + bits)
    sc = O.bitquantizer_scale_c(bits)
    for x in edge_cases(rng):
        c0, c1 = O.quantize_c(x, bits, sc), ctx.quantize_c(x, bits, sc)
        assert np.array_equal(c0, c1)
        assert beq(O.dequantize_cd(c0, sc, 0.0), ctx.dequantize_cd(c1, sc, 0.0))
    x = rng.random(999).astype(F)
    assert np.array_equal(O.quantize_c(x, bits, 0.013, 0.25), ctx.quantize_c(x, bits, 0.013, 0.25))


def test_c_quantize_vectors(ctx, O):
    import dllm_b200
    import dllm_b200.kvquant as KQ
    rng = np.random.default_rng(5)
    emb = rng.random((7, 3, 8)).astype(F)
    for cfg, bits in (([4, 6, 8, 16], [4]), ([4, 6, 8, 16], [2, 4]), ([1, 2, 4, 8], [1, 2, 4]), ([4, 6, 8, 16], [6, 2, 4])):
        assert np.array_equal(ctx.kvquant_quantize_vectors(emb, cfg, bits), O.kvquant_quantize_vectors(emb, cfg, bits))
    with pytest.raises(dllm_b200.ReferencePanic):           # quantizers[4] of a 4-entry Vec, lib.rs:133
        ctx.kvquant_quantize_vectors(emb, [4, 6, 8, 16], [8])
    pk = KQ.PrefillKVQuant(KQ.SystemConfig(), ctx)
    toks = [KQ.TokenizedVector(str(i), [i], emb[i]) for i in range(7)]
    out = pk.quantize_vectors(toks, [4, 2])
    exp = O.kvquant_quantize_vectors(emb, [4, 6, 8, 16], [4, 2])
    for i, cv in enumerate(out):
        assert cv.bits == [4, 2][i % 2] and cv.original_shape == [3, 8]
        assert np.array_equal(cv.data, exp[i].ravel())


# ---------------------------------------------------------------- quantizer D
def test_d_reference_kat(ctx):
    import dllm_b200.kvquant as KQ
    k = KAT["D_prefill_kv_test_quantization"]
    cache = KQ.KVCache(ctx=ctx)
    cv = cache.compress_vector("test", np.array(k["input"], F), 4)
    assert cv.data.tolist() == k["codes"] == [1, 7, 14, 0]
    assert beq(cv.quant_scale, F(k["scale"])) and beq(cv.quant_zero_point, F(k["zero_point"]))
    dec = cache.decompress_vector(cv)
    assert beq(dec, np.array(k["dequant"], F))
    assert np.all(np.abs(np.array(k["input"], F) - dec) < 0.1)           # the reference's own assertion
    k = KAT["D_fusion_ann_rows"]
    out = KQ.FusionANN(ctx).quantize(np.array(k["input"], F), k["bits"])
    assert len(out) == 2 and out[0].bits == 4 and out[1].bits == 8       # fusion_ann.rs:160-164
    for cv, exp in zip(out, k["rows"]):
        assert cv.data.tolist() == exp["codes"] and beq(cv.quant_scale, F(exp["scale"]))


@pytest.mark.parametrize("dim", [1, 4, 7, 8, 100, 128, 512, 2048, 4096, 8192, 20000])
def test_d_rows_bit_exact(ctx, O, dim):
    rng = np.random.default_rng(dim)
    rows = 37 if dim <= 4096 else 5
    x = rng.standard_normal((rows, dim)).astype(F)
    x[1] = 0.75                                   # constant row: 0/0 -> NaN -> code 0 (prefill_kv.rs:57-58,107)
    if dim > 4:
        x[2, 3] = np.nan
        x[3, 1] = np.inf
    for bits in ([4], [8], [1, 2, 4, 8], [2, 16]):
        c0, s0, z0 = O.quantize_d_rows(x, bits)
        c1, s1, z1 = ctx.quantize_d_rows(x, bits)
        assert beq(s0, s1) and beq(z0, z1)
        assert np.array_equal(c0, c1)
        assert beq(O.dequantize_d_rows(c0, s0, z0), ctx.dequantize_d_rows(c1, s1, z1))


@pytest.mark.parametrize("bits", [1, 2, 4, 8])
@pytest.mark.parametrize("dim", [128, 1024, 4096, 8192])
def test_d_rows_packed_device_path(ctx, O, bits, dim):
    rng = np.random.default_rng(dim + bits)
    rows = 19
    x = rng.standard_normal((rows, dim)).astype(F)
    n = rows * dim
    dx, dc, ds, dz, dout = ctx.malloc(n * 4), ctx.malloc(n), ctx.malloc(rows * 4), ctx.malloc(rows * 4), ctx.malloc(n * 4)
    ctx.h2d(dx, x)
    ctx.quantize_d_rows_dev(dx, rows, dim, bits, True, dc, ds, dz)
    ctx.dequantize_d_rows_dev(dc, rows, dim, bits, True, ds, dz, dout)
    ctx.sync()
    c0, s0, z0 = O.quantize_d_rows(x, [bits])
    assert beq(ctx.d2h(ds, (rows,), F), s0) and beq(ctx.d2h(dz, (rows,), F), z0)
    assert np.array_equal(ctx.d2h(dc, (n * bits // 8,), np.uint8), O.pack(c0, bits))
    assert beq(ctx.d2h(dout, (rows, dim), F), O.dequantize_d_rows(c0, s0, z0))
    for p in (dx, dc, ds, dz, dout):
        ctx.free(p)


# ---------------------------------------------------------------- pack / unpack
@pytest.mark.parametrize("bits", [1, 2, 4, 8])
@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 15, 16, 17, 1023, 4096, 65537])
def test_pack_unpack_bit_exact(ctx, O, bits, n):
    rng = np.random.default_rng(n * 10 + bits)
    codes = rng.integers(0, 1 << bits, n).astype(np.uint8)
    p = ctx.pack(codes, bits)
    assert np.array_equal(p, O.pack(codes, bits))
    assert np.array_equal(ctx.unpack(p, n, bits), codes)


def test_pack_rejects_other_widths(ctx):
    import dllm_b200
    for bad in (0, 3, 5, 6, 7, 16):
        with pytest.raises(dllm_b200.InvalidParams):
            ctx.pack(np.zeros(8, np.uint8), bad)


# ---------------------------------------------------------------- KV cache entry
@pytest.mark.parametrize("bits", [2, 4, 8, 3])
def test_kv_entry_tensor_scheme(ctx, O, bits):
    """QuantizedKVCacheEntry::new — one (scale, zp) per tensor, quantization.rs:140-157"""
    from dllm_b200.quantization import QuantizedKVCacheEntry
    rng = np.random.default_rng(bits)
    k = rng.standard_normal((3, 17, 64)).astype(F)
    v = (rng.standard_normal((3, 17, 64)) * 3 + 1).astype(F)
    e = QuantizedKVCacheEntry(k, v, bits, ctx)
    assert e.seq_len == 17
    for t, src in ((e.keys, k), (e.values, v)):
        c0, s0, z0 = O.quantize_tensor(src, bits)
        assert np.array_equal(t.data, c0) and beq(t.scale, s0) and beq(t.zero_point, z0)
    assert beq(e.dequantize_keys().ravel(), O.dequantize_tensor(*O.quantize_tensor(k, bits)))
    assert beq(e.dequantize_values().ravel(), O.dequantize_tensor(*O.quantize_tensor(v, bits)))
    assert e.memory_usage() == 2 * ((k.size * bits + 7) // 8)            # lib.rs:284-285
    e.close()


@pytest.mark.parametrize("scheme_bits", [(1, 4), (1, 8), (1, 2), (2, 4), (2, 8)])
def test_kv_entry_row_and_fixed_schemes(ctx, O, scheme_bits):
    from dllm_b200 import _lib as L
    from dllm_b200.quantization import QuantizedKVCacheEntry
    scheme, bits = scheme_bits
    rng = np.random.default_rng(10 * scheme + bits)
    k = rng.random((2, 9, 128)).astype(F)
    v = rng.random((2, 9, 128)).astype(F)
    e = QuantizedKVCacheEntry(k, v, bits, ctx, scheme=scheme)
    kc, vc, ks, kz, vs, vz = e._export()
    if scheme == L.KV_ROW_D:
        c0, s0, z0 = O.quantize_d_rows(k.reshape(-1, 128), [bits])
        assert np.array_equal(kc, c0.ravel()) and beq(ks, s0) and beq(kz, z0)
        assert beq(e.dequantize_keys().reshape(-1, 128), O.dequantize_d_rows(c0, s0, z0))
        c0, s0, z0 = O.quantize_d_rows(v.reshape(-1, 128), [bits])
        assert np.array_equal(vc, c0.ravel()) and beq(vs, s0) and beq(vz, z0)
    else:
        sc = O.bitquantizer_scale_c(bits)
        c0 = O.quantize_c(k, bits, sc)
        assert np.array_equal(kc, c0) and beq(ks[0], sc)
        assert beq(e.dequantize_keys().ravel(), O.dequantize_cd(c0, sc, 0.0))
    e.close()


@pytest.mark.parametrize("scheme_bits", [(1, 4), (1, 8), (1, 2), (1, 3), (2, 4), (2, 8)])
@pytest.mark.parametrize("hidden", [64, 512])
def test_kv_entry_append_only(ctx, O, scheme_bits, hidden):
    """Growing an entry token chunk by token chunk (only the new tokens are quantized) must give exactly the entry the
    reference builds by re-quantizing the whole [layers, seq, hidden] tensor (lib.rs:246-276), and the oracle's codes."""
    import dllm_b200
    from dllm_b200 import _lib as L
    from dllm_b200.quantization import QuantizedKVCacheEntry
    scheme, bits = scheme_bits
    rng = np.random.default_rng(100 * scheme + bits + hidden)
    layers, seq = 3, 37
    k = (rng.random((layers, seq, hidden)) if scheme == L.KV_FIXED_C else rng.standard_normal((layers, seq, hidden))).astype(F)
    v = (rng.random((layers, seq, hidden)) if scheme == L.KV_FIXED_C else rng.standard_normal((layers, seq, hidden))).astype(F)
    whole = QuantizedKVCacheEntry(k, v, bits, ctx, scheme=scheme)
    grown = QuantizedKVCacheEntry.with_capacity(layers, 40, hidden, bits, scheme, ctx)
    assert grown.seq_len == 0
    s = 0
    for t in (5, 1, 16, 0, 15):
        grown.append(k[:, s:s + t], v[:, s:s + t])
        s += t
        assert grown.seq_len == s and grown.shape == [layers, s, hidden]
    for a, b in zip(whole._export(), grown._export()):
        assert beq(a, b)
    assert beq(whole.dequantize_keys(), grown.dequantize_keys()) and beq(whole.dequantize_values(), grown.dequantize_values())
    assert whole.memory_usage() == grown.memory_usage()
    if scheme == L.KV_ROW_D:
        c0, s0, z0 = O.quantize_d_rows(k.reshape(-1, hidden), [bits])
        kc, _, ks, kz, _, _ = grown._export()
        assert np.array_equal(kc, c0.ravel()) and beq(ks, s0) and beq(kz, z0)
    with pytest.raises(dllm_b200.DllmError):                          # 37 + 4 > capacity 40
        grown.append(k[:, :4], v[:, :4])
    assert grown.seq_len == 37
    grown.append(k[:, :3], v[:, :3])                                    # exactly full
    assert grown.seq_len == 40
    whole.close(); grown.close()


def test_kv_entry_per_tensor_cannot_grow(ctx):
    import dllm_b200
    from dllm_b200 import _lib as L
    from dllm_b200.quantization import QuantizedKVCacheEntry
    with pytest.raises(dllm_b200.UnsupportedOperation):
        QuantizedKVCacheEntry.with_capacity(2, 8, 64, 4, L.KV_TENSOR_B, ctx)


def test_phase_aware_cache_entry(ctx, O):
    """KVCacheEntry, diffuse-llm-rs/src/lib.rs:122-313"""
    from dllm_b200.diffuse_llm import KVCacheEntry
    rng = np.random.default_rng(1)
    k = rng.standard_normal((2, 5, 32)).astype(F)
    v = rng.standard_normal((2, 5, 32)).astype(F)
    e = KVCacheEntry(k, v, 8, 4, ctx)
    assert e.get_current_quant_bits() == 8 and e.len() == 5 and not e.is_empty()
    assert beq(e.get_keys().ravel(), O.dequantize_tensor(*O.quantize_tensor(k, 8)))
    e.set_phase(False)
    assert e.get_current_quant_bits() == 4
    assert beq(e.get_values().ravel(), O.dequantize_tensor(*O.quantize_tensor(v, 4)))
    assert e.memory_usage() == 2 * (k.size * 8 // 8) + 2 * (k.size * 4 // 8)
    k2 = rng.standard_normal((2, 6, 32)).astype(F)
    e.update(k2, k2)
    assert e.len() == 6 and beq(e.get_keys().ravel(), O.dequantize_tensor(*O.quantize_tensor(k2, 4)))
    empty = KVCacheEntry(np.zeros((2, 0, 32), F), np.zeros((2, 0, 32), F), 8, 4, ctx)
    assert empty.is_empty() and empty.get_keys().shape == (2, 0, 32)


def test_adaptive_quantizer(ctx, O):
    from dllm_b200.quantization import AdaptiveQuantizer
    q = AdaptiveQuantizer(4, 4.0, ctx)
    data = (np.arange(1000) / 1000.0).astype(F)                           # quantization.rs:267-277
    q.update_stats(data)
    scale, zp = q.compute_params()
    assert scale > 0.0 and zp >= 0.0
    codes, s, z = q.quantize(data)
    assert np.array_equal(codes, O.quantize_codes_b(data, 4, s, z))
    # against the oracle's restatement of the sketch the reference feeds (CKMS(0.01), queried at q = 0 and q = 1 only,
    # quantization.rs:198-216): the same parameters bit for bit — before any data, after the reference's test data, and after
    # several chunks with outliers (the device min / max per chunk, the running extremes on the host)
    s_ref, z_ref, _ = O.adaptive_compute_params([data], 4)
    assert beq([scale, zp], [s_ref, z_ref])
    fresh = AdaptiveQuantizer(6, 4.0, ctx)
    s0, z0, _ = O.adaptive_compute_params([], 6)
    assert beq(list(fresh.compute_params()), [s0, z0])                    # unwrap_or(0.0) / unwrap_or(1.0), :208-209
    rng = np.random.default_rng(12)
    for bits in (2, 4, 8):
        chunks = [rng.standard_normal(777).astype(F) * 3 - 1, rng.standard_normal(1500).astype(F), np.array([41.5, -17.25], F),
                  rng.standard_normal(300).astype(F) * 1e-3]
        aq = AdaptiveQuantizer(bits, 4.0, ctx)
        for c in chunks:
            aq.update_stats(c)
        s_ref, z_ref, _ = O.adaptive_compute_params(chunks, bits)
        assert beq(list(aq.compute_params()), [s_ref, z_ref])
        probe = rng.standard_normal(2000).astype(F) * 10
        cq, sq, zq = aq.quantize(probe)
        assert np.array_equal(cq, O.quantize_codes_b(probe, bits, s_ref, z_ref)) and beq([sq, zq], [s_ref, z_ref])


# ---------------------------------------------------------------- full-size properties
def _torch_round_half_away(t):
    import torch
    r = torch.trunc(t)
    return r + ((t - r).abs() >= 0.5).to(t.dtype) * torch.sign(t)


def test_row_division_matches_ieee_division(ctx):
    """The per-token quantizer divides by the row's scale with a hoisted correctly-rounded reciprocal and two exact-FMA
    corrections (common.cuh: div_row).  Brute force on the device: 2^20 random divisors x (every code boundary
    1..256 x +-4 ulps, both signs, + a random numerator) = 4.8e9 quotients, all bit-identical to IEEE division."""
    import ctypes as C
    bad = C.c_uint64(123)
    ctx._ck(ctx._lib.dllm_selftest_division(ctx.h, 1 << 20, 42, C.byref(bad)))
    assert bad.value == 0


@pytest.mark.parametrize("dim", [512, 1000, 4096, 14000])
def test_d_rows_many_rows_bit_exact(ctx, O, dim):
    """Many rows (several per resident CTA, so the row loop and its register reuse are exercised): same codes, scales
    and zero-points as the oracle, with the reference's degenerate rows (constant row -> 0/0 = NaN -> code 0), NaN and
    inf entries, a cycling bit-width table (fusion_ann.rs:58) and every packed width."""
    rng = np.random.default_rng(dim)
    rows = 1000
    x = rng.standard_normal((rows, dim)).astype(F)
    x[1] = 0.75
    x[2, 3] = np.nan
    x[3, 1] = np.inf
    x[997, dim - 1] = -np.inf
    for bits in ([4], [1, 2, 4, 8], [2, 16]):
        c0, s0, z0 = O.quantize_d_rows(x, bits)
        c1, s1, z1 = ctx.quantize_d_rows(x, bits)
        assert beq(s0, s1) and beq(z0, z1)
        assert np.array_equal(c0, c1)
    n = rows * dim
    dx, dc, ds, dz = ctx.malloc(n * 4), ctx.malloc(n), ctx.malloc(rows * 4), ctx.malloc(rows * 4)
    ctx.h2d(dx, x)
    for bits in (1, 2, 4, 8):
        if (dim * bits // 8) % 4 or dim * bits % 8:
            continue                                   # packed rows must be whole 32-bit words
        ctx.quantize_d_rows_dev(dx, rows, dim, bits, True, dc, ds, dz)
        ctx.sync()
        c0, s0, z0 = O.quantize_d_rows(x, [bits])
        assert beq(ctx.d2h(ds, (rows,), F), s0) and beq(ctx.d2h(dz, (rows,), F), z0)
        assert np.array_equal(ctx.d2h(dc, (n * bits // 8,), np.uint8), O.pack(c0, bits))
    for p in (dx, dc, ds, dz):
        ctx.free(p)


@pytest.mark.parametrize("bits", [4, 8])
def test_kv_full_row_size_against_torch_restatement(ctx, bits):
    """Per-token KV quantize at the BASELINE config-4 row width (hidden 4096) on 2^16 rows
    (1 GiB of f32): bit-exact against an independent torch (IEEE f32) restatement, plus the
    round-trip bound |x - deq| < scale and code range."""
    import torch
    rows, dim = 1 << 16, 4096
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.randn(rows, dim, device="cuda", generator=g)
    codes = torch.empty(rows * dim * bits // 8, dtype=torch.uint8, device="cuda")
    scales = torch.empty(rows, device="cuda")
    zps = torch.empty(rows, device="cuda")
    out = torch.empty_like(x)
    torch.cuda.synchronize()
    ctx.quantize_d_rows_dev(x.data_ptr(), rows, dim, bits, True, codes.data_ptr(), scales.data_ptr(), zps.data_ptr())
    ctx.dequantize_d_rows_dev(codes.data_ptr(), rows, dim, bits, True, scales.data_ptr(), zps.data_ptr(), out.data_ptr())
    ctx.sync()
    mn, mx = x.min(dim=1).values, x.max(dim=1).values
    levels = float((1 << bits) - 1)
    # tensor divisors everywhere: torch turns `tensor / python_scalar` into a multiply by the reciprocal
    s_ref = (mx - mn) / torch.full_like(mx, levels)
    assert torch.equal(scales, s_ref) and torch.equal(zps, mn)
    q_ref = torch.clamp((x - mn[:, None]) / s_ref[:, None], 0.0, levels).trunc().to(torch.uint8)
    if bits == 8:
        assert torch.equal(codes.view(rows, dim), q_ref)
    else:
        lo, hi = q_ref[:, 0::2], q_ref[:, 1::2]
        assert torch.equal(codes.view(rows, dim // 2), lo | (hi << 4))
    deq_ref = q_ref.float() * s_ref[:, None] + mn[:, None]
    assert torch.equal(out, deq_ref)
    assert bool(((x - out).abs() <= s_ref[:, None] * 1.0001).all())


def test_tensor_quantize_large_against_torch_restatement(ctx):
    """Quantizer B over 2^28 elements (1 GiB): bit-exact against torch, packed 4-bit."""
    import torch
    n, bits = 1 << 28, 4
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(n, device="cuda", generator=g)
    codes = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
    params = torch.empty(4, device="cuda")
    torch.cuda.synchronize()
    ctx.quantize_tensor_dev(x.data_ptr(), n, bits, True, codes.data_ptr(), params.data_ptr())
    ctx.sync()
    mn, mx = x.min(), x.max()
    qmax = float((1 << bits) - 1)
    scale = (mx - mn) / torch.full_like(mx, qmax)      # true division (see above)
    zp = _torch_round_half_away(torch.clamp(0.0 - mn / scale, 0.0, qmax))
    assert torch.equal(params[0], scale) and torch.equal(params[1], zp)
    assert torch.equal(params[2], mn) and torch.equal(params[3], mx)
    q = torch.clamp(_torch_round_half_away(x / scale.expand_as(x) + zp), 0, qmax).to(torch.uint8)
    assert torch.equal(codes, q[0::2] | (q[1::2] << 4))
    del q
    out = torch.empty_like(x)
    ctx.dequantize_tensor_dev(codes.data_ptr(), n, bits, True, params.data_ptr(), out.data_ptr())
    ctx.sync()
    assert bool(((x - out).abs() <= scale * 0.5001).all())
