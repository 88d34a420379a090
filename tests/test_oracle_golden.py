"""Pins the CPU oracle: (1) the reference's own known-answer vectors (tests/golden/
reference_kat.json), (2) bit-for-bit agreement with the independent numpy
restatement on seeded random, ragged, empty and degenerate inputs."""
import json
import os

import numpy as np
import pytest

import np_restatement as R
from oracle import pyoracle as O

F = np.float32
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.json")))
KAT = {k["name"]: k for k in KAT}


def bits_equal(a, b):
    a, b = np.asarray(a, F), np.asarray(b, F)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---------------- reference known-answer vectors ----------------
def test_kat_d_prefill_kv():
    k = KAT["D_prefill_kv_test_quantization"]
    codes, scales, zps = O.quantize_d_rows(np.array([k["input"]], F), [k["bits"]])
    assert codes[0].tolist() == k["codes"] == [1, 7, 14, 0]
    assert bits_equal(scales[0], F(k["scale"])) and bits_equal(zps[0], F(k["zero_point"]))
    deq = O.dequantize_d_rows(codes, scales, zps)[0]
    assert bits_equal(deq, np.array(k["dequant"], F))
    # the reference's own assertion (prefill_kv.rs:156-159)
    assert np.all(np.abs(np.array(k["input"], F) - deq) < 0.1)


@pytest.mark.parametrize("name", ["B_test_quantized_tensor", "B_test_quantization"])
def test_kat_b(name):
    k = KAT[name]
    codes, s, z = O.quantize_tensor(k["input"], k["bits"])
    assert codes.tolist() == k["codes"]
    assert bits_equal(s, F(k["scale"])) and bits_equal(z, F(k["zero_point"]))
    deq = O.dequantize_tensor(codes, s, z)
    assert bits_equal(deq, np.array(k["dequant"], F))
    if name == "B_test_quantized_tensor":
        # the reference's own assertions (quantization.rs:263-264)
        assert len(deq) == 4
        r = O.compression_ratio(4, len(codes), 4)
        assert r > 4.0 and r == F(k["compression_ratio"])
    else:
        # quantization.rs:249-251 asserts <0.1, which the reference's own arithmetic violates
        # (2.0/0.26666668 = 7.4999995 -> 7).  Recorded, not "fixed".
        assert not k["ref_assert_holds"]
        assert codes.tolist() == [4, 7, 11, 15, 15]


@pytest.mark.parametrize("name", ["A_roundtrip_int8", "A_example_basic"])
def test_kat_a(name):
    k = KAT[name]
    codes = O.quantize_a(k["input"], k["qtype"], k["scale"], k["zero_point"])
    assert codes.tolist() == k["codes"]
    deq = O.dequantize_a(codes, k["scale"], k["zero_point"])
    assert bits_equal(deq, np.array(k["dequant"], F))


def test_kat_calibration():
    k = KAT["A_calibration_1_to_6"]
    s, z = O.calibrate_params(k["min"], k["max"], k["total_samples"], k["bits"], k["symmetric"])
    assert bits_equal(s, F(k["scale"])) and z == k["zero_point"] == -51
    with pytest.raises(O.OracleError):
        O.calibrate_params(0.0, 1.0, 0, 8, False)       # CalibrationRequired
    assert O.calibrate_params(2.0, 2.0, 4, 8, False) == (F(1.0), 0)  # range <= EPSILON
    s, z = O.calibrate_params(-3.0, 1.0, 4, 4, True)
    assert bits_equal(s, F(F(3.0) * F(2.0) / F(15.0))) and z == 7


def test_kat_fusion_ann_rows():
    k = KAT["D_fusion_ann_rows"]
    codes, scales, zps = O.quantize_d_rows(np.array(k["input"], F), k["bits"])
    for r, exp in enumerate(k["rows"]):
        assert codes[r].tolist() == exp["codes"]
        assert bits_equal(scales[r], F(exp["scale"])) and bits_equal(zps[r], F(exp["zero_point"]))


# ---------------- oracle == numpy restatement ----------------
def _cases(rng):
    yield rng.standard_normal(1000).astype(F)
    yield (rng.standard_normal(4097) * 0.02).astype(F)
    yield rng.random(333).astype(F)
    yield np.array([], F)
    yield np.array([3.25], F)
    yield np.full(17, -2.5, F)                       # constant: B scale->1.0, D scale 0 -> NaN
    yield np.array([0.0, -0.0, 0.0], F)
    yield np.array([1.0, np.nan, -1.0, 2.0], F)
    yield np.array([np.inf, 1.0, -1.0], F)
    yield np.array([-np.inf, 1.0, np.inf], F)
    yield np.array([1e38, -1e38, 0.5], F)            # max-min overflows to inf
    yield np.array([1e-45, 0.0, 2e-45], F)           # subnormals
    yield (rng.standard_normal(64) * 1e6).astype(F)


@pytest.mark.parametrize("bits", [1, 2, 3, 4, 5, 8])
def test_b_matches_numpy(bits):
    rng = np.random.default_rng(1234 + bits)
    for x in _cases(rng):
        c0, s0, z0 = O.quantize_tensor(x, bits)
        c1, s1, z1 = R.quantize_tensor(x, bits)
        assert bits_equal(s0, s1) and bits_equal(z0, z1), (x[:8], s0, s1, z0, z1)
        assert np.array_equal(c0, c1)
        assert bits_equal(O.dequantize_tensor(c0, s0, z0), R.dequantize_tensor(c0, s0, z0))


def test_b_bits_range_is_an_error():
    for bad in (0, 9, 255):
        with pytest.raises(O.OracleError):          # quantization.rs:39 assert!
            O.quantize_tensor([1.0, 2.0], bad)


def test_b_empty_tensor():
    c, s, z = O.quantize_tensor(np.array([], F), 4)
    assert c.size == 0 and s == F(-np.inf) and z == 0.0


@pytest.mark.parametrize("qtype", [0, 1, 2, 3])
def test_a_matches_numpy(qtype):
    rng = np.random.default_rng(99 + qtype)
    for x in _cases(rng):
        for scale, zp in ((1.0, 0), (0.05, 3), (0.5, -51)):
            assert np.array_equal(O.quantize_a(x * 10, qtype, scale, zp),
                                  R.quantize_a(x * 10, qtype, scale, zp))


@pytest.mark.parametrize("bits", [1, 2, 4, 6, 8, 16])
def test_c_matches_numpy(bits):
    rng = np.random.default_rng(7 + bits)
    sc = O.bitquantizer_scale_c(bits)
    assert bits_equal(sc, R.scale_c(bits))
    for x in _cases(rng):
        assert np.array_equal(O.quantize_c(x, bits, sc), R.quantize_c(x, bits, sc))
        c = O.quantize_c(x, bits, sc)
        assert bits_equal(O.dequantize_cd(c, sc, 0.0), R.dequantize_cd(c, sc, 0.0))


def test_c_quantize_vectors_index_rule():
    rng = np.random.default_rng(5)
    emb = rng.random((6, 3, 8)).astype(F)
    # default config [4,6,8,16]: 4 bits -> quantizers[2] = the 8-bit scale (lib.rs:133)
    codes = O.kvquant_quantize_vectors(emb, [4, 6, 8, 16], [4])
    assert np.array_equal(codes[0].ravel(), R.quantize_c(emb[0], 4, R.scale_c(8)))
    # bits cycle over vectors
    codes = O.kvquant_quantize_vectors(emb, [4, 6, 8, 16], [2, 4])
    assert np.array_equal(codes[0].ravel(), R.quantize_c(emb[0], 2, R.scale_c(6)))
    assert np.array_equal(codes[1].ravel(), R.quantize_c(emb[1], 4, R.scale_c(8)))
    # 8 bits -> index 4 of a 4-entry Vec: the reference panics
    with pytest.raises(O.OracleError) as e:
        O.kvquant_quantize_vectors(emb, [4, 6, 8, 16], [8])
    assert e.value.code == O.ERR_INDEX


@pytest.mark.parametrize("bits", [1, 2, 4, 8])
def test_d_matches_numpy(bits):
    rng = np.random.default_rng(21 + bits)
    for x in _cases(rng):
        if x.size == 0:
            continue
        codes, scales, zps = O.quantize_d_rows(x[None, :], [bits])
        c1, s1, z1 = R.quantize_d_row(x, bits)
        assert np.array_equal(codes[0], c1)
        assert bits_equal(scales[0], s1) and bits_equal(zps[0], z1)


def test_d_constant_row_is_nan_scale_code_zero():
    codes, scales, zps = O.quantize_d_rows(np.full((1, 8), 0.75, F), [4])
    assert scales[0] == 0.0 and np.all(codes == 0)   # 0/0 = NaN -> clamp NaN -> as u8 = 0


@pytest.mark.parametrize("bits", [1, 2, 4, 8])
@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 1023, 4096])
def test_pack_roundtrip_and_layout(bits, n):
    rng = np.random.default_rng(n * 10 + bits)
    codes = rng.integers(0, 1 << bits, n).astype(np.uint8)
    p = O.pack(codes, bits)
    assert p.size == (n * bits + 7) // 8 == O.packed_len(n, bits)
    assert np.array_equal(p, R.pack(codes, bits))
    assert np.array_equal(O.unpack(p, n, bits), codes)
    assert np.array_equal(R.unpack(p, n, bits), codes)


def test_pack_rejects_other_widths():
    for bad in (0, 3, 5, 6, 7, 16):
        with pytest.raises(O.OracleError):
            O.pack(np.zeros(8, np.uint8), bad)


def test_grouped_weight_quant_is_b_per_group():
    rng = np.random.default_rng(3)
    w = (rng.standard_normal((256, 24)) * 0.02).astype(F)
    codes, scales, zps = O.quantize_weight_grouped(w, 4, 128)
    for g in range(2):
        for n in (0, 5, 23):
            c, s, z = R.quantize_tensor(w[g * 128:(g + 1) * 128, n], 4)
            assert np.array_equal(codes[g * 128:(g + 1) * 128, n], c)
            assert bits_equal(scales[g, n], s) and bits_equal(zps[g, n], z)
    wd = O.dequantize_weight_grouped(codes, scales, zps, 128)
    assert np.max(np.abs(wd - w)) <= np.max(scales) * 0.5001


def test_linear_f32_vs_f64():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 300)).astype(F)
    w = (rng.standard_normal((300, 40)) * 0.02).astype(F)
    b = rng.standard_normal(40).astype(F)
    y32, y64 = O.linear_f32(x, w, b), O.linear_f64(x, w, b)
    assert np.allclose(y32, y64, rtol=1e-5, atol=1e-5)
    assert np.allclose(y64, x.astype(np.float64) @ w.astype(np.float64) + b, rtol=1e-12)
    assert np.array_equal(O.linear_f32(x, w, b, threads=3), y32)


def test_linear_i8_exact():
    rng = np.random.default_rng(1)
    qx = rng.integers(0, 256, (3, 64)).astype(np.uint8)
    qw = rng.integers(0, 16, (64, 10)).astype(np.uint8)
    acc = O.linear_i8_exact(qx, 128, qw, 7)
    ref = (qx.astype(np.int64) - 128) @ (qw.astype(np.int64) - 7)
    assert np.array_equal(acc, ref)


def test_beta_schedules_and_coeffs():
    T = 1000
    lin = O.beta_schedule(O.BETA_LINEAR, T)
    assert lin.size == 1000 and lin[0] == F(1e-4) and abs(lin[-1] - 0.02) < 1e-7
    t = np.arange(T, dtype=F)
    exp = (F(1e-4) + ((F(0.02) - F(1e-4)) * t).astype(F) / F(T - 1)).astype(F)
    assert bits_equal(lin, exp)
    quad = O.beta_schedule(O.BETA_QUADRATIC, T)
    assert np.all(np.diff(quad) >= 0) and quad[0] == F(1e-4)
    cos = O.beta_schedule(O.BETA_COSINE, T)
    assert np.all(cos <= F(0.999)) and cos[0] == 0.0
    # t=1: alpha_bar_prev = alpha_bars[0] = 1 -> c2 = 0, std = 0  (lib.rs:1162-1192)
    c1, c2, sd = O.p_sample_coeffs(lin, 1)
    assert c2 == 0.0 and sd == 0.0 and np.isfinite(c1)
    # t=0: 1 - alpha_bar_0 == 0 -> literal arithmetic is inf / NaN
    c1, c2, sd = O.p_sample_coeffs(lin, 0)
    assert np.isinf(c1) and np.isnan(c2)


def test_p_sample_guard_and_noise_rule():
    rng = np.random.default_rng(2)
    betas = O.beta_schedule(O.BETA_LINEAR, 50)
    x = rng.standard_normal((3, 16)).astype(F)
    pred = rng.standard_normal((3, 16)).astype(F)
    z = rng.standard_normal((3, 16)).astype(F)
    out = O.p_sample(x, pred, z, [7, 7, 7], betas)
    c1, c2, sd = O.p_sample_coeffs(betas, 7)
    exp = ((c1 * x).astype(F) + (c2 * pred).astype(F)).astype(F) + (sd * z).astype(F)
    assert bits_equal(out, exp.astype(F))
    # t[0]==0 -> no noise for the whole batch (lib.rs:1199-1205); guarded rows keep x_t
    out0 = O.p_sample(x, pred, z, [0, 0, 0], betas, guard_t0=True)
    assert bits_equal(out0, x)
    lit = O.p_sample(x, pred, z, [0, 0, 0], betas, guard_t0=False)
    assert np.all(np.isnan(lit))


def test_add_noise_matches_numpy_restatement():
    # lib.rs:1100-1137: noisy = x*sqrt(ab_t) + noise*sqrt(1-ab_t), t clamped to T-1 per row, f32, no FMA
    rng = np.random.default_rng(11)
    T = 50
    betas = O.beta_schedule(O.BETA_LINEAR, T)
    ab = np.ones(T, F)
    for i in range(1, T):
        ab[i] = F(ab[i - 1] * F(F(1.0) - betas[i - 1]))
    x = rng.standard_normal((5, 24)).astype(F)
    nz = rng.standard_normal((5, 24)).astype(F)
    t = [0, 1, 17, 49, 1000]
    out = O.add_noise(x, t, nz, betas)
    for b, tb in enumerate(t):
        a = ab[min(tb, T - 1)]
        sa, sd = np.sqrt(a, dtype=F), np.sqrt(F(F(1.0) - a), dtype=F)
        exp = ((x[b] * sa).astype(F) + (nz[b] * sd).astype(F)).astype(F)
        assert bits_equal(out[b], exp)
    # t = 0: alpha_bar_0 = 1 -> the input comes back unchanged (x*1 + noise*0)
    assert bits_equal(out[0], (x[0] + (nz[0] * F(0.0)).astype(F)).astype(F))
    assert O.add_noise(np.zeros((0, 8), F), [], np.zeros((0, 8), F), betas).shape == (0, 8)


KAT_NOISE_42_3 = [1067104742, 3195617084, 1019446180, 3198130443]   # f32 bit patterns of noise(seed 42, stream 3)[0:4]


def test_noise_generator_properties():
    """"dllm_noise v1": counter-based (element i of stream s depends on (seed, s, i) only), standard normal."""
    z = O.noise_normal(42, 3, 400_000)
    assert np.all(np.isfinite(z)) and abs(float(z.mean())) < 6e-3 and abs(float(z.std()) - 1) < 6e-3
    assert abs(float((z.astype(np.float64) ** 4).mean()) - 3.0) < 0.1                 # kurtosis of N(0,1)
    assert bits_equal(O.noise_normal(42, 3, 100, i0=1001), z[1001:1101])              # random access
    assert not np.array_equal(O.noise_normal(42, 4, 64), z[:64]) and not np.array_equal(O.noise_normal(43, 3, 64), z[:64])
    assert abs(float(np.corrcoef(z[0::2], z[1::2])[0, 1])) < 0.01                     # the two members of a Box-Muller pair
    # known answers (pin the constants of the specification)
    assert [int(v) for v in O.noise_normal(42, 3, 4).view(np.uint32)] == KAT_NOISE_42_3


def test_progressive_bits():
    # lib.rs:886-897 with defaults decode=4, min=2, num_steps=64
    assert O.progressive_bits(64, 63) == (int(F(4) * (F(1) - F(1 / 32)) + F(2) * F(1 / 32)), True)
    assert O.progressive_bits(64, 32) == (2, False)      # progress = 1.0
    assert O.progressive_bits(64, 0) == (0, False)       # progress = 2.0 -> 0 bits
    assert O.progressive_bits(64, 48)[1] is True


def test_linear_i8_oracle_against_plain_loops_and_the_float_oracle():
    """oracle/pyoracle.py: linear_i8 (the checker of dllm_qlinear_forward_i8): equals a plain-Python integer loop on a small
    case, and composed with the scale it is the f64 linear of the dequantized weight (quantization.rs:81-85, lib.rs:812)."""
    from oracle import pyoracle as O
    rng = np.random.default_rng(12)
    K, N, M = 24, 7, 5
    w = (rng.standard_normal((K, N)) * 0.02).astype(np.float32)
    codes, scale, zp = O.quantize_tensor(w, 4)
    codes = codes.reshape(K, N)
    xq = rng.integers(-128, 128, (M, K)).astype(np.int8)
    y = O.linear_i8(xq, codes, zp)
    exp = [[sum(int(xq[m, k]) * (int(codes[k, n]) - int(zp)) for k in range(K)) for n in range(N)] for m in range(M)]
    assert y.dtype == np.int64 and y.tolist() == exp
    wd = O.dequantize_tensor(codes.reshape(-1), scale, zp).reshape(K, N)
    yf = O.linear_f64(xq.astype(np.float32), wd, None)
    bound = np.abs(xq).astype(np.float64) @ np.abs(wd).astype(np.float64)      # wd carries one f32 rounding per weight
    assert np.all(np.abs(yf - np.float64(scale) * y) <= 1e-6 * bound + 1e-12)
    with pytest.raises(AssertionError):
        O.linear_i8(xq, codes, 2.5)                       # quantize_tensor's zero-points are integers


def test_int8_mode_restatement_stays_within_the_stated_bound_of_the_f64_stack():
    """The oracle's restatement of the int8 denoise mode (DLLM_PATH_I8; not a reference mode): bf16 rounding equals torch's,
    the activation quantizer's codes / sums / steps are self-consistent, and the stack stays within 1e-2 * sqrt(n_linears)
    relative Frobenius error of the reference's arithmetic in f64 (lib.rs:812 on dequantize_tensor's output)."""
    import torch
    from oracle import pyoracle as O
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((64, 1000)) * 3).astype(np.float32)
    x[0, :3] = [0.0, -0.0, 1e-40]
    assert np.array_equal(O.bf16_round(x).view(np.uint32), torch.from_numpy(x).bfloat16().float().numpy().view(np.uint32))
    q, rs, sm = O.rowquant_i8(O.bf16_round(x), 0.5)
    assert q.dtype == np.int8 and np.abs(q).max() == 127 and np.array_equal(sm, q.astype(np.int64).sum(1))
    assert np.all(np.abs(q.astype(np.float64) * (rs / 0.5)[:, None] - O.bf16_round(x)) <= 0.5001 * (rs / 0.5)[:, None])
    q0, rs0, sm0 = O.rowquant_i8(np.zeros((2, 64), np.float32), 0.25)
    assert not q0.any() and np.all(rs0 == 0.25) and not sm0.any()
    for bits in (4, 8):
        dims = [256, 256, 512, 256, 256]
        layers = []
        for K, N in zip(dims[:-1], dims[1:]):
            w = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
            b = (rng.standard_normal(N) * 0.1).astype(np.float32)
            c, s, z = O.quantize_tensor(w, bits)
            layers.append((c.reshape(K, N), s, z, b))
        xt = rng.standard_normal((256, 256)).astype(np.float32)
        h = xt.astype(np.float64)
        for c, s, z, b in layers:
            h = h @ ((c.astype(np.float64) - z) * s) + b
        y = O.model_forward_i8(xt, layers)
        assert np.linalg.norm(y - h) <= 1e-2 * np.sqrt(len(layers)) * np.linalg.norm(h)
        c, s, z, b = layers[0]
        y1 = O.linear_i8_deq(xt, c, s, z, b)
        h1 = xt.astype(np.float64) @ ((c.astype(np.float64) - z) * s) + b
        assert np.linalg.norm(y1 - h1) <= 1e-2 * np.linalg.norm(h1)


def test_ckms_sketch_answers_q0_and_q1_with_the_exact_extremes():
    """AdaptiveQuantizer (quantization.rs:179-216) asks its CKMS(0.01) sketch for q = 0.0 and q = 1.0 only.  The oracle's
    restatement of the published algorithm returns the exact minimum / maximum of everything inserted so far — for random,
    sorted, reverse-sorted, constant, duplicate-heavy and outlier streams, at every prefix checked — while it does compress
    (far fewer samples than inserts) and stays within its rank error in between."""
    from oracle import pyoracle as O
    rng = np.random.default_rng(7)
    streams = {
        "normal": rng.standard_normal(6000),
        "sorted": np.sort(rng.standard_normal(4000)),
        "reverse": np.sort(rng.standard_normal(4000))[::-1],
        "constant": np.full(1500, 0.25),
        "duplicates": rng.integers(-3, 4, 5000).astype(np.float64),
        "outliers": np.concatenate([rng.standard_normal(3000), [1e30, -1e30], rng.standard_normal(1000) * 1e-20]),
        "ramp": np.arange(1000) / 1000.0,                       # the reference's own test data (quantization.rs:267-277)
    }
    for name, data in streams.items():
        data = data.astype(np.float32)
        sk = O.CKMS(0.01)
        for i, v in enumerate(data):
            sk.insert(v)
            if i % 397 == 0 or i + 1 == len(data):
                assert sk.query(0.0)[1] == data[: i + 1].min(), name
                assert sk.query(1.0)[1] == data[: i + 1].max(), name
        assert sk.n == len(data) and sum(e[1] for e in sk.samples) == len(data)
        if name in ("normal", "sorted", "reverse"):
            assert len(sk.samples) < len(data) // 4, (name, len(sk.samples))          # it does compress
            srt = np.sort(data)
            for q in (0.1, 0.5, 0.9):                                                    # and stays a quantile sketch
                _, v = sk.query(q)
                rank = np.searchsorted(srt, v, side="left")
                assert abs(rank - q * len(data)) <= 0.01 * 2 * q * len(data) + 2, (name, q, rank)
    assert O.CKMS(0.01).query(0.0) is None                                                # -> unwrap_or(0.0) / unwrap_or(1.0)
    # compute_params on the reference's test data: scale = (0.999 - 0) / 15, zero-point 0
    scale, zp, _ = O.adaptive_compute_params([streams["ramp"]], 4)
    assert scale == np.float32(np.float32(0.999) / np.float32(15)) and zp == 0.0
    s0, z0, _ = O.adaptive_compute_params([], 4)
    assert s0 == np.float32(np.float32(1.0) / np.float32(15)) and z0 == 0.0
