"""GPU parity of the quantized weight object and the dequant-matmul paths against the oracle.

Tolerances (BASELINE.json north_star):
  * weight quantization / repack / export: bit-exact;
  * SIMT path (f32 dequant exactly as the reference, f32 accumulate): max|y - y64| <= 2e-5 * (|x|·|W|)
    — f32 summation-order noise only;
  * tcgen05 path (bf16 operands, f32 accumulate): max|y - y64| <= 1e-2 * max|y64|  and
    relative Frobenius error <= 1e-2.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F = np.float32


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


def make_w(rng, K, N, std=0.02):
    return (rng.standard_normal((K, N)) * std).astype(F)     # lib.rs:792-796 init


# ---------------------------------------------------------------- weight object
@pytest.mark.parametrize("bits", [2, 3, 4, 5, 8])
@pytest.mark.parametrize("shape", [(128, 128), (256, 200), (512, 384), (1024, 130)])
def test_grouped_weight_quantize_bit_exact(ctx, O, bits, shape):
    from dllm_b200 import QWeight
    K, N = shape
    rng = np.random.default_rng(K + N + bits)
    w = make_w(rng, K, N)
    qw = QWeight.quantize(ctx, w, bits, 128)
    codes, scales, zps = qw.export()
    c0, s0, z0 = O.quantize_weight_grouped(w, bits, 128)
    assert beq(scales, s0) and beq(zps, z0)
    assert np.array_equal(codes, c0)
    qw.close()


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("shape", [(64, 128), (100, 50), (333, 257)])
def test_per_tensor_weight_quantize_bit_exact(ctx, O, bits, shape):
    """group == 0: the reference's per-tensor quantize_tensor over the whole [K,N] array."""
    from dllm_b200 import QWeight
    K, N = shape
    rng = np.random.default_rng(K * N + bits)
    w = make_w(rng, K, N)
    qw = QWeight.quantize(ctx, w, bits, 0)
    codes, scales, zps = qw.export()
    c0, s0, z0 = O.quantize_tensor(w, bits)
    assert beq(scales[0, 0], s0) and beq(zps[0, 0], z0)
    assert np.array_equal(codes.ravel(), c0)
    qw.close()


@pytest.mark.parametrize("bits", [2, 4, 8])
def test_from_codes_roundtrip(ctx, O, bits):
    from dllm_b200 import QWeight
    rng = np.random.default_rng(bits)
    K, N = 384, 200
    codes = rng.integers(0, 1 << bits, (K, N)).astype(np.uint8)
    scales = (rng.random((K // 128, N)) * 0.01 + 0.001).astype(F)
    zps = rng.integers(0, 1 << bits, (K // 128, N)).astype(F)
    qw = QWeight.from_codes(ctx, codes, scales, zps, bits, 128)
    c1, s1, z1 = qw.export()
    assert np.array_equal(c1, codes) and beq(s1, scales) and beq(z1, zps)
    assert qw.packed_bytes == ((N + 127) // 128) * 128 * K * (2 if bits <= 2 else 4 if bits <= 4 else 8) // 8
    qw.close()


def test_weight_argument_errors(ctx):
    import dllm_b200
    from dllm_b200 import QWeight
    w = np.zeros((128, 16), F)
    with pytest.raises(dllm_b200.InvalidParams):
        QWeight.quantize(ctx, w, 0, 128)
    with pytest.raises(dllm_b200.InvalidParams):
        QWeight.quantize(ctx, w, 9, 128)
    with pytest.raises(dllm_b200.ShapeMismatch):
        QWeight.quantize(ctx, w, 4, 96)            # group must be a multiple of 64
    with pytest.raises(dllm_b200.ShapeMismatch):
        QWeight.quantize(ctx, np.zeros((192, 16), F), 4, 128)   # K % group != 0


# ---------------------------------------------------------------- linear, SIMT path
def ref_linear(O, x, w, bits, group, bias):
    if group:
        c, s, z = O.quantize_weight_grouped(w, bits, group)
        wd = O.dequantize_weight_grouped(c, s, z, group)
    else:
        c, s, z = O.quantize_tensor(w, bits)
        wd = O.dequantize_tensor(c, s, z).reshape(w.shape)
    y64 = O.linear_f64(x, wd, bias)
    bound = np.abs(x).astype(np.float64) @ np.abs(wd).astype(np.float64)
    return wd, y64, bound


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("M", [1, 2, 3, 4, 7, 8, 16, 33])
def test_qlinear_simt_matches_oracle(ctx, O, bits, M):
    from dllm_b200 import QWeight, PATH_SIMT
    rng = np.random.default_rng(M * 10 + bits)
    K, N = 1024, 384
    w = make_w(rng, K, N)
    x = rng.standard_normal((M, K)).astype(F)
    bias = rng.standard_normal(N).astype(F)
    qw = QWeight.quantize(ctx, w, bits, 128, bias)
    y = qw.forward(x, PATH_SIMT)
    wd, y64, bound = ref_linear(O, x, w, bits, 128, bias)
    assert np.all(np.abs(y - y64) <= 2e-5 * bound + 1e-6)
    # and it agrees with the oracle's own f32 sequential linear to f32 noise
    y32 = O.linear_f32(x, wd, bias)
    assert np.allclose(y, y32, rtol=0, atol=2e-5 * bound.max())
    qw.close()


@pytest.mark.parametrize("shape", [(64, 10), (200, 130), (4096, 256)])
def test_qlinear_simt_ragged_shapes_and_per_tensor(ctx, O, shape):
    from dllm_b200 import QWeight, PATH_SIMT
    K, N = shape
    rng = np.random.default_rng(K + N)
    w = make_w(rng, K, N)
    x = rng.standard_normal((5, K)).astype(F)
    qw = QWeight.quantize(ctx, w, 4, 0)            # per-tensor (reference behaviour), K, N ragged
    y = qw.forward(x, PATH_SIMT)
    _, y64, bound = ref_linear(O, x, w, 4, 0, None)
    assert np.all(np.abs(y - y64) <= 2e-5 * bound + 1e-6)
    qw.close()


def test_dequant_matmul_example_config0(ctx, O):
    """BASELINE.json configs[0]: 4-bit group-quantize + dequant-matmul, 4096x4096 weight, 8 tokens."""
    import dllm_b200
    rng = np.random.default_rng(42)
    K = N = 4096
    w = make_w(rng, K, N)
    x = rng.standard_normal((8, K)).astype(F)
    c, s, z = O.quantize_weight_grouped(w, 4, 128)
    y = dllm_b200.dequant_matmul(ctx, c, s, z, x, 4, 128, None, dllm_b200.PATH_SIMT)
    wd = O.dequantize_weight_grouped(c, s, z, 128)
    y32 = O.linear_f32(x, wd, None, threads=8)
    bound = float((np.abs(x).astype(np.float64) @ np.abs(wd).astype(np.float64)).max())
    assert np.max(np.abs(y - y32)) <= 4e-5 * bound


# ---------------------------------------------------------------- linear, HBM-bound GEMV path (1..16 tokens)
def gemv_check(y, y64, bound, x=None, wd=None):
    """int8 tensor path: codes, zero-points and the integer sums are exact; x is rounded to block fixed point per aligned
    block of 128 activations (two 64-k blocks share their power-of-two step, so that their integer sums can be added before
    the float work; error <= 2^-22 of the block's max |x| for 1-2 tokens, 2^-14 for 3-16), scale and accumulation across
    blocks are f32.  With x and the dequantized weight given, every output is checked against
    B2 = sum over blocks of max|x_block| * sum|w_block|: 2^-14 B2 for 3-16 tokens; for 1-2 tokens the f32 rounding of the
    per-block terms (each up to max|x_block| * sum|w_block|) is of the same order as the 2^-22 digit error: 2^-20 B2.
    Always also against 2^-12 * sum|x||w| (ample for Gaussian x)."""
    err = np.abs(y - y64)
    if x is not None:
        M, K = x.shape
        kb = (K + 127) // 128
        xp = np.zeros((M, kb * 128)); xp[:, :K] = np.abs(x)
        wp = np.zeros((kb * 128, wd.shape[1])); wp[:K] = np.abs(wd)
        bmax = xp.reshape(M, kb, 128).max(axis=2)                        # [M, kb]
        wsum = wp.reshape(kb, 128, -1).sum(axis=1)                        # [kb, N]
        tight = (2.0 ** -20 if M <= 2 else 2.0 ** -14 + 2.0 ** -20) * (bmax @ wsum)
        assert np.all(err <= tight + 1e-6), float((err / (tight + 1e-30)).max())
    else:
        assert np.all(err <= (2.0 ** -12 + 5e-6) * bound + 1e-6), float((err / (bound + 1e-30)).max())
    assert np.linalg.norm(y - y64) <= 1e-3 * np.linalg.norm(y64)


@pytest.mark.parametrize("bits", [2, 3, 4, 5, 8])
@pytest.mark.parametrize("M", [1, 2, 3, 4, 7, 8, 9, 16])
def test_qlinear_gemv_matches_oracle(ctx, O, bits, M):
    from dllm_b200 import QWeight, PATH_GEMV
    rng = np.random.default_rng(M * 100 + bits)
    K, N = 1024, 384
    w = make_w(rng, K, N)
    x = rng.standard_normal((M, K)).astype(F)
    bias = rng.standard_normal(N).astype(F)
    qw = QWeight.quantize(ctx, w, bits, 128, bias)
    y = qw.forward(x, PATH_GEMV)
    wd, y64, bound = ref_linear(O, x, w, bits, 128, bias)
    gemv_check(y, y64, bound, x, wd)
    qw.close()


@pytest.mark.parametrize("M", [1, 2, 4, 16])
def test_qlinear_gemv_activation_outliers(ctx, O, M):
    """Activations with a few huge entries, exact zeros, tiny values and whole zero blocks: the block fixed point must
    stay within its stated bound (the outlier's block loses the small values' low bits, nothing else does)."""
    from dllm_b200 import QWeight, PATH_GEMV
    rng = np.random.default_rng(900 + M)
    K, N = 1024, 256
    w = make_w(rng, K, N)
    x = rng.standard_normal((M, K)).astype(F)
    x[:, 5] *= 1e4
    x[:, 700] = -3e6
    x[:, 128:192] = 0.0
    x[:, 300:364] *= 1e-30
    x[0, 64:128] = 2.0 ** 10                                            # a block of equal powers of two
    qw = QWeight.quantize(ctx, w, 4, 128)
    y = qw.forward(x, PATH_GEMV)
    wd, y64, bound = ref_linear(O, x, w, 4, 128, None)
    gemv_check(y, y64, bound, x, wd)
    qw.close()


@pytest.mark.parametrize("shape", [(64, 10, 5), (200, 130, 5), (4096, 256, 3), (8192, 512, 16), (14336, 256, 16),
                                   (14336, 128, 8), (2048, 4096, 1)])
@pytest.mark.parametrize("group", [0, 128])
def test_qlinear_gemv_shapes(ctx, O, shape, group):
    """ragged K / N (padding), per-tensor parameters (group 0, the reference's own scheme), and K large
    enough that the activations are split into k segments (several CTAs reduce into one tile)."""
    from dllm_b200 import QWeight, PATH_GEMV
    K, N, M = shape
    if group and K % group:
        pytest.skip("grouped weights need K % group == 0")
    rng = np.random.default_rng(K + N + M)
    w = make_w(rng, K, N)
    x = rng.standard_normal((M, K)).astype(F)
    qw = QWeight.quantize(ctx, w, 4, group)
    y = qw.forward(x, PATH_GEMV)
    _, y64, bound = ref_linear(O, x, w, 4, group, None)
    gemv_check(y, y64, bound)
    y2 = qw.forward(x, PATH_GEMV)              # tickets re-armed; fixed reduction order
    assert beq(y, y2)
    qw.close()


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("group,K", [(64, 448), (192, 1152), (256, 2048), (128, 1408), (0, 1472), (128, 14336)])
@pytest.mark.parametrize("M", [1, 5, 16])
def test_qlinear_gemv_stage_shapes(ctx, O, bits, group, K, M):
    """The ring stages hold aligned groups of 2 or 4 k-blocks and the two k-blocks of an aligned pair are summed in int32
    when they share their parameters: groups of 64 (never shared), 192 (three k-blocks: pairs straddle groups), 256 and
    per-tensor parameters (whole stages shared), odd k-block counts, and stream-K ranges that start in the middle of a pair."""
    from dllm_b200 import QWeight, PATH_GEMV
    rng = np.random.default_rng(K + M + bits + group)
    N = 384 if K < 8192 else 256
    w = make_w(rng, K, N)
    x = rng.standard_normal((M, K)).astype(F)
    x[:, 64:128] *= 50.0                         # the two halves of an aligned pair differ in magnitude
    bias = rng.standard_normal(N).astype(F)
    qw = QWeight.quantize(ctx, w, bits, group, bias)
    y = qw.forward(x, PATH_GEMV)
    wd, y64, bound = ref_linear(O, x, w, bits, group, bias)
    gemv_check(y, y64, bound, x, wd)
    assert beq(y, qw.forward(x, PATH_GEMV))
    qw.close()


def test_qlinear_auto_path_small_m_is_gemv(ctx, O):
    from dllm_b200 import QWeight, PATH_AUTO, PATH_GEMV
    rng = np.random.default_rng(5)
    w = make_w(rng, 512, 256)
    x = rng.standard_normal((4, 512)).astype(F)
    qw = QWeight.quantize(ctx, w, 4, 128)
    assert beq(qw.forward(x, PATH_AUTO), qw.forward(x, PATH_GEMV))
    import dllm_b200
    with pytest.raises(dllm_b200.QuantizationError):
        qw.forward(rng.standard_normal((17, 512)).astype(F), PATH_GEMV)
    qw.close()


def test_fractional_zero_points_take_the_f32_path(ctx, O):
    """Hand-made parameters with non-integer zero-points (quantizer B never produces them): the 16-bit-operand
    kernels subtract the zero-point exactly only when it is an integer, so AUTO must fall back to the f32-faithful
    SIMT kernel and an explicit request for the other paths must fail loudly."""
    import dllm_b200
    from dllm_b200 import QWeight, PATH_AUTO, PATH_SIMT, PATH_GEMV, PATH_UMMA
    rng = np.random.default_rng(11)
    K, N = 256, 128
    codes = rng.integers(0, 16, (K, N)).astype(np.uint8)
    scales = (rng.random((K // 128, N)) * 0.01 + 0.001).astype(F)
    zps = (rng.integers(0, 15, (K // 128, N)) + 0.5).astype(F)
    qw = QWeight.from_codes(ctx, codes, scales, zps, 4, 128)
    x = rng.standard_normal((4, K)).astype(F)
    y = qw.forward(x, PATH_AUTO)
    assert beq(y, qw.forward(x, PATH_SIMT))
    wd = O.dequantize_weight_grouped(codes, scales, zps, 128)
    bound = np.abs(x).astype(np.float64) @ np.abs(wd).astype(np.float64)
    assert np.all(np.abs(y - O.linear_f64(x, wd, None)) <= 2e-5 * bound + 1e-6)
    for path in (PATH_GEMV, PATH_UMMA):
        with pytest.raises(dllm_b200.UnsupportedOperation):
            qw.forward(x, path)
    qw.close()


@pytest.mark.parametrize("bits,M", [(4, 1), (4, 16), (2, 4), (8, 8)])
def test_qlinear_gemv_full_width(ctx, bits, M):
    """Full BASELINE width (K=N=14336): the GEMV path against the f32-faithful SIMT path, run-to-run
    determinism, and exact linearity under power-of-two scaling."""
    import torch
    from dllm_b200 import QWeight, PATH_SIMT, PATH_GEMV
    K = N = 14336
    g = torch.Generator(device="cuda").manual_seed(bits * 100 + M)
    w = torch.randn(K, N, device="cuda", generator=g) * 0.02
    x = torch.randn(M, K, device="cuda", generator=g)
    y1, y2, y3, y4 = (torch.empty(M, N, device="cuda") for _ in range(4))
    torch.cuda.synchronize()
    qw = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128)
    qw.forward_dev(x.data_ptr(), M, y1.data_ptr(), PATH_SIMT)
    qw.forward_dev(x.data_ptr(), M, y2.data_ptr(), PATH_GEMV)
    qw.forward_dev(x.data_ptr(), M, y3.data_ptr(), PATH_GEMV)
    x2 = x * 4
    torch.cuda.synchronize()
    qw.forward_dev(x2.data_ptr(), M, y4.data_ptr(), PATH_GEMV)
    ctx.sync()
    assert float((y1 - y2).abs().max()) <= 4e-3 * float(y1.abs().max())
    assert float(torch.linalg.norm(y1 - y2)) <= 3e-3 * float(torch.linalg.norm(y1))
    assert torch.equal(y2, y3)
    # power-of-two scaling commutes with every rounding of the path except for the handful of activations that are
    # fp16-subnormal (|x| < 6.1e-5) before the scaling
    assert float((y4 - y2 * 4).abs().max()) <= 1e-5 * float(y2.abs().max())
    qw.close()


# ---------------------------------------------------------------- exact int8 linear (tcgen05 kind::i8)
@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("shape", [(64, 128, 1), (256, 130, 16), (1024, 384, 100), (4160, 256, 300), (2048, 1000, 33)])
def test_qlinear_i8_exact(ctx, O, bits, shape):
    """int8 activations x per-tensor quantized codes -> int32, tolerance 0 (BASELINE.json north_star): every output equals
    the int64 oracle, for ragged N / token counts, K with an odd number of k-blocks, and activations that include -128."""
    from dllm_b200 import QWeight
    K, N, M = shape
    rng = np.random.default_rng(K + N + M + bits)
    w = make_w(rng, K, N)
    qw = QWeight.quantize(ctx, w, bits, 0)
    codes, scales, zps = qw.export()
    c0, s0, z0 = O.quantize_tensor(w, bits)
    assert np.array_equal(codes.reshape(-1), c0) and float(np.ravel(scales)[0]) == s0 and float(np.ravel(zps)[0]) == z0
    xq = rng.integers(-128, 128, (M, K)).astype(np.int8)
    xq[0, :7] = -128
    xq[-1, -5:] = 127
    y = qw.forward_i8(xq)
    exp = O.linear_i8(xq, codes.reshape(K, N), z0)
    assert y.dtype == np.int32 and np.array_equal(y.astype(np.int64), exp)
    # composed with the scales it is the reference's float linear (f32-faithful path on the same integers)
    from dllm_b200 import PATH_SIMT
    yf = qw.forward(xq.astype(F), PATH_SIMT)
    assert np.allclose(yf, s0 * exp, rtol=2e-5, atol=2e-5 * float(np.abs(s0 * exp).max()))
    qw.close()


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("shape", [(64, 128, 1), (256, 130, 16), (1024, 384, 100), (2048, 1000, 333), (4160, 256, 300)])
def test_qlinear_i8_mode_matches_oracle(ctx, O, bits, shape):
    """int8 denoise mode (DLLM_PATH_I8) on f32 activations: per-token int8 activation quantizer + exact kind::i8 contraction
    + fused `(q - zp) * scale` (quantization.rs:83) epilogue.  Against the oracle's restatement of the same arithmetic: 1 ulp
    of f32 (integer part exact, one fused multiply-add); against the reference's f64 linear: <= 1e-2 relative Frobenius error
    (the activation step is max|x_row| / 127; measured 0.7e-2 on N(0,1) rows)."""
    from dllm_b200 import QWeight, PATH_I8
    K, N, M = shape
    rng = np.random.default_rng(K + N + M + bits)
    w = make_w(rng, K, N)
    b = (rng.standard_normal(N) * 0.1).astype(F)
    qw = QWeight.quantize(ctx, w, bits, 0, b)
    codes, scales, zps = qw.export()
    s0, z0 = float(np.ravel(scales)[0]), float(np.ravel(zps)[0])
    x = rng.standard_normal((M, K)).astype(F)
    x[0] *= 1e-3                       # a quiet token keeps its own step
    if M > 2:
        x[1] = 0.0                     # an all-zero token: codes 0, output = bias
    y = qw.forward(x, PATH_I8)
    exp = O.linear_i8_deq(x, codes.reshape(K, N), s0, z0, b)
    assert np.all(np.abs(y.astype(np.float64) - exp) <= 1.2e-7 * np.abs(exp) + 1e-30)
    if M > 2:
        assert beq(y[1], b)
    y64 = x.astype(np.float64) @ ((codes.reshape(K, N).astype(np.float64) - z0) * s0) + b
    assert np.linalg.norm(y - y64) <= 1e-2 * np.linalg.norm(y64)
    qw.close()


@pytest.mark.parametrize("bits", [2, 4])
@pytest.mark.parametrize("shape", [(512, 1024, 2560), (576, 1000, 2600), (2048, 2048, 4096), (192, 1536, 1111)])
def test_qlinear_i8_mode_dense_shapes(ctx, O, bits, shape):
    """The same on dense shapes (>= 1024 tokens): the CTA-pair kind::i8 kernel (256-token tiles), f32 output.  Ragged token
    counts and columns, an odd number of k-blocks; same two bounds as above."""
    from dllm_b200 import QWeight, PATH_I8
    K, N, M = shape
    rng = np.random.default_rng(K + N + M + bits)
    w = make_w(rng, K, N)
    b = (rng.standard_normal(N) * 0.1).astype(F)
    qw = QWeight.quantize(ctx, w, bits, 0, b)
    codes, scales, zps = qw.export()
    s0, z0 = float(np.ravel(scales)[0]), float(np.ravel(zps)[0])
    x = (rng.standard_normal((M, K)) * rng.uniform(0.1, 10.0, (M, 1))).astype(F)     # every token its own range
    x[5] = 0.0
    y = qw.forward(x, PATH_I8)
    exp = O.linear_i8_deq(x, codes.reshape(K, N), s0, z0, b)
    assert np.all(np.abs(y.astype(np.float64) - exp) <= 1.2e-7 * np.abs(exp) + 1e-30)
    assert beq(y[5], b)
    y64 = x.astype(np.float64) @ ((codes.reshape(K, N).astype(np.float64) - z0) * s0) + b
    assert np.linalg.norm(y - y64) <= 1e-2 * np.linalg.norm(y64)
    qw.close()


def test_qlinear_i8_mode_refuses_grouped_weights(ctx):
    import dllm_b200
    from dllm_b200 import QWeight, PATH_I8
    rng = np.random.default_rng(3)
    qw = QWeight.quantize(ctx, make_w(rng, 256, 128), 4, 128)
    with pytest.raises(dllm_b200.UnsupportedOperation):
        qw.forward(rng.standard_normal((4, 256)).astype(F), PATH_I8)
    qw.close()


def test_qlinear_i8_extremes_and_errors(ctx, O):
    """All-maximum operands (the int32 accumulator's worst case for K = 4096), and the shapes the int8 path refuses."""
    import dllm_b200
    from dllm_b200 import QWeight
    K, N, M = 4096, 128, 16
    codes = np.full((K, N), 255, np.uint8)
    codes[:, 1] = 0
    qw = QWeight.from_codes(ctx, codes, np.array([0.01], F), np.array([3.0], F), 8, 0)
    xq = np.full((M, K), -128, np.int8)
    xq[1] = 127
    assert np.array_equal(qw.forward_i8(xq).astype(np.int64), O.linear_i8(xq, codes, 3.0))
    qw.close()
    rng = np.random.default_rng(3)
    grouped = QWeight.quantize(ctx, make_w(rng, 256, 128), 4, 128)
    with pytest.raises(dllm_b200.UnsupportedOperation):
        grouped.forward_i8(np.zeros((4, 256), np.int8))
    grouped.close()
    ragged = QWeight.quantize(ctx, make_w(rng, 200, 128), 4, 0)
    with pytest.raises(dllm_b200.UnsupportedOperation):
        ragged.forward_i8(np.zeros((4, 200), np.int8))
    ragged.close()


# ---------------------------------------------------------------- linear, tcgen05 path
def umma_check(y, y64):
    err = np.abs(y - y64)
    assert err.max() <= 1e-2 * np.abs(y64).max(), (err.max(), np.abs(y64).max())
    assert np.linalg.norm(y - y64) <= 1e-2 * np.linalg.norm(y64)


@pytest.mark.parametrize("bits", [4, 2, 8])
@pytest.mark.parametrize("M", [1, 8, 16, 128, 300])
def test_qlinear_umma_matches_oracle(ctx, O, bits, M):
    from dllm_b200 import QWeight, PATH_UMMA
    rng = np.random.default_rng(M + bits)
    K, N = 1024, 512
    w = make_w(rng, K, N)
    x = rng.standard_normal((M, K)).astype(F)
    bias = rng.standard_normal(N).astype(F)
    qw = QWeight.quantize(ctx, w, bits, 128, bias)
    y = qw.forward(x, PATH_UMMA)
    _, y64, _ = ref_linear(O, x, w, bits, 128, bias)
    umma_check(y, y64)
    qw.close()


@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 130, 512), (2048, 2048, 256), (4096, 1024, 8192 // 8)])
def test_qlinear_umma_shapes(ctx, O, shape):
    from dllm_b200 import QWeight, PATH_UMMA
    K, N, M = shape
    rng = np.random.default_rng(K + N + M)
    w = make_w(rng, K, N)
    x = rng.standard_normal((M, K)).astype(F)
    qw = QWeight.quantize(ctx, w, 4, 128)
    y = qw.forward(x, PATH_UMMA)
    c, s, z = O.quantize_weight_grouped(w, 4, 128)
    wd = O.dequantize_weight_grouped(c, s, z, 128)
    y64 = x.astype(np.float64) @ wd.astype(np.float64)
    umma_check(y, y64)
    qw.close()


def test_qlinear_umma_vs_simt_full_width(ctx):
    """Full BASELINE width (K=N=14336, 4-bit, M=16): both device paths agree to bf16 tolerance.
    (The CPU oracle needs minutes at this size; the two GPU paths share only the packed weights.)"""
    import torch
    from dllm_b200 import QWeight, PATH_SIMT, PATH_UMMA
    K = N = 14336
    M = 16
    g = torch.Generator(device="cuda").manual_seed(1)
    w = torch.randn(K, N, device="cuda", generator=g) * 0.02
    x = torch.randn(M, K, device="cuda", generator=g)
    y1 = torch.empty(M, N, device="cuda")
    y2 = torch.empty(M, N, device="cuda")
    torch.cuda.synchronize()
    qw = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, 4, 128)
    qw.forward_dev(x.data_ptr(), M, y1.data_ptr(), PATH_SIMT)
    qw.forward_dev(x.data_ptr(), M, y2.data_ptr(), PATH_UMMA)
    ctx.sync()
    assert float((y1 - y2).abs().max()) <= 1e-2 * float(y1.abs().max())
    # linearity property of the SIMT path: f(2x) == 2 f(x) exactly in f32 (power-of-two scaling)
    x2 = x * 2
    y3 = torch.empty_like(y1)
    torch.cuda.synchronize()
    qw.forward_dev(x2.data_ptr(), M, y3.data_ptr(), PATH_SIMT)
    ctx.sync()
    assert torch.equal(y3, y1 * 2)
    qw.close()


def _pair_check(mode, shape=None, ntok2=None):
    import os
    import subprocess
    import sys
    import tempfile
    script = os.path.join(os.path.dirname(__file__), "umma_pair_check.py")
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "y.npy")
        env = dict(os.environ, DLLM_UMMA_PAIR=str(mode))
        env.pop("DLLM_UMMA_NTOK2", None)
        if ntok2:
            env["DLLM_UMMA_NTOK2"] = str(ntok2)
        cmd = [sys.executable, script, out] + ([str(v) for v in shape] if shape else [])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        assert "PAIR_CHECK_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
        return np.load(out)


def test_umma_cta_pair_variant_is_bit_identical():
    """The CTA-pair (tcgen05 cta_group::2) kernels — the 256-token pair kernel the denoise step runs (DLLM_UMMA_PAIR=2, the
    default for dense problems) and round 1's 128-token variant (=1) — must produce the 1-CTA kernel's output (=0) bit for
    bit on a dense, ragged shape with an odd number of column tiles."""
    base = _pair_check(0)
    for mode in (1, 2):
        assert np.array_equal(base.view(np.uint32), _pair_check(mode).view(np.uint32)), f"pair mode {mode}"


@pytest.mark.parametrize("ntok2", [128, 160, 192, 224, 256])
def test_umma_pair2_tile_widths_bit_identical(ntok2):
    """Every tile width the 256-token pair kernel may pick (MMA N = ntok/2 = 64..128, token halves of 32..64 rows per CTA),
    on a shape with an odd k-block count (K = 448 with groups of 64: seven k-blocks, the last stage holds one), ragged tokens and ragged columns."""
    shape = (448, 7000, 2100, 64)   # >= 4 tiles per SM: the 1-CTA kernel runs whole tiles too (same summation order)
    base = _pair_check(0, shape)
    assert np.array_equal(base.view(np.uint32), _pair_check(2, shape, ntok2).view(np.uint32))


@pytest.mark.parametrize("KN", [(2048, 2048), (2048, 8192), (8192, 2048)])
def test_umma_benchmark_shapes_match_oracle(ctx, O, KN):
    """The launches the headline number is made of: 8192 tokens x the three layer shapes of the 1B-class stack, full size,
    through the default dense path.  Weight codes / scales / zero-points must equal the oracle's bit for bit; the output is
    checked against the f64-accumulated oracle on a 256-token x 64-column sample spread over the tile grid, and must be
    finite everywhere.  Bound: 1e-2 relative (bf16 operands, f32 accumulate — BASELINE.json north_star)."""
    import torch
    from dllm_b200 import QWeight, PATH_AUTO
    K, N = KN
    M = 8192
    g = torch.Generator(device="cuda").manual_seed(K + N)
    w = torch.randn(K, N, device="cuda", generator=g) * (1.0 / K ** 0.5)
    x = torch.randn(M, K, device="cuda", generator=g)
    y = torch.empty(M, N, device="cuda")
    torch.cuda.synchronize()
    qw = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, 4, 128)
    qw.forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_AUTO)
    ctx.sync()
    codes, scales, zps = qw.export()
    oc, os_, oz = O.quantize_weight_grouped(w.cpu().numpy(), 4, 128)
    assert np.array_equal(codes, oc) and np.array_equal(scales.view(np.uint32), os_.view(np.uint32)) and np.array_equal(zps, oz)
    rng = np.random.default_rng(7)
    toks = np.unique(np.concatenate([rng.integers(0, M, 250), [0, 127, 128, 223, 224, 255, 256, M - 1]]))
    cols = np.unique(np.concatenate([rng.integers(0, N, 60), [0, 127, 128, 255, 256, N - 1]]))
    wd = O.dequantize_weight_grouped(oc, os_, oz, 128)[:, cols].astype(np.float64)
    y64 = x.cpu().numpy()[toks].astype(np.float64) @ wd
    ys = y.cpu().numpy()
    assert np.all(np.isfinite(ys))
    got = ys[np.ix_(toks, cols)]
    assert np.linalg.norm(got - y64) <= 1e-2 * np.linalg.norm(y64)
    assert np.abs(got - y64).max() <= 2e-2 * np.abs(y64).max()
    qw.close()


@pytest.mark.parametrize("bits,group,bias", [(4, 128, True), (2, 128, False), (8, 64, True), (3, 0, False), (1, 0, True)])
def test_qweight_container_round_trip(ctx, O, tmp_path, bits, group, bias):
    """Packed-weights container "DLLMQW01" (SURVEY.md 8f-3): serialize -> deserialize and save -> load give back the same
    codes / scales / zero-points (bit for bit, equal to the oracle's quantization) and the same forward result; the header
    carries K, N, bits, group; a flipped byte or a truncated buffer is refused."""
    import struct
    from dllm_b200 import QWeight, PATH_SIMT, DllmError
    from dllm_b200 import _lib as L
    rng = np.random.default_rng(bits * 7 + group)
    K, N = 256, 200                                     # ragged N: codes do not fill whole tiles
    w = make_w(rng, K, N)
    b = rng.standard_normal(N).astype(F) if bias else None
    qw = QWeight.quantize(ctx, w, bits, group, b)
    blob = qw.serialize()
    magic, version, hbits, hK, hN, hgroup, scheme, has_bias, codes_bytes, _ = struct.unpack("<8sIIQQQIIQQ", blob[:64])
    pw = 1 if bits <= 1 else 2 if bits <= 2 else 4 if bits <= 4 else 8
    assert (magic, version, hbits, hK, hN, hgroup, scheme, has_bias) == (b"DLLMQW01", 1, bits, K, N, group, 0, int(bias))
    assert codes_bytes == (K * N * pw + 7) // 8
    G = K // group if group else 1
    assert len(blob) == 64 + codes_bytes + 2 * 4 * (G * N if group else 1) + (4 * N if bias else 0) + 4
    # the packed codes in the container are dllm_pack's layout of the oracle's codes
    if group:
        oc, os_, oz = O.quantize_weight_grouped(w, bits, group)
    else:
        oc, s1, z1 = O.quantize_tensor(w.ravel(), bits)
        oc, os_, oz = oc.reshape(K, N), np.array([[s1]], F), np.array([[z1]], F)
    assert np.array_equal(np.frombuffer(blob[64:64 + codes_bytes], np.uint8), O.pack(oc.ravel(), pw))
    x = rng.standard_normal((5, K)).astype(F)
    y0 = qw.forward(x, PATH_SIMT)
    path = str(tmp_path / "w.dllmqw")
    qw.save(path)
    for q2 in (QWeight.deserialize(ctx, blob), QWeight.load(ctx, path)):
        assert (q2.K, q2.N, q2.bits, q2.group) == (K, N, bits, group)
        c2, s2, z2 = q2.export()
        assert np.array_equal(c2, oc) and np.array_equal(s2.view(np.uint32), os_.view(np.uint32)) and np.array_equal(z2, oz)
        assert np.array_equal(q2.forward(x, PATH_SIMT).view(np.uint32), y0.view(np.uint32))
        q2.close()
    bad = bytearray(blob)
    bad[70] ^= 0x10
    for data in (bytes(bad), blob[:-5], b"NOTAFILE" + blob[8:]):
        with pytest.raises(DllmError) as ei:
            QWeight.deserialize(ctx, data)
        assert ei.value.code in (L.ERR_SERIALIZATION, L.ERR_INVALID_DATA_FORMAT)
    with pytest.raises(DllmError) as ei:
        QWeight.load(ctx, str(tmp_path / "missing.dllmqw"))
    assert ei.value.code == L.ERR_IO
    qw.close()
