"""Host-side arithmetic of the library (no GPU needed) against the oracle: compression ratio,
calibration parameters, beta schedules, progressive decode bits, fixed C scales, packed sizes."""
import ctypes as C

import numpy as np
import pytest

from dllm_b200 import _lib as L
from oracle import pyoracle as O

F = np.float32


def beq(a, b):
    return np.array_equal(np.asarray(a, F).view(np.uint32), np.asarray(b, F).view(np.uint32))


def test_compression_ratio():
    for numel, dlen, bits in [(4, 4, 4), (1000, 1000, 1), (7, 7, 3), (1 << 20, 1 << 20, 8)]:
        assert beq(L.lib().dllm_compression_ratio(numel, dlen, bits), O.compression_ratio(numel, dlen, bits))
    assert L.lib().dllm_compression_ratio(4, 4, 4) == 8.0      # quantization.rs:254-265


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("bits", [1, 2, 4, 8])
def test_calibrate_params(bits, sym):
    rng = np.random.default_rng(bits)
    for _ in range(50):
        a, b = np.sort(rng.standard_normal(2).astype(F) * F(10))
        s, z = C.c_float(), C.c_int32()
        assert L.lib().dllm_calibrate_params(float(a), float(b), 5, bits, int(sym), C.byref(s), C.byref(z)) == 0
        es, ez = O.calibrate_params(a, b, 5, bits, sym)
        assert beq(s.value, es) and z.value == ez
    s, z = C.c_float(), C.c_int32()
    assert L.lib().dllm_calibrate_params(0.0, 1.0, 0, 8, 0, C.byref(s), C.byref(z)) == L.ERR_CALIBRATION_REQUIRED
    assert L.lib().dllm_calibrate_params(1.0, 6.0, 6, 8, 0, C.byref(s), C.byref(z)) == 0
    assert beq(s.value, F(0.019607844)) and z.value == -51     # calibrate.rs:123-132 as the code computes it


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("T", [1, 2, 50, 1000])
def test_beta_schedule(kind, T):
    betas = np.empty(T, F)
    assert L.lib().dllm_beta_schedule(kind, T, 1e-4, 0.02, betas.ctypes.data) == 0
    exp = O.beta_schedule(kind, T)
    assert np.array_equal(betas.view(np.uint32), exp.view(np.uint32))   # NaN-safe bit compare (T=1: 0/0)
    assert L.lib().dllm_beta_schedule(7, T, 1e-4, 0.02, betas.ctypes.data) == L.ERR_INVALID_PARAMS
    assert L.lib().dllm_beta_schedule(kind, 0, 1e-4, 0.02, betas.ctypes.data) == L.ERR_INVALID_PARAMS


def test_progressive_bits():
    for steps in (2, 10, 64, 1000):
        for t in range(steps):
            pre = C.c_int32()
            b = L.lib().dllm_progressive_bits(steps, t, 4, 2, C.byref(pre))
            eb, ep = O.progressive_bits(steps, t, 4, 2)
            assert (b, bool(pre.value)) == (eb, ep)


def test_bitquantizer_scale_and_packed_len():
    for bits in (1, 2, 4, 6, 8, 16):
        assert beq(L.lib().dllm_bitquantizer_scale(bits), O.bitquantizer_scale_c(bits))
    for n in (0, 1, 7, 8, 9, 4097):
        for bits in (1, 2, 4, 8):
            assert L.lib().dllm_packed_len(n, bits) == O.packed_len(n, bits) == (n * bits + 7) // 8
