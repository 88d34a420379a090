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


def test_kv_cache_accounting_and_eviction_policy():
    """lib.rs:988-1084 on the host: the accounting quirks (f32 size charged before quantization, saturating growth of an
    updated entry) and the eviction order (largest entry first, despite the function's name) — no GPU involved: the
    entries are stand-ins that report the reference's memory_usage formula (:279-302)."""
    from dllm_b200.diffuse_llm import DiffuseLLM, DiffusionConfig

    class Entry:
        def __init__(self, k, v, pre, dec):
            self.k, self.bits = k, (pre, dec)

        def memory_usage(self):
            return sum(2 * ((self.k.size * b + 7) // 8) for b in self.bits if b > 0) or self.k.size * 8

        def update(self, k, v):
            self.k = k

    class LLM(DiffuseLLM):
        def _new_entry(self, keys, values, pre, dec):
            return Entry(keys, values, pre, dec)

    cfg = DiffusionConfig(hidden_size=8, num_layers=2, kv_quant_bits=4, max_cache_size=3000)
    llm = LLM(cfg, ctx=object())
    k = lambda seq: np.zeros((2, seq, 8), np.float32)
    llm.update_kv_cache("a", k(4), k(4))                       # 64 elements: quantized size 2 copies x 2 tensors x 32 B = 128
    assert llm.kv_cache_memory_usage() == 128
    llm.update_kv_cache("b", k(16), k(16))                     # 256 elements -> 512 B
    assert llm.kv_cache_memory_usage() == 640
    llm.update_kv_cache("a", k(8), k(8))                       # in place: grows by max(0, 128*4*2 - 128) = 896
    assert llm.kv_cache_memory_usage() == 640 + 896 and llm.kv_cache["a"].k.shape[1] == 8
    # next insert: 1536 + 32*4*2*... = over the 3000-byte budget -> evict the LARGEST entry ("b", 512 B), not the oldest
    llm.update_kv_cache("c", k(12), k(12))                     # entry_size = 192*4*2 = 1536 -> new_usage 3072 > 3000
    assert "b" not in llm.kv_cache and "a" in llm.kv_cache and "c" in llm.kv_cache
    assert llm.kv_cache_memory_usage() == 1536 - 512 + 384     # 'c' is charged its quantized size (192 el -> 384 B)
    llm.clear_kv_cache()
    assert llm.kv_cache == {} and llm.kv_cache_memory_usage() == 0
    cfg.use_kv_cache = False
    llm.update_kv_cache("z", k(4), k(4))
    assert llm.kv_cache == {}
