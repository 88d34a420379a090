"""Mirror of diffuse_llm_rs::quantization (diffuse-llm-rs/src/quantization.rs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .runtime import Context, default_context


def quantize_tensor(data, bits: int, ctx: Context | None = None):
    """quantize_tensor(&[f32], bits) -> (Vec<u8>, scale, zero_point)   quantization.rs:38-68.
    bits outside 1..=8 panics in the reference -> InvalidParams here."""
    return (ctx or default_context()).quantize_tensor(data, bits)


def dequantize_tensor(data, scale, zero_point, ctx: Context | None = None):
    """dequantize_tensor(&[u8], scale, zp) -> Vec<f32>   quantization.rs:81-85"""
    return (ctx or default_context()).dequantize_tensor(data, scale, zero_point)


class QuantizedTensor:
    """quantization.rs:89-125"""

    def __init__(self, data, shape, scale, zero_point, bits, ctx: Context | None = None):
        self.data = np.ascontiguousarray(data, np.uint8)
        self.shape = list(shape)
        self.scale = np.float32(scale)
        self.zero_point = np.float32(zero_point)
        self.bits = int(bits)
        self._ctx = ctx

    def dequantize(self):
        return dequantize_tensor(self.data, self.scale, self.zero_point, self._ctx)

    def compression_ratio(self):
        numel = int(np.prod(self.shape)) if len(self.shape) else 1
        return np.float32(L.lib().dllm_compression_ratio(numel, self.data.size, self.bits))


class QuantizedKVCacheEntry:
    """quantization.rs:129-176: K and V [layers, seq, hidden] each quantized as ONE tensor.
    The codes stay resident (bit-packed) in HBM; `keys`/`values` materialise the reference view."""

    def __init__(self, keys, values, bits: int, ctx: Context | None = None, scheme: int = L.KV_TENSOR_B):
        self._ctx = ctx or default_context()
        keys = np.ascontiguousarray(keys, np.float32)      # as_slice().unwrap(): contiguous, :143
        values = np.ascontiguousarray(values, np.float32)
        assert keys.ndim == 3 and keys.shape == values.shape
        self.shape = list(keys.shape)
        self.bits = int(bits)
        self.seq_len = keys.shape[1]
        h = C.c_void_p()
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_kv_quantize(self._ctx.h, keys.ctypes.data, values.ctypes.data,
                                                          keys.shape[0], keys.shape[1], keys.shape[2], bits, scheme,
                                                          C.byref(h)))
        self.h = h
        self.scheme = scheme

    @classmethod
    def with_capacity(cls, layers: int, capacity: int, hidden: int, bits: int, scheme: int = L.KV_ROW_D, ctx: Context | None = None):
        """An empty entry with room for `capacity` tokens per layer (per-token or fixed-scale schemes): `append` quantizes
        only the new tokens, where the reference's KVCacheEntry::update (lib.rs:246-276) re-quantizes the whole cache."""
        self = cls.__new__(cls)
        self._ctx = ctx or default_context()
        self.shape = [layers, 0, hidden]
        self.bits = int(bits)
        self.seq_len = 0
        self.scheme = scheme
        h = C.c_void_p()
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_kv_create(self._ctx.h, layers, capacity, hidden, bits, scheme, C.byref(h)))
        self.h = h
        return self

    def append(self, keys_new, values_new):
        """keys_new / values_new: [layers, t_new, hidden]; they land after the tokens every layer already holds."""
        keys_new = np.ascontiguousarray(keys_new, np.float32)
        values_new = np.ascontiguousarray(values_new, np.float32)
        assert keys_new.ndim == 3 and keys_new.shape == values_new.shape
        assert keys_new.shape[0] == self.shape[0] and keys_new.shape[2] == self.shape[2]
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_kv_append(self._ctx.h, self.h, keys_new.ctypes.data, values_new.ctypes.data,
                                                        keys_new.shape[1]))
            self.seq_len = int(self._ctx._lib.dllm_kv_seq_len(self.h))
        self.shape[1] = self.seq_len

    def _export(self):
        n = int(np.prod(self.shape))
        rows = self.shape[0] * self.shape[1]
        np_ = rows if self.scheme == L.KV_ROW_D else 1
        kc, vc = np.empty(n, np.uint8), np.empty(n, np.uint8)
        ks, kz, vs, vz = (np.empty(np_, np.float32) for _ in range(4))
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_kv_export(self._ctx.h, self.h, kc.ctypes.data, vc.ctypes.data,
                                                        ks.ctypes.data, kz.ctypes.data, vs.ctypes.data, vz.ctypes.data))
        return kc, vc, ks, kz, vs, vz

    @property
    def keys(self) -> QuantizedTensor:
        kc, _, ks, kz, _, _ = self._export()
        return QuantizedTensor(kc, self.shape, ks[0], kz[0], self.bits, self._ctx)

    @property
    def values(self) -> QuantizedTensor:
        _, vc, _, _, vs, vz = self._export()
        return QuantizedTensor(vc, self.shape, vs[0], vz[0], self.bits, self._ctx)

    def _deq(self, want_k, want_v):
        k = np.empty(self.shape, np.float32) if want_k else None
        v = np.empty(self.shape, np.float32) if want_v else None
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_kv_dequantize(self._ctx.h, self.h, k.ctypes.data if want_k else None,
                                                            v.ctypes.data if want_v else None))
        return k, v

    def dequantize_keys(self):
        return self._deq(True, False)[0]

    def dequantize_values(self):
        return self._deq(False, True)[1]

    def memory_usage(self) -> int:
        return int(self._ctx._lib.dllm_kv_memory_usage(self.h))

    def close(self):
        if getattr(self, "h", None):
            self._ctx.sync()
            self._ctx._lib.dllm_kv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AdaptiveQuantizer:
    """quantization.rs:179-235.  The reference feeds a CKMS(0.01) quantile sketch (un-vendored
    `quantiles` crate) and queries q=0 / q=1 only; the published algorithm answers those two
    queries with the exact extremes it has seen (restated and tested on the checker's side,
    tests/test_oracle_golden.py), so the statistics here are a running exact
    min/max — the device reduction per chunk — and the parameters equal the sketch's bit for
    bit (tests/test_gpu_quantizers.py).  The code step runs on the GPU."""

    def __init__(self, bits: int, target_ratio: float, ctx: Context | None = None):
        self.bits, self.target_ratio = int(bits), float(target_ratio)
        self._ctx = ctx or default_context()
        self._min = None
        self._max = None

    def update_stats(self, data):
        if np.size(data) == 0:
            return
        mn, mx = self._ctx.minmax(data)
        self._min = mn if self._min is None else min(self._min, mn)
        self._max = mx if self._max is None else max(self._max, mx)

    def compute_params(self):
        F = np.float32
        mn = F(0.0) if self._min is None else F(self._min)      # unwrap_or(0.0)  :208
        mx = F(1.0) if self._max is None else F(self._max)      # unwrap_or(1.0)  :209
        q_max = F(F(1 << self.bits) - F(1))
        with np.errstate(all="ignore"):
            scale = F(F(mx - mn) / q_max)                       # :213
            r = F(-mn) / scale
            t = np.trunc(r)
            r = t + np.copysign(F(1), r) if abs(r - t) >= 0.5 else t   # f32::round
            zp = F(min(max(F(r), F(0)), q_max))                 # .round().clamp(0, q_max)  :214
        return scale, zp

    def quantize(self, data):
        scale, zp = self.compute_params()
        return self._ctx.quantize_codes(data, self.bits, scale, zp), scale, zp
