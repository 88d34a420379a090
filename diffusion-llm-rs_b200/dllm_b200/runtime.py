"""Context + thin numpy front-ends of the C ABI.  Host-pointer calls take/return numpy arrays
(the reference's `&[f32]` / `Vec<u8>`); `*_dev` calls take raw device addresses (ints, e.g.
`torch.Tensor.data_ptr()`), enqueue on the context's stream and do not synchronise."""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib as L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Context:
    """One CUDA device + one stream (include/dllm_b200.h `dllm_ctx`).  Not shared between threads:
    a lock serialises callers, which is how the Rust wrapper satisfies `Send + Sync`."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = L.lib()
        h = C.c_void_p()
        if stream is None:
            rc = self._lib.dllm_ctx_create(device, C.byref(h))
        else:
            rc = self._lib.dllm_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h))
        if rc != L.OK:
            raise L._ERR.get(rc, L.DllmError)(rc, "dllm_ctx_create failed (no sm_100 CUDA device?) — no CPU fallback")
        self.h = h
        self.device = device
        self.lock = threading.RLock()

    def close(self):
        if getattr(self, "h", None):
            self._lib.dllm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        L.check(rc, self.h)

    # ---- plumbing ----
    def sync(self):
        self._ck(self._lib.dllm_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self._lib.dllm_ctx_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(self._lib.dllm_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(self._lib.dllm_sm_count(self.h))

    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self._lib.dllm_malloc(self.h, nbytes, C.byref(p)))
        return int(p.value)

    def free(self, dptr: int):
        self._ck(self._lib.dllm_free(self.h, C.c_void_p(dptr)))

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self._ck(self._lib.dllm_memcpy_h2d(self.h, C.c_void_p(dptr), _ptr(arr), arr.nbytes))
        self.sync()

    def d2h(self, dptr: int, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype)
        self._ck(self._lib.dllm_memcpy_d2h(self.h, _ptr(out), C.c_void_p(dptr), out.nbytes))
        return out

    @property
    def copy_bytes(self):
        """(host->device, device->host) bytes this context has moved so far"""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.dllm_copy_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def graph_replays(self) -> int:
        return int(self._lib.dllm_graph_replay_count(self.h))

    # ---- "dllm_noise v1": the seeded loop's counter-based N(0,1) generator (csrc/noise.cuh) ----
    def noise_fill(self, seed: int, stream: int, n: int, first: int = 0) -> np.ndarray:
        out = np.empty(int(n), np.float32)
        with self.lock:
            self._ck(self._lib.dllm_noise_fill(self.h, seed, stream, first, int(n), _ptr(out)))
        return out

    def noise_fill_dev(self, seed: int, stream: int, n: int, out_dev: int, first: int = 0):
        self._ck(self._lib.dllm_noise_fill_dev(self.h, seed, stream, first, int(n), out_dev))

    # ---- quantizer B ----
    def quantize_tensor(self, x, bits):
        x = _f32(x).ravel()
        codes = np.empty(x.size, np.uint8)
        s, z = C.c_float(), C.c_float()
        with self.lock:
            self._ck(self._lib.dllm_quantize_tensor(self.h, _ptr(x), x.size, bits, _ptr(codes), C.byref(s), C.byref(z)))
        return codes, np.float32(s.value), np.float32(z.value)

    def dequantize_tensor(self, codes, scale, zp):
        codes = _u8(codes).ravel()
        out = np.empty(codes.size, np.float32)
        with self.lock:
            self._ck(self._lib.dllm_dequantize_tensor(self.h, _ptr(codes), codes.size, float(scale), float(zp), _ptr(out)))
        return out

    def quantize_codes(self, x, bits, scale, zp):
        x = _f32(x).ravel()
        codes = np.empty(x.size, np.uint8)
        with self.lock:
            self._ck(self._lib.dllm_quantize_codes(self.h, _ptr(x), x.size, bits, float(scale), float(zp), _ptr(codes)))
        return codes

    def quantize_tensor_dev(self, x_dev, n, bits, packed, codes_dev, params_dev):
        self._ck(self._lib.dllm_quantize_tensor_dev(self.h, x_dev, n, bits, int(packed), codes_dev, params_dev))

    def dequantize_tensor_dev(self, codes_dev, n, bits, packed, params_dev, out_dev):
        self._ck(self._lib.dllm_dequantize_tensor_dev(self.h, codes_dev, n, bits, int(packed), params_dev, out_dev))

    # ---- quantizer A ----
    def quantize_a(self, x, qtype, scale=1.0, zero_point=0):
        x = _f32(x).ravel()
        codes = np.empty(x.size, np.uint8)
        with self.lock:
            self._ck(self._lib.dllm_quantize_a(self.h, _ptr(x), x.size, qtype, float(scale), int(zero_point), _ptr(codes)))
        return codes

    def dequantize_a(self, codes, scale=1.0, zero_point=0):
        codes = _u8(codes).ravel()
        out = np.empty(codes.size, np.float32)
        with self.lock:
            self._ck(self._lib.dllm_dequantize_a(self.h, _ptr(codes), codes.size, float(scale), int(zero_point), _ptr(out)))
        return out

    def minmax(self, x):
        x = _f32(x).ravel()
        mn, mx = C.c_float(), C.c_float()
        with self.lock:
            self._ck(self._lib.dllm_minmax(self.h, _ptr(x), x.size, C.byref(mn), C.byref(mx)))
        return np.float32(mn.value), np.float32(mx.value)

    # ---- quantizers C / D ----
    def quantize_c(self, x, bits, scale, zp=0.0):
        x = _f32(x).ravel()
        codes = np.empty(x.size, np.uint8)
        with self.lock:
            self._ck(self._lib.dllm_quantize_c(self.h, _ptr(x), x.size, bits, float(scale), float(zp), _ptr(codes)))
        return codes

    def dequantize_cd(self, codes, scale, zp):
        codes = _u8(codes).ravel()
        out = np.empty(codes.size, np.float32)
        with self.lock:
            self._ck(self._lib.dllm_dequantize_cd(self.h, _ptr(codes), codes.size, float(scale), float(zp), _ptr(out)))
        return out

    def kvquant_quantize_vectors(self, emb, cfg_bits, bits):
        emb = _f32(emb)
        nvec = emb.shape[0]
        per = int(np.prod(emb.shape[1:])) if emb.ndim > 1 else 1
        cfg, b = _u8(cfg_bits), _u8(bits)
        codes = np.zeros(emb.shape, np.uint8)
        with self.lock:
            self._ck(self._lib.dllm_kvquant_quantize_vectors(self.h, _ptr(emb), nvec, per, _ptr(cfg), cfg.size,
                                                             _ptr(b), b.size, _ptr(codes)))
        return codes

    def quantize_d_rows(self, x, bits):
        x = _f32(x)
        rows, dim = x.shape
        b = _u8(np.atleast_1d(bits))
        codes = np.empty((rows, dim), np.uint8)
        scales = np.empty(rows, np.float32)
        zps = np.empty(rows, np.float32)
        with self.lock:
            self._ck(self._lib.dllm_quantize_d_rows(self.h, _ptr(x), rows, dim, _ptr(b), b.size, _ptr(codes),
                                                    _ptr(scales), _ptr(zps)))
        return codes, scales, zps

    def dequantize_d_rows(self, codes, scales, zps):
        codes = _u8(codes)
        rows, dim = codes.shape
        scales, zps = _f32(scales), _f32(zps)
        out = np.empty((rows, dim), np.float32)
        with self.lock:
            self._ck(self._lib.dllm_dequantize_d_rows(self.h, _ptr(codes), rows, dim, _ptr(scales), _ptr(zps), _ptr(out)))
        return out

    def quantize_d_rows_dev(self, x_dev, rows, dim, bits, packed, codes_dev, scales_dev, zps_dev):
        self._ck(self._lib.dllm_quantize_d_rows_dev(self.h, x_dev, rows, dim, bits, int(packed), codes_dev, scales_dev, zps_dev))

    def dequantize_d_rows_dev(self, codes_dev, rows, dim, bits, packed, scales_dev, zps_dev, out_dev):
        self._ck(self._lib.dllm_dequantize_d_rows_dev(self.h, codes_dev, rows, dim, bits, int(packed), scales_dev, zps_dev, out_dev))

    # ---- pack / unpack ----
    def pack(self, codes, bits):
        codes = _u8(codes).ravel()
        out = np.empty(int(self._lib.dllm_packed_len(codes.size, bits)) if bits in (1, 2, 4, 8) else 0, np.uint8)
        with self.lock:
            self._ck(self._lib.dllm_pack(self.h, _ptr(codes), codes.size, bits, _ptr(out)))
        return out

    def unpack(self, packed, n, bits):
        packed = _u8(packed).ravel()
        out = np.empty(n, np.uint8)
        with self.lock:
            self._ck(self._lib.dllm_unpack(self.h, _ptr(packed), n, bits, _ptr(out)))
        return out

    def pack_dev(self, codes_dev, n, bits, packed_dev):
        self._ck(self._lib.dllm_pack_dev(self.h, codes_dev, n, bits, packed_dev))

    def unpack_dev(self, packed_dev, n, bits, codes_dev):
        self._ck(self._lib.dllm_unpack_dev(self.h, packed_dev, n, bits, codes_dev))


class QWeight:
    """A quantized [K,N] weight resident in HBM (`dllm_qweight`)."""

    def __init__(self, ctx: Context, handle, K, N, bits, group):
        self.ctx, self.h, self.K, self.N, self.bits, self.group = ctx, handle, K, N, bits, group

    @classmethod
    def quantize(cls, ctx: Context, w, bits, group=128, bias=None):
        w = _f32(w)
        K, N = w.shape
        b = _f32(bias) if bias is not None else None
        h = C.c_void_p()
        with ctx.lock:
            ctx._ck(ctx._lib.dllm_qweight_quantize(ctx.h, _ptr(w), K, N, bits, group, _ptr(b), C.byref(h)))
        return cls(ctx, h, K, N, bits, group)

    @classmethod
    def quantize_dev(cls, ctx: Context, w_dev: int, K, N, bits, group=128, bias_dev=None):
        h = C.c_void_p()
        ctx._ck(ctx._lib.dllm_qweight_quantize_dev(ctx.h, w_dev, K, N, bits, group, bias_dev, C.byref(h)))
        return cls(ctx, h, K, N, bits, group)

    @classmethod
    def from_codes(cls, ctx: Context, codes, scales, zps, bits, group=128, bias=None):
        codes = _u8(codes)
        K, N = codes.shape
        scales, zps = _f32(scales), _f32(zps)
        b = _f32(bias) if bias is not None else None
        h = C.c_void_p()
        with ctx.lock:
            ctx._ck(ctx._lib.dllm_qweight_from_codes(ctx.h, _ptr(codes), _ptr(scales), _ptr(zps), K, N, bits, group,
                                                     _ptr(b), C.byref(h)))
        return cls(ctx, h, K, N, bits, group)

    # ---- packed-weights container "DLLMQW01" (include/dllm_b200.h) ----
    def serialize(self) -> bytes:
        n = int(self.ctx._lib.dllm_qweight_serialized_size(self.h))
        buf = (C.c_uint8 * n)()
        wr = C.c_size_t()
        with self.ctx.lock:
            self.ctx._ck(self.ctx._lib.dllm_qweight_serialize(self.ctx.h, self.h, buf, n, C.byref(wr)))
        return bytes(buf[:wr.value])

    @classmethod
    def deserialize(cls, ctx: "Context", data: bytes):
        arr = (C.c_uint8 * len(data)).from_buffer_copy(data)
        h = C.c_void_p()
        with ctx.lock:
            ctx._ck(ctx._lib.dllm_qweight_deserialize(ctx.h, arr, len(data), C.byref(h)))
        return cls._from_handle(ctx, h)

    def save(self, path: str):
        with self.ctx.lock:
            self.ctx._ck(self.ctx._lib.dllm_qweight_save(self.ctx.h, self.h, path.encode()))

    @classmethod
    def load(cls, ctx: "Context", path: str):
        h = C.c_void_p()
        with ctx.lock:
            ctx._ck(ctx._lib.dllm_qweight_load(ctx.h, path.encode(), C.byref(h)))
        return cls._from_handle(ctx, h)

    @classmethod
    def _from_handle(cls, ctx, h):
        K, N, g, pb = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        bits = C.c_uint8()
        ctx._lib.dllm_qweight_info(h, C.byref(K), C.byref(N), C.byref(bits), C.byref(g), C.byref(pb))
        return cls(ctx, h, int(K.value), int(N.value), int(bits.value), int(g.value))

    def export(self):
        G = 1 if self.group == 0 else self.K // self.group
        codes = np.empty((self.K, self.N), np.uint8)
        ncol = 1 if self.group == 0 else self.N
        scales = np.empty((G, ncol), np.float32)
        zps = np.empty((G, ncol), np.float32)
        with self.ctx.lock:
            self.ctx._ck(self.ctx._lib.dllm_qweight_export(self.ctx.h, self.h, _ptr(codes), _ptr(scales), _ptr(zps)))
        return codes, scales, zps

    @property
    def packed_bytes(self) -> int:
        pb = C.c_size_t()
        self.ctx._lib.dllm_qweight_info(self.h, None, None, None, None, C.byref(pb))
        return int(pb.value)

    def forward(self, x, path=L.PATH_AUTO):
        x = _f32(x)
        M, K = x.shape
        assert K == self.K
        y = np.empty((M, self.N), np.float32)
        with self.ctx.lock:
            self.ctx._ck(self.ctx._lib.dllm_qlinear_forward(self.ctx.h, self.h, _ptr(x), M, _ptr(y), path))
        return y

    def forward_dev(self, x_dev: int, M: int, y_dev: int, path=L.PATH_AUTO):
        self.ctx._ck(self.ctx._lib.dllm_qlinear_forward_dev(self.ctx.h, self.h, x_dev, M, y_dev, path))

    def forward_i8(self, xq):
        """Exact integer linear for a per-tensor quantized weight: int8 xq[M,K] -> int32 y[M,N] = sum_k xq (q - zp)
        (tcgen05 kind::i8; dequantize_tensor composed with the matmul is tensor_scale * x_scale * y)."""
        xq = np.ascontiguousarray(xq, dtype=np.int8)
        M, K = xq.shape
        assert K == self.K
        y = np.empty((M, self.N), np.int32)
        with self.ctx.lock:
            self.ctx._ck(self.ctx._lib.dllm_qlinear_forward_i8(self.ctx.h, self.h, _ptr(xq), M, _ptr(y)))
        return y

    def forward_i8_dev(self, xq_dev: int, M: int, y_dev: int):
        self.ctx._ck(self.ctx._lib.dllm_qlinear_forward_i8_dev(self.ctx.h, self.h, xq_dev, M, y_dev))

    def close(self):
        if self.h:
            self.ctx.sync()
            self.ctx._lib.dllm_qweight_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dequant_matmul(ctx: Context, codes, scales, zps, x, bits, group=128, bias=None, path=L.PATH_AUTO):
    """`quantization` crate extension: y = x · dequant(codes) + b (BASELINE.json north_star)."""
    codes = _u8(codes)
    K, N = codes.shape
    x = _f32(x)
    M = x.shape[0]
    scales, zps = _f32(scales), _f32(zps)
    b = _f32(bias) if bias is not None else None
    y = np.empty((M, N), np.float32)
    with ctx.lock:
        ctx._ck(ctx._lib.dllm_dequant_matmul(ctx.h, _ptr(codes), _ptr(scales), _ptr(zps), K, N, bits, group, _ptr(b),
                                             _ptr(x), M, _ptr(y), path))
    return y


_default_ctx = None


def default_context() -> Context:
    """Process-wide context on device LOCAL_RANK (or 0)."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx
