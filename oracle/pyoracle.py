"""ctypes/numpy front-end of the CPU oracle (oracle/dllm_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdllm_oracle.so")

QT_INT8, QT_INT4, QT_BINARY, QT_FLOAT8 = 0, 1, 2, 3
BETA_LINEAR, BETA_QUADRATIC, BETA_COSINE = 0, 1, 2
OK, ERR_INVALID_PARAMS, ERR_SHAPE, ERR_INDEX = 0, 1, 3, 8


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dllm_oracle.c")
    hdr = os.path.join(_HERE, "dllm_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_compression_ratio.restype = C.c_float
        _lib.orc_bitquantizer_scale_c.restype = C.c_float
        _lib.orc_packed_len.restype = C.c_size_t
        _lib.orc_progressive_bits.restype = C.c_uint8
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle status {code}")
        self.code = code


def _chk(rc):
    if rc != 0:
        raise OracleError(rc)


# ---- quantizer B ----
def quantize_tensor(x, bits):
    x = _f32(x).ravel()
    codes = np.empty(x.size, np.uint8)
    s, z = C.c_float(), C.c_float()
    _chk(lib().orc_quantize_tensor(_p(x, C.c_float), C.c_size_t(x.size), C.c_uint8(bits),
                                   _p(codes, C.c_uint8), C.byref(s), C.byref(z)))
    return codes, np.float32(s.value), np.float32(z.value)


def dequantize_tensor(codes, scale, zp):
    codes = _u8(codes).ravel()
    out = np.empty(codes.size, np.float32)
    lib().orc_dequantize_tensor(_p(codes, C.c_uint8), C.c_size_t(codes.size), C.c_float(scale),
                                C.c_float(zp), _p(out, C.c_float))
    return out


def quantize_codes_b(x, bits, scale, zp):
    x = _f32(x).ravel()
    codes = np.empty(x.size, np.uint8)
    lib().orc_quantize_codes_b(_p(x, C.c_float), C.c_size_t(x.size), C.c_uint8(bits),
                               C.c_float(scale), C.c_float(zp), _p(codes, C.c_uint8))
    return codes


def compression_ratio(numel, data_len, bits):
    return np.float32(lib().orc_compression_ratio(C.c_size_t(numel), C.c_size_t(data_len),
                                                  C.c_uint8(bits)))


def quantize_weight_grouped(w, bits, group=128):
    w = _f32(w)
    K, N = w.shape
    codes = np.empty((K, N), np.uint8)
    scales = np.empty((K // group, N), np.float32)
    zps = np.empty((K // group, N), np.float32)
    _chk(lib().orc_quantize_weight_grouped(_p(w, C.c_float), C.c_size_t(K), C.c_size_t(N),
                                           C.c_uint8(bits), C.c_size_t(group),
                                           _p(codes, C.c_uint8), _p(scales, C.c_float),
                                           _p(zps, C.c_float)))
    return codes, scales, zps


def dequantize_weight_grouped(codes, scales, zps, group=128):
    codes = _u8(codes)
    K, N = codes.shape
    scales, zps = _f32(scales), _f32(zps)
    w = np.empty((K, N), np.float32)
    lib().orc_dequantize_weight_grouped(_p(codes, C.c_uint8), C.c_size_t(K), C.c_size_t(N),
                                        C.c_size_t(group), _p(scales, C.c_float),
                                        _p(zps, C.c_float), _p(w, C.c_float))
    return w


# ---- quantizer A ----
def quantize_a(x, qtype, scale=1.0, zero_point=0):
    x = _f32(x).ravel()
    codes = np.empty(x.size, np.uint8)
    _chk(lib().orc_quantize_a(_p(x, C.c_float), C.c_size_t(x.size), C.c_int32(qtype),
                              C.c_float(scale), C.c_int32(zero_point), _p(codes, C.c_uint8)))
    return codes


def dequantize_a(codes, scale=1.0, zero_point=0):
    codes = _u8(codes).ravel()
    out = np.empty(codes.size, np.float32)
    lib().orc_dequantize_a(_p(codes, C.c_uint8), C.c_size_t(codes.size), C.c_float(scale),
                           C.c_int32(zero_point), _p(out, C.c_float))
    return out


def calibrate_params(mn, mx, total, bits, symmetric):
    s, z = C.c_float(), C.c_int32()
    _chk(lib().orc_calibrate_params(C.c_float(mn), C.c_float(mx), C.c_size_t(total),
                                    C.c_uint8(bits), C.c_int32(int(symmetric)),
                                    C.byref(s), C.byref(z)))
    return np.float32(s.value), int(z.value)


# ---- quantizer C ----
def bitquantizer_scale_c(bits):
    return np.float32(lib().orc_bitquantizer_scale_c(C.c_uint8(bits)))


def quantize_c(x, bits, scale, zp=0.0):
    x = _f32(x).ravel()
    codes = np.empty(x.size, np.uint8)
    _chk(lib().orc_quantize_c(_p(x, C.c_float), C.c_size_t(x.size), C.c_uint8(bits),
                              C.c_float(scale), C.c_float(zp), _p(codes, C.c_uint8)))
    return codes


def dequantize_cd(codes, scale, zp):
    codes = _u8(codes).ravel()
    out = np.empty(codes.size, np.float32)
    lib().orc_dequantize_cd(_p(codes, C.c_uint8), C.c_size_t(codes.size), C.c_float(scale),
                            C.c_float(zp), _p(out, C.c_float))
    return out


def kvquant_quantize_vectors(emb, cfg_bits, bits):
    emb = _f32(emb)
    nvec = emb.shape[0]
    per = int(np.prod(emb.shape[1:])) if emb.ndim > 1 else 1
    cfg = _u8(cfg_bits)
    b = _u8(bits)
    codes = np.zeros(emb.shape, np.uint8)
    _chk(lib().orc_kvquant_quantize_vectors(_p(emb, C.c_float), C.c_size_t(nvec), C.c_size_t(per),
                                            _p(cfg, C.c_uint8), C.c_size_t(cfg.size),
                                            _p(b, C.c_uint8), C.c_size_t(b.size),
                                            _p(codes, C.c_uint8)))
    return codes


# ---- quantizer D ----
def quantize_d_rows(x, bits):
    x = _f32(x)
    rows, dim = x.shape
    b = _u8(np.atleast_1d(bits))
    codes = np.empty((rows, dim), np.uint8)
    scales = np.empty(rows, np.float32)
    zps = np.empty(rows, np.float32)
    _chk(lib().orc_quantize_d_rows(_p(x, C.c_float), C.c_size_t(rows), C.c_size_t(dim),
                                   _p(b, C.c_uint8), C.c_size_t(b.size), _p(codes, C.c_uint8),
                                   _p(scales, C.c_float), _p(zps, C.c_float)))
    return codes, scales, zps


def dequantize_d_rows(codes, scales, zps):
    codes = _u8(codes)
    rows, dim = codes.shape
    scales, zps = _f32(scales), _f32(zps)
    out = np.empty((rows, dim), np.float32)
    lib().orc_dequantize_d_rows(_p(codes, C.c_uint8), C.c_size_t(rows), C.c_size_t(dim),
                                _p(scales, C.c_float), _p(zps, C.c_float), _p(out, C.c_float))
    return out


# ---- pack / unpack ----
def packed_len(n, bits):
    return int(lib().orc_packed_len(C.c_size_t(n), C.c_uint8(bits)))


def pack(codes, bits):
    codes = _u8(codes).ravel()
    out = np.empty(packed_len(codes.size, bits), np.uint8)
    _chk(lib().orc_pack(_p(codes, C.c_uint8), C.c_size_t(codes.size), C.c_uint8(bits),
                        _p(out, C.c_uint8)))
    return out


def unpack(packed, n, bits):
    packed = _u8(packed).ravel()
    out = np.empty(n, np.uint8)
    _chk(lib().orc_unpack(_p(packed, C.c_uint8), C.c_size_t(n), C.c_uint8(bits),
                          _p(out, C.c_uint8)))
    return out


# ---- linear ----
def linear_f32(x, w, bias=None, threads=1):
    x, w = _f32(x), _f32(w)
    M, K = x.shape
    K2, N = w.shape
    assert K == K2
    y = np.empty((M, N), np.float32)
    bp = _p(_f32(bias), C.c_float) if bias is not None else None
    if threads > 1:
        lib().orc_linear_f32_mt(_p(x, C.c_float), _p(w, C.c_float), bp, C.c_size_t(M),
                                C.c_size_t(K), C.c_size_t(N), _p(y, C.c_float), C.c_int(threads))
    else:
        lib().orc_linear_f32(_p(x, C.c_float), _p(w, C.c_float), bp, C.c_size_t(M),
                             C.c_size_t(K), C.c_size_t(N), _p(y, C.c_float))
    return y


def linear_f64(x, w, bias=None):
    x, w = _f32(x), _f32(w)
    M, K = x.shape
    _, N = w.shape
    y = np.empty((M, N), np.float64)
    bp = _p(_f32(bias), C.c_float) if bias is not None else None
    lib().orc_linear_f64(_p(x, C.c_float), _p(w, C.c_float), bp, C.c_size_t(M), C.c_size_t(K),
                         C.c_size_t(N), _p(y, C.c_double))
    return y


def linear_i8_exact(qx, zx, qw, zw):
    qx, qw = _u8(qx), _u8(qw)
    M, K = qx.shape
    _, N = qw.shape
    acc = np.empty((M, N), np.int64)
    lib().orc_linear_i8_exact(_p(qx, C.c_uint8), C.c_int32(zx), _p(qw, C.c_uint8), C.c_int32(zw),
                              C.c_size_t(M), C.c_size_t(K), C.c_size_t(N), _p(acc, C.c_int64))
    return acc


# ---- schedules / p_sample ----
def beta_schedule(kind, T, beta_start=1e-4, beta_end=0.02):
    betas = np.empty(T, np.float32)
    _chk(lib().orc_beta_schedule(C.c_int32(kind), C.c_size_t(T), C.c_float(beta_start),
                                 C.c_float(beta_end), _p(betas, C.c_float)))
    return betas


def p_sample_coeffs(betas, t):
    betas = _f32(betas)
    c1, c2, sd = C.c_float(), C.c_float(), C.c_float()
    lib().orc_p_sample_coeffs(_p(betas, C.c_float), C.c_size_t(betas.size), C.c_size_t(t),
                              C.byref(c1), C.byref(c2), C.byref(sd))
    return np.float32(c1.value), np.float32(c2.value), np.float32(sd.value)


def p_sample(x_t, noise_pred, z, t, betas, guard_t0=True):
    x_t, noise_pred, betas = _f32(x_t), _f32(noise_pred), _f32(betas)
    batch, feat = x_t.shape
    tt = np.ascontiguousarray(t, dtype=np.uint64)
    out = np.empty_like(x_t)
    zp = _p(_f32(z), C.c_float) if z is not None else None
    lib().orc_p_sample(_p(x_t, C.c_float), _p(noise_pred, C.c_float), zp,
                       tt.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(batch),
                       C.c_size_t(feat), _p(betas, C.c_float), C.c_size_t(betas.size),
                       C.c_int32(int(guard_t0)), _p(out, C.c_float))
    return out


def add_noise(x_start, t, noise, betas):
    """diffuse-llm-rs/src/lib.rs:1100-1137 with the noise supplied (the reference's own draw is unseeded)."""
    x_start, noise, betas = _f32(x_start), _f32(noise), _f32(betas)
    batch, feat = x_start.shape
    tt = np.ascontiguousarray(t, dtype=np.uint64)
    out = np.empty_like(x_start)
    lib().orc_add_noise(_p(x_start, C.c_float), _p(noise, C.c_float), tt.ctypes.data_as(C.POINTER(C.c_size_t)),
                        C.c_size_t(batch), C.c_size_t(feat), _p(betas, C.c_float), C.c_size_t(betas.size),
                        _p(out, C.c_float))
    return out


def noise_normal(seed, stream, n, i0=0):
    """"dllm_noise v1": n standard normals of stream `stream` under `seed`, starting at element i0."""
    out = np.empty(int(n), np.float32)
    lib().orc_noise_normal(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(i0), C.c_size_t(int(n)), _p(out, C.c_float))
    return out


def progressive_bits(num_steps, t, decode_bits=4, min_bits=2):
    pre = C.c_int32()
    b = lib().orc_progressive_bits(C.c_size_t(num_steps), C.c_size_t(t), C.c_uint8(decode_bits),
                                   C.c_uint8(min_bits), C.byref(pre))
    return int(b), bool(pre.value)


# ---- composed paths (reference composition: dequantize_tensor ∘ x.dot(W)+b) ----
def qlinear_forward(x, codes, scales, zps, bias=None, group=128, threads=1):
    """diffuse-llm-rs/src/quantization.rs:81-85 composed with lib.rs:812."""
    w = dequantize_weight_grouped(codes, scales, zps, group)
    return linear_f32(x, w, bias, threads=threads)


def model_forward(x_tokens, layers, threads=1):
    """layers: list of (codes[K,N], scales, zps, bias|None, group). Chain of x·W+b."""
    h = _f32(x_tokens)
    for (codes, scales, zps, bias, group) in layers:
        h = qlinear_forward(h, codes, scales, zps, bias, group, threads=threads)
    return h


def sample(x0, layers, hidden, num_steps, betas, noises, guard_t0=True, threads=1):
    """DiffuseLLM::sample without cache, lib.rs:875-927, noise injected.
    x0: [batch, hidden*seq]; noises: list indexed by t of [batch, feat] (or None)."""
    x = _f32(x0).copy()
    batch, feat = x.shape
    for t in range(num_steps - 1, -1, -1):
        pred = model_forward(x.reshape(-1, hidden), layers, threads=threads).reshape(batch, feat)
        z = noises[t] if (noises is not None and t > 0) else None
        x = p_sample(x, pred, z, np.full(batch, t), betas, guard_t0)
    return x


def linear_i8(xq, codes, zp):
    """Exact integer form of `x.dot(dequantize_tensor(codes))` (diffuse-llm-rs/src/lib.rs:812 composed with
    quantization.rs:81-85, `(q - zp) * scale`) for a per-tensor quantized weight and int8 activations:
    y[m,n] = sum_k xq[m,k] * (codes[k,n] - zp), in int64 (numpy), the scale factored out.  Checker for
    dllm_qlinear_forward_i8 (tolerance 0)."""
    zi = int(zp)
    assert zi == zp, "per-tensor zero-points of quantize_tensor are integers (quantization.rs:55-56)"
    return np.asarray(xq, dtype=np.int64) @ (np.asarray(codes, dtype=np.int64) - zi)



# ---- int8 denoise mode (DLLM_PATH_I8): an arithmetic mode of this repo, not of the reference -----------------------------
# The reference's linear is f32 (lib.rs:812 on dequantize_tensor's output, quantization.rs:81-85).  The int8 mode keeps the
# reference's per-tensor codes, quantizes each token row of the activations to symmetric int8 and contracts exactly in
# integers; what is restated here is that arithmetic (so that the device result can be held to 1 ulp), and the tests bound its
# distance to the reference's f64 stack separately.
def bf16_round(x):
    """f32 -> bf16 (round to nearest even) -> f32, the activation format between the layers of the tcgen05 stacks."""
    u = _f32(x).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    nan = np.isnan(_f32(x))
    out = r.astype(np.uint32).view(np.float32).copy()
    out[nan] = np.nan
    return out


def rowquant_i8(x_bf16, wscale):
    """Per-token symmetric int8 activation quantizer: q = clamp(rint(x * (127 / max|x_row|)), -127, 127) in f32 arithmetic,
    rowscale = wscale * (max|x_row| / 127) (wscale for an all-zero row), rowsum = sum of the row's codes."""
    x = _f32(x_bf16)
    m = np.minimum(np.abs(x).max(axis=1), np.float32(3.0e38)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(m > 0, np.float32(127.0) / m, np.float32(0)).astype(np.float32)
    q = np.clip(np.rint(x * inv[:, None]), -127, 127).astype(np.int8)
    rowscale = (np.float32(wscale) * np.where(m > 0, m / np.float32(127.0), np.float32(1)).astype(np.float32)).astype(np.float32)
    return q, rowscale, q.astype(np.int64).sum(axis=1)


def linear_i8_deq(x, codes, scale, zp, bias=None):
    """One linear of the int8 mode on f32 activations x[M,K] and per-tensor codes[K,N] (quantize_tensor's, quantization.rs:38-79):
    y = (sum_k q_x q_w - zp sum_k q_x) * (scale * step_m) + b — `(q - zp) * scale` of quantization.rs:83 with the token's step
    factored out.  The integer part is exact; the float epilogue is one int->f32 conversion and one fused multiply-add,
    evaluated here in f64 and rounded once (the device result may differ by 1 ulp of f32)."""
    q, rowscale, rowsum = rowquant_i8(bf16_round(x), scale)
    # (every partial sum is an integer below 2^53: the f64 matmul is exact, and BLAS-fast)
    e = (q.astype(np.float64) @ np.asarray(codes, dtype=np.float64)).astype(np.int64) - int(zp) * rowsum[:, None]
    b = 0.0 if bias is None else _f32(bias).astype(np.float64)[None, :]
    return (e.astype(np.float32).astype(np.float64) * rowscale.astype(np.float64)[:, None] + b).astype(np.float32)


def model_forward_i8(x_tokens, layers):
    """Stack of int8-mode linears; layers: list of (codes[K,N], scale, zp, bias|None); bf16 activations between the layers
    (the last layer's output stays f32), as dllm_model_forward(path = DLLM_PATH_I8) runs it."""
    h = _f32(x_tokens)
    for i, (codes, scale, zp, bias) in enumerate(layers):
        h = linear_i8_deq(h, codes, scale, zp, bias)
        if i + 1 < len(layers):
            h = bf16_round(h)
    return h


# ---- AdaptiveQuantizer's statistics: the CKMS sketch (diffuse-llm-rs/src/quantization.rs:179-216) ------------------------------
# The reference keeps `quantiles::ckms::CKMS<f32>` (crate `quantiles = "0.7"`, diffuse-llm-rs/Cargo.toml:33 — un-vendored, no
# Cargo.lock) with error 0.01 and asks it for q = 0.0 and q = 1.0 (:208-209).  What follows restates the PUBLISHED algorithm the
# crate implements — Cormode, Korn, Muthukrishnan, Srivastava, "Effective Computation of Biased Quantiles over Data Streams"
# (ICDE 2005), low-biased invariant f(r, n) = 2 eps r: samples (v, g, delta) kept sorted by v; insert with g = 1 and delta =
# floor(f(r)) - 1 (0 at either end); compress every 1 / (2 eps) inserts, merging sample i into i + 1 when g_i + g_{i+1} +
# delta_{i+1} <= max(1, floor(f(r_i))); query(q) = the last sample whose successor's maximum rank exceeds q n + f(q n) / 2.
# Pure Python (small streams only): it is here to show what q = 0 and q = 1 return — the exact extremes seen so far, because a
# merge always keeps the LARGER value and its rank budget max(1, floor(2 eps r)) is 1 below rank 1 / eps, so neither the last
# sample nor the first is ever merged away — which is what the product's running min / max computes (parity of the crate's
# own code stays unpinned: it is not in /root/reference).
class CKMS:
    def __init__(self, error=0.01):
        self.eps = min(max(float(error), 1e-10), 0.99)
        self.every = max(1, int(1.0 / (2.0 * self.eps)))
        self.n = 0
        self.inserts = 0
        self.samples = []                     # [v, g, delta], sorted by v

    def _f(self, r):
        return max(1, int(np.floor(2.0 * self.eps * r)))

    def insert(self, v):
        v = float(np.float32(v))
        smp = self.samples
        i = 0
        r = 0
        while i < len(smp) and not (v < smp[i][0]):
            r += smp[i][1]
            i += 1
        delta = 0 if (i == 0 or i == len(smp)) else max(0, self._f(r) - 1)
        smp.insert(i, [v, 1, delta])
        self.n += 1
        self.inserts = (self.inserts + 1) % self.every
        if self.inserts == 0:
            self.compress()

    def compress(self):
        smp = self.samples
        if len(smp) < 3:
            return
        ranks = np.cumsum([e[1] for e in smp])          # rank of sample i = sum of g up to and including i
        i = len(smp) - 2
        while i >= 1:
            r_prev = int(ranks[i - 1])
            if smp[i][1] + smp[i + 1][1] + smp[i + 1][2] <= self._f(r_prev):
                smp[i + 1][1] += smp[i][1]
                del smp[i]
            i -= 1

    def query(self, q):
        smp = self.samples
        if not smp:
            return None
        nphi = q * self.n
        rhs = nphi + self._f(nphi) / 2.0
        r = 0
        for i in range(1, len(smp)):
            r += smp[i - 1][1]
            if r + smp[i][1] + smp[i][2] > rhs:
                return r, np.float32(smp[i - 1][0])
        return len(smp), np.float32(smp[-1][0])


def adaptive_compute_params(chunks, bits, error=0.01):
    """AdaptiveQuantizer::update_stats over `chunks` then compute_params (quantization.rs:198-216) with the sketch above:
    min = query(0.0) or 0.0, max = query(1.0) or 1.0, scale = (max - min) / q_max, zp = clamp(round(-min / scale), 0, q_max),
    all in f32 with f32::round (half away from zero)."""
    F32 = np.float32
    sk = CKMS(error)
    for c in chunks:
        for v in np.asarray(c, dtype=np.float32).ravel():
            sk.insert(v)
    lo, hi = sk.query(0.0), sk.query(1.0)
    mn = F32(0.0) if lo is None else lo[1]
    mx = F32(1.0) if hi is None else hi[1]
    q_max = F32(F32(1 << bits) - F32(1))
    with np.errstate(all="ignore"):
        scale = F32(F32(mx - mn) / q_max)
        r = F32(F32(-mn) / scale)
        t = np.trunc(r)
        r = F32(t + np.copysign(F32(1), r)) if abs(r - t) >= 0.5 else F32(t)
        zp = F32(min(max(r, F32(0)), q_max))
    return scale, zp, sk
