"""Independent numpy (float32) restatement of the reference quantizers.

Second opinion for the C oracle: written from the same reference source lines but sharing
no code with oracle/dllm_oracle.c.  tests/test_oracle_golden.py requires the two to agree
bit-for-bit.  All arithmetic stays in np.float32 (IEEE binary32, true division).
"""
import numpy as np

F = np.float32


def round_half_away(x):
    """Rust f32::round. np.round is half-to-even, so build it from trunc (exact in f32)."""
    x = np.asarray(x, dtype=F)
    r = np.trunc(x)
    d = np.abs(x - r)  # exact
    with np.errstate(invalid="ignore"):
        out = np.where(d >= F(0.5), r + np.copysign(F(1), x), r)
    return out.astype(F)


def clamp_rs(x, lo, hi):
    """Rust f32::clamp (NaN passes through)."""
    x = np.asarray(x, dtype=F).copy()
    with np.errstate(invalid="ignore"):
        x = np.where(x < F(lo), F(lo), x)
        x = np.where(x > F(hi), F(hi), x)
    return x.astype(F)


def as_u8(x):
    """Rust `f32 as u8`: toward zero, saturating, NaN -> 0."""
    x = np.asarray(x, dtype=F)
    with np.errstate(invalid="ignore"):
        y = np.where(np.isnan(x), F(0), x)
        y = np.clip(y, F(0), F(255))
    return np.trunc(y).astype(np.uint8)


def as_i32(x):
    x = np.asarray(x, dtype=F)
    with np.errstate(invalid="ignore"):
        y = np.where(np.isnan(x), F(0), x).astype(np.float64)
        y = np.clip(y, -2147483648.0, 2147483647.0)
    return np.trunc(y).astype(np.int64)


def fold_max(x):
    # f32::max ignores NaN; start from -inf
    x = np.asarray(x, dtype=F)
    x = x[~np.isnan(x)]
    return F(-np.inf) if x.size == 0 else F(x.max())


def fold_min(x):
    x = np.asarray(x, dtype=F)
    x = x[~np.isnan(x)]
    return F(np.inf) if x.size == 0 else F(x.min())


# quantizer B — diffuse-llm-rs/src/quantization.rs:38-68
def quantize_tensor(data, bits):
    assert 1 <= bits <= 8
    data = np.asarray(data, dtype=F).ravel()
    with np.errstate(all="ignore"):
        mx, mn = fold_max(data), fold_min(data)
        q_max = F(F(1 << bits) - F(1))
        scale = F((mx - mn) / q_max)
        if scale == 0:
            scale = F(1)
        zp = F(F(0) - F(mn / scale))
        zp = as_u8(round_half_away(clamp_rs(zp, 0, q_max)))[()]
        v = round_half_away((data / scale).astype(F) + F(zp))
        codes = np.clip(as_i32(v), 0, (1 << bits) - 1).astype(np.uint8)
    return codes, scale, F(zp)


def dequantize_tensor(codes, scale, zp):
    return ((np.asarray(codes, np.uint8).astype(F) - F(zp)).astype(F) * F(scale)).astype(F)


# quantizer A — quantization/src/quantize.rs:111-154
_A_RANGE = {0: (-128.0, 127.0), 1: (-8.0, 7.0), 2: (0.0, 1.0), 3: (-127.0, 127.0)}


def quantize_a(data, qtype, scale=1.0, zero_point=0):
    lo, hi = _A_RANGE[qtype]
    data = np.asarray(data, dtype=F).ravel()
    with np.errstate(all="ignore"):
        v = (data / F(scale)).astype(F) + F(zero_point)
        v = np.fmin(np.fmax(v, F(lo)), F(hi))  # fmax/fmin ignore NaN like f32::max/min
        return as_u8(round_half_away(v))


# quantizer C — prefill-kvquant-rs/lib.rs:39-46, :105
def scale_c(bits):
    return F(F(1) / F((1 << bits) - 1))


def quantize_c(data, bits, scale, zp=0.0):
    data = np.asarray(data, dtype=F).ravel()
    with np.errstate(all="ignore"):
        scaled = ((data - F(zp)).astype(F) / F(scale)).astype(F)
        return as_u8(clamp_rs(scaled, 0, F((1 << bits) - 1)))


def dequantize_cd(codes, scale, zp):
    return ((np.asarray(codes, np.uint8).astype(F) * F(scale)).astype(F) + F(zp)).astype(F)


# quantizer D — diffusion_prefill/src/prefill_kv.rs:104-121
def quantize_d_row(row, bits):
    row = np.asarray(row, dtype=F).ravel()
    with np.errstate(all="ignore"):
        mn, mx = fold_min(row), fold_max(row)
        levels = F((1 << bits) - 1)
        scale = F((mx - mn) / levels)
        scaled = ((row - mn).astype(F) / scale).astype(F)
        codes = as_u8(clamp_rs(scaled, 0, levels))
    return codes, scale, mn


# pack — layout defined by this build (LSB-first)
def pack(codes, bits):
    codes = np.asarray(codes, np.uint8).ravel()
    per = 8 // bits
    n = codes.size
    pad = (-n) % per
    c = np.concatenate([codes & ((1 << bits) - 1), np.zeros(pad, np.uint8)]).reshape(-1, per)
    out = np.zeros(c.shape[0], np.uint16)
    for j in range(per):
        out |= c[:, j].astype(np.uint16) << (j * bits)
    return out.astype(np.uint8)


def unpack(packed, n, bits):
    packed = np.asarray(packed, np.uint8).ravel()
    per = 8 // bits
    idx = np.arange(n)
    return ((packed[idx // per] >> ((idx % per) * bits)) & ((1 << bits) - 1)).astype(np.uint8)
