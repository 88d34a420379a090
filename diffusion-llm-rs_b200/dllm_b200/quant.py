"""Mirror of the `quantization` crate (quantization/src/{quantize,types,calibrate,error}.rs)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import Enum
from typing import Optional

import numpy as np

from . import _lib as L
from ._lib import QuantizationError  # noqa: F401  (re-export, quantization/src/lib.rs:30)
from .runtime import Context, default_context, dequant_matmul  # noqa: F401


class QuantizationType(Enum):
    """quantize.rs:62-78"""
    Int8 = L.QT_INT8
    Int4 = L.QT_INT4
    Binary = L.QT_BINARY
    Float8 = L.QT_FLOAT8

    def bits(self) -> int:
        return {L.QT_INT8: 8, L.QT_INT4: 4, L.QT_BINARY: 1, L.QT_FLOAT8: 8}[self.value]


@dataclass
class QuantizationParams:
    """types.rs:21-40"""
    bits: int = 8
    scale: float = 1.0
    zero_point: int = 0
    symmetric: bool = True
    axis: Optional[int] = None


@dataclass
class QuantizationConfig:
    """types.rs:112-131 (unused by the reference's code; group_size feeds the grouped linears here)"""
    quant_method: str = "gptq"
    bits: int = 4
    group_size: int = 128
    sym: bool = True
    desc_act: bool = True
    true_sequential: bool = True


@dataclass
class QuantizedTensor:
    """types.rs:42-81"""
    data: np.ndarray
    shape: list
    params: QuantizationParams = field(default_factory=QuantizationParams)
    _ctx: Optional[Context] = None

    def len(self) -> int:
        return int(np.prod(self.shape)) if len(self.shape) else 1

    def is_empty(self) -> bool:
        return self.data.size == 0

    def dequantize(self):
        ctx = self._ctx or default_context()
        return ctx.dequantize_a(self.data, self.params.scale, self.params.zero_point).reshape(self.shape)


class Quantizer:
    """trait Quantizer, quantize.rs:81-90"""

    def quantize(self, data, qtype: QuantizationType) -> QuantizedTensor:
        raise NotImplementedError

    def dequantize(self, tensor: QuantizedTensor):
        raise NotImplementedError

    def get_params(self) -> QuantizationParams:
        raise NotImplementedError


class DefaultQuantizer(Quantizer):
    """quantize.rs:93-185.  `new` hard-codes scale 1.0 / zero_point 0 exactly like the reference
    (:98-108); `with_params` is the setter the reference lacks so calibration can feed it."""

    def __init__(self, bits: int, symmetric: bool, axis: Optional[int] = None, ctx: Context | None = None):
        self.params = QuantizationParams(bits=bits, scale=1.0, zero_point=0, symmetric=symmetric, axis=axis)
        self._ctx = ctx or default_context()

    @classmethod
    def with_params(cls, params: QuantizationParams, ctx: Context | None = None):
        q = cls(params.bits, params.symmetric, params.axis, ctx)
        q.params = params
        return q

    def quantize(self, data, qtype: QuantizationType) -> QuantizedTensor:
        data = np.asarray(data, np.float32)
        codes = self._ctx.quantize_a(data, qtype.value, self.params.scale, self.params.zero_point)
        return QuantizedTensor(codes, list(data.shape), QuantizationParams(**vars(self.params)), self._ctx)

    def dequantize(self, tensor: QuantizedTensor):
        return self._ctx.dequantize_a(tensor.data, tensor.params.scale, tensor.params.zero_point).reshape(tensor.shape)

    def get_params(self) -> QuantizationParams:
        return self.params


class quant_utils:
    """quantize.rs:188-215 (`pub mod utils`, re-exported as quant_utils, lib.rs:33)"""

    @staticmethod
    def quantize(data, qtype: QuantizationType, symmetric: bool, axis: Optional[int] = None, ctx=None):
        return DefaultQuantizer(qtype.bits(), symmetric, axis, ctx).quantize(data, qtype)

    @staticmethod
    def dequantize(tensor: QuantizedTensor, ctx=None):
        return DefaultQuantizer(tensor.params.bits, tensor.params.symmetric, tensor.params.axis,
                                ctx or tensor._ctx).dequantize(tensor)


class CalibrationData:
    """calibrate.rs:19-116.  The min/max fold runs on the GPU (dllm_minmax)."""

    def __init__(self, num_bins: int, per_channel: bool, ctx: Context | None = None):
        self.min = np.finfo(np.float32).max       # f32::MAX  :34
        self.max = np.finfo(np.float32).min       # f32::MIN  :35
        self.histogram = [0] * num_bins
        self.num_bins = num_bins
        self.total_samples = 0
        self.per_channel_stats = {} if per_channel else None
        self._ctx = ctx or default_context()

    def update(self, data, channel: Optional[int] = None):
        data = np.ascontiguousarray(data, np.float32)
        F = np.float32
        if data.size:
            mn, mx = self._ctx.minmax(data)
            mn, mx = min(F(np.finfo(F).max), mn), max(F(np.finfo(F).min), mx)   # fold seeds :43
        else:
            mn, mx = F(np.finfo(F).max), F(np.finfo(F).min)
        self.min, self.max = min(F(self.min), mn), max(F(self.max), mx)
        self.total_samples += data.size
        if channel is not None and self.per_channel_stats is not None:
            cmn, cmx = self.per_channel_stats.get(channel, (F(np.finfo(F).max), F(np.finfo(F).min)))
            self.per_channel_stats[channel] = (min(cmn, mn), max(cmx, mx))
        if self.max > self.min:                    # histogram :60-69 (bookkeeping, host)
            width = F(F(self.max - self.min) / F(self.num_bins))
            v = data.ravel()
            v = v[(v >= self.min) & (v <= self.max)]
            bins = np.minimum(np.floor((v - F(self.min)) / width).astype(np.int64), self.num_bins - 1)
            for b, c in zip(*np.unique(bins, return_counts=True)):
                self.histogram[int(b)] += int(c)

    def compute_params(self, bits: int, symmetric: bool) -> QuantizationParams:
        s, z = C.c_float(), C.c_int32()
        L.check(L.lib().dllm_calibrate_params(float(self.min), float(self.max), self.total_samples, bits,
                                              int(symmetric), C.byref(s), C.byref(z)))
        return QuantizationParams(bits=bits, scale=np.float32(s.value), zero_point=int(z.value),
                                  symmetric=symmetric, axis=None)

    def get_per_channel_stats(self):
        return self.per_channel_stats
