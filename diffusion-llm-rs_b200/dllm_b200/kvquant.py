"""Mirrors of the KV-quantization seams:
  prefill_kvquant_rs::kvquant::{Quantizer, BitQuantizer, CompressedVector, PrefillKVQuant}
      (prefill-kvquant-rs/lib.rs)                                   — quantizer C
  diffusion_prefill::prefill_kv::{BitQuantizer, KVCache, CompressedVector}, FusionANN::quantize
      (diffusion_prefill/src/prefill_kv.rs, fusion_ann.rs)          — quantizer D
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _lib as L
from .runtime import Context, default_context


class Quantizer:
    """trait Quantizer: Send + Sync (prefill-kvquant-rs/lib.rs:29-32, prefill_kv.rs:42-45)"""

    def quantize(self, input, bits: int):
        raise NotImplementedError

    def dequantize(self, input, bits: int):
        raise NotImplementedError


class BitQuantizer(Quantizer):
    """{scale, zero_point} fixed quantizer, truncating (lib.rs:34-53, prefill_kv.rs:48-67)"""

    def __init__(self, scale: float, zero_point: float, ctx: Context | None = None):
        self.scale, self.zero_point = np.float32(scale), np.float32(zero_point)
        self._ctx = ctx or default_context()

    def quantize(self, input, bits: int):
        return self._ctx.quantize_c(input, bits, self.scale, self.zero_point)

    def dequantize(self, input, bits: int):
        return self._ctx.dequantize_cd(input, self.scale, self.zero_point)


@dataclass
class CompressedVector:
    id: str
    data: np.ndarray
    bits: int
    original_shape: List[int]
    quant_scale: float = 0.0        # prefill_kv.rs:31-32 (absent in prefill-kvquant-rs/lib.rs:61-67)
    quant_zero_point: float = 0.0


@dataclass
class SystemConfig:
    """prefill-kvquant-rs/lib.rs:77-91"""
    num_quantizers: int = 4
    cache_size: int = 1024
    quantization_bits: List[int] = field(default_factory=lambda: [4, 6, 8, 16])


@dataclass
class TokenizedVector:
    id: str
    tokens: List[int]
    embeddings: np.ndarray   # Array2<f32>


class PrefillKVQuant:
    """prefill-kvquant-rs/lib.rs:99-147"""

    def __init__(self, config: SystemConfig, ctx: Context | None = None):
        self._ctx = ctx or default_context()
        self.config = config
        self.quantizers = [BitQuantizer(L.lib().dllm_bitquantizer_scale(b), 0.0, self._ctx)
                           for b in config.quantization_bits]           # :102-110
        self.compression_ratio = 1.0

    def quantize_vectors(self, tokens: List[TokenizedVector], bits: List[int]) -> List[CompressedVector]:
        """:127-146.  `quantizers[bits/2]` indexing and the bits cycle are the reference's; an index
        past the Vec raises ReferencePanic."""
        if not tokens or not bits:
            return []
        shapes = {t.embeddings.shape for t in tokens}
        if len(shapes) == 1:   # one batched launch set
            emb = np.stack([np.asarray(t.embeddings, np.float32) for t in tokens])
            codes = self._ctx.kvquant_quantize_vectors(emb, self.config.quantization_bits, bits)
            per = [codes[i].ravel() for i in range(len(tokens))]
        else:
            per = []
            for i, t in enumerate(tokens):
                b = bits[i % len(bits)]
                per.append(self._ctx.kvquant_quantize_vectors(np.asarray(t.embeddings, np.float32)[None],
                                                              self.config.quantization_bits, [b])[0].ravel())
        return [CompressedVector(t.id, per[i], bits[i % len(bits)], list(t.embeddings.shape))
                for i, t in enumerate(tokens)]


class KVCache:
    """diffusion_prefill/src/prefill_kv.rs:35-139 — the quantizer-D store (host dict of compressed rows)"""

    def __init__(self, embedding_dim: int = 768, ctx: Context | None = None):
        self.embedding_dim = embedding_dim
        self.store = {}
        self._ctx = ctx or default_context()

    def compress_vector(self, id: str, vector, bits: int) -> CompressedVector:
        v = np.ascontiguousarray(vector, np.float32).reshape(1, -1)
        codes, scales, zps = self._ctx.quantize_d_rows(v, [bits])
        return CompressedVector(id, codes[0], bits, [v.shape[1]], scales[0], zps[0])

    def compress_batch(self, vectors, bits: List[int]) -> List[CompressedVector]:
        """FusionANN::quantize (fusion_ann.rs:53-63): row i uses bits[i % len]; one fused launch."""
        x = np.ascontiguousarray(vectors, np.float32)
        codes, scales, zps = self._ctx.quantize_d_rows(x, bits)
        return [CompressedVector(str(i), codes[i], bits[i % len(bits)], [x.shape[1]], scales[i], zps[i])
                for i in range(x.shape[0])]

    def decompress_vector(self, vector: CompressedVector):
        return self._ctx.dequantize_d_rows(vector.data.reshape(1, -1), [vector.quant_scale],
                                           [vector.quant_zero_point])[0]

    def insert_batch(self, vectors: List[CompressedVector]):
        for v in vectors:
            self.store[v.id] = v

    def get_batch(self, ids):
        out = []
        for i in ids:
            v = self.store.get(str(i))
            out.append(self.decompress_vector(v) if v is not None else np.zeros(self.embedding_dim, np.float32))
        return out

    def size_bytes(self) -> int:
        return sum(v.data.size + len(v.id) for v in self.store.values())


class FusionANN:
    """fusion_ann.rs:53-63 — only the quantize seam is on the path"""

    def __init__(self, ctx: Context | None = None):
        self._cache = KVCache(ctx=ctx)

    def quantize(self, vectors, bits: List[int]) -> List[CompressedVector]:
        return self._cache.compress_batch(vectors, bits)
