"""dllm_b200 — Python front-end of libdllm_b200.so, mirroring the reference's Rust interfaces
for the quantized-linear / quantize / KV-quant hot path (names, argument meaning, error
behaviour).  Everything computes on the GPU through the C ABI; there is no CPU fallback.

  dllm_b200.quantization  <- diffuse_llm_rs::quantization      (diffuse-llm-rs/src/quantization.rs)
  dllm_b200.quant         <- the `quantization` crate          (quantization/src/*.rs)
  dllm_b200.kvquant       <- prefill_kvquant_rs::kvquant and diffusion_prefill::prefill_kv / fusion_ann
  dllm_b200.diffuse_llm   <- diffuse_llm_rs::diffuse_llm       (diffuse-llm-rs/src/lib.rs)
"""
from . import _lib
from ._lib import (DllmError, QuantizationError, InvalidParams, UnsupportedOperation, ShapeMismatch,
                   CalibrationRequired, ReferencePanic, NoDevice, PATH_AUTO, PATH_SIMT, PATH_UMMA, PATH_GEMV, PATH_I8,
                   KV_TENSOR_B, KV_ROW_D, KV_FIXED_C)
from .runtime import Context, QWeight, dequant_matmul, default_context

__all__ = ["Context", "QWeight", "dequant_matmul", "default_context", "DllmError", "QuantizationError",
           "InvalidParams", "UnsupportedOperation", "ShapeMismatch", "CalibrationRequired", "ReferencePanic",
           "NoDevice", "PATH_AUTO", "PATH_SIMT", "PATH_UMMA", "PATH_GEMV", "PATH_I8", "KV_TENSOR_B", "KV_ROW_D", "KV_FIXED_C"]
