"""ctypes loader for libdllm_b200.so (include/dllm_b200.h).

The product path has no CPU fallback: if the shared library is missing, or no sm_100 CUDA
device is present, every entry point raises.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# DLLM_B200_LIB: another build of the same library (timing experiments: scripts/build_variants.sh)
LIB_PATH = os.environ.get("DLLM_B200_LIB") or os.path.join(_PKG, "lib", "libdllm_b200.so")

OK = 0
ERR_INVALID_PARAMS, ERR_UNSUPPORTED, ERR_SHAPE, ERR_CALIBRATION_REQUIRED = 1, 2, 3, 4
ERR_IO, ERR_SERIALIZATION, ERR_INVALID_DATA_FORMAT, ERR_INDEX = 5, 6, 7, 8
ERR_CUDA, ERR_NO_DEVICE, ERR_NCCL, ERR_OOM, ERR_NULL = 100, 101, 102, 103, 104

QT_INT8, QT_INT4, QT_BINARY, QT_FLOAT8 = 0, 1, 2, 3
BETA_LINEAR, BETA_QUADRATIC, BETA_COSINE = 0, 1, 2
KV_TENSOR_B, KV_ROW_D, KV_FIXED_C = 0, 1, 2
PATH_AUTO, PATH_SIMT, PATH_UMMA, PATH_GEMV, PATH_I8 = 0, 1, 2, 3, 4


class DllmError(RuntimeError):
    """Base error; `code` is the DLLM_ERR_* status."""

    def __init__(self, code, msg=""):
        super().__init__(f"[dllm status {code}] {msg}")
        self.code = code


class QuantizationError(DllmError):
    """quantization/src/error.rs:19-40"""


class InvalidParams(QuantizationError):
    pass


class UnsupportedOperation(QuantizationError):
    pass


class ShapeMismatch(QuantizationError):
    pass


class CalibrationRequired(QuantizationError):
    pass


class ReferencePanic(DllmError):
    """The reference would panic here (assert!, index out of bounds)."""


class NoDevice(DllmError):
    pass


_ERR = {
    ERR_INVALID_PARAMS: InvalidParams, ERR_UNSUPPORTED: UnsupportedOperation, ERR_SHAPE: ShapeMismatch,
    ERR_CALIBRATION_REQUIRED: CalibrationRequired, ERR_INDEX: ReferencePanic, ERR_NO_DEVICE: NoDevice,
}

_lib = None

c_f32p = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_uint8)
c_vp = C.c_void_p
c_sz = C.c_size_t

# name -> (restype, argtypes).  Kept in one table so tests can check it against the header.
SIGNATURES = {
    "dllm_version": (C.c_char_p, []),
    "dllm_device_count": (C.c_int32, []),
    "dllm_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(c_vp)]),
    "dllm_ctx_create_on_stream": (C.c_int32, [C.c_int32, c_vp, C.POINTER(c_vp)]),
    "dllm_ctx_destroy": (None, [c_vp]),
    "dllm_ctx_sync": (C.c_int32, [c_vp]),
    "dllm_ctx_stream": (c_vp, [c_vp]),
    "dllm_last_error": (C.c_char_p, [c_vp]),
    "dllm_launch_count": (C.c_uint64, [c_vp]),
    "dllm_sm_count": (C.c_int32, [c_vp]),
    "dllm_graph_replay_count": (C.c_uint64, [c_vp]),
    "dllm_copy_bytes": (C.c_int32, [c_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "dllm_selftest_division": (C.c_int32, [c_vp, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "dllm_profile_begin": (C.c_int32, [c_vp]),
    "dllm_profile_end": (C.c_int32, [c_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]),
    "dllm_malloc": (C.c_int32, [c_vp, c_sz, C.POINTER(c_vp)]),
    "dllm_free": (C.c_int32, [c_vp, c_vp]),
    "dllm_memcpy_h2d": (C.c_int32, [c_vp, c_vp, c_vp, c_sz]),
    "dllm_memcpy_d2h": (C.c_int32, [c_vp, c_vp, c_vp, c_sz]),
    "dllm_host_alloc": (C.c_int32, [c_sz, C.POINTER(c_vp)]),
    "dllm_host_free": (C.c_int32, [c_vp]),
    "dllm_quantize_tensor": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, c_vp, c_f32p, c_f32p]),
    "dllm_dequantize_tensor": (C.c_int32, [c_vp, c_vp, c_sz, C.c_float, C.c_float, c_vp]),
    "dllm_quantize_tensor_dev": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, C.c_int32, c_vp, c_vp]),
    "dllm_dequantize_tensor_dev": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, C.c_int32, c_vp, c_vp]),
    "dllm_quantize_codes": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, C.c_float, C.c_float, c_vp]),
    "dllm_compression_ratio": (C.c_float, [c_sz, c_sz, C.c_uint8]),
    "dllm_quantize_a": (C.c_int32, [c_vp, c_vp, c_sz, C.c_int32, C.c_float, C.c_int32, c_vp]),
    "dllm_dequantize_a": (C.c_int32, [c_vp, c_vp, c_sz, C.c_float, C.c_int32, c_vp]),
    "dllm_minmax": (C.c_int32, [c_vp, c_vp, c_sz, c_f32p, c_f32p]),
    "dllm_calibrate_params": (C.c_int32, [C.c_float, C.c_float, c_sz, C.c_uint8, C.c_int32, c_f32p,
                                          C.POINTER(C.c_int32)]),
    "dllm_bitquantizer_scale": (C.c_float, [C.c_uint8]),
    "dllm_quantize_c": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, C.c_float, C.c_float, c_vp]),
    "dllm_dequantize_cd": (C.c_int32, [c_vp, c_vp, c_sz, C.c_float, C.c_float, c_vp]),
    "dllm_kvquant_quantize_vectors": (C.c_int32, [c_vp, c_vp, c_sz, c_sz, c_vp, c_sz, c_vp, c_sz, c_vp]),
    "dllm_quantize_d_rows": (C.c_int32, [c_vp, c_vp, c_sz, c_sz, c_vp, c_sz, c_vp, c_vp, c_vp]),
    "dllm_dequantize_d_rows": (C.c_int32, [c_vp, c_vp, c_sz, c_sz, c_vp, c_vp, c_vp]),
    "dllm_quantize_d_rows_dev": (C.c_int32, [c_vp, c_vp, c_sz, c_sz, C.c_uint8, C.c_int32, c_vp, c_vp, c_vp]),
    "dllm_dequantize_d_rows_dev": (C.c_int32, [c_vp, c_vp, c_sz, c_sz, C.c_uint8, C.c_int32, c_vp, c_vp, c_vp]),
    "dllm_packed_len": (c_sz, [c_sz, C.c_uint8]),
    "dllm_pack": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, c_vp]),
    "dllm_unpack": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, c_vp]),
    "dllm_pack_dev": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, c_vp]),
    "dllm_unpack_dev": (C.c_int32, [c_vp, c_vp, c_sz, C.c_uint8, c_vp]),
    "dllm_qweight_quantize": (C.c_int32, [c_vp, c_vp, c_sz, c_sz, C.c_uint8, c_sz, c_vp, C.POINTER(c_vp)]),
    "dllm_qweight_quantize_dev": (C.c_int32, [c_vp, c_vp, c_sz, c_sz, C.c_uint8, c_sz, c_vp, C.POINTER(c_vp)]),
    "dllm_qweight_from_codes": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, C.c_uint8, c_sz, c_vp,
                                            C.POINTER(c_vp)]),
    "dllm_qweight_export": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dllm_qweight_info": (C.c_int32, [c_vp, C.POINTER(c_sz), C.POINTER(c_sz), c_u8p, C.POINTER(c_sz),
                                      C.POINTER(c_sz)]),
    "dllm_qweight_destroy": (None, [c_vp]),
    "dllm_qweight_serialized_size": (c_sz, [c_vp]),
    "dllm_qweight_serialize": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, C.POINTER(c_sz)]),
    "dllm_qweight_deserialize": (C.c_int32, [c_vp, c_vp, c_sz, C.POINTER(c_vp)]),
    "dllm_qweight_save": (C.c_int32, [c_vp, c_vp, C.c_char_p]),
    "dllm_qweight_load": (C.c_int32, [c_vp, C.c_char_p, C.POINTER(c_vp)]),
    "dllm_qlinear_forward": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, c_vp, C.c_int32]),
    "dllm_qlinear_forward_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, c_vp, C.c_int32]),
    "dllm_qlinear_forward_i8": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, c_vp]),
    "dllm_qlinear_forward_i8_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, c_vp]),
    "dllm_dequant_matmul": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, C.c_uint8, c_sz, c_vp, c_vp, c_sz,
                                        c_vp, C.c_int32]),
    "dllm_model_create": (C.c_int32, [c_vp, c_sz, C.POINTER(c_vp), c_sz, c_sz, C.c_int32, C.c_float, C.c_float,
                                      C.POINTER(c_vp)]),
    "dllm_model_destroy": (None, [c_vp]),
    "dllm_model_forward": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, c_vp, C.c_int32]),
    "dllm_model_forward_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, c_sz, c_vp, C.c_int32]),
    "dllm_beta_schedule": (C.c_int32, [C.c_int32, c_sz, C.c_float, C.c_float, c_vp]),
    "dllm_p_sample": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, C.c_int32, c_vp]),
    "dllm_add_noise": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, c_vp]),
    "dllm_add_noise_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, c_sz, c_vp]),
    "dllm_denoise_step_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, c_sz, C.c_int32, C.c_int32]),
    "dllm_denoise_step": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, c_sz, C.c_int32, C.c_int32]),
    "dllm_last_step_breakdown": (C.c_int32, [c_vp, c_f32p, c_f32p, c_f32p]),
    "dllm_sample": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, c_sz, C.c_int32, C.c_int32, c_vp]),
    "dllm_noise_fill": (C.c_int32, [c_vp, C.c_uint64, C.c_uint64, C.c_uint64, c_sz, c_vp]),
    "dllm_noise_fill_dev": (C.c_int32, [c_vp, C.c_uint64, C.c_uint64, C.c_uint64, c_sz, c_vp]),
    "dllm_denoise_step_seeded_dev": (C.c_int32, [c_vp, c_vp, c_vp, C.c_uint64, c_sz, c_sz, c_sz, C.c_int32, C.c_int32]),
    "dllm_p_sample_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_sz, c_sz, C.c_int32, c_vp]),
    "dllm_p_sample_seeded_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, C.c_uint64, c_sz, c_sz, c_sz, C.c_int32, c_vp]),
    "dllm_sample_seeded": (C.c_int32, [c_vp, c_vp, c_vp, C.c_uint64, c_sz, c_sz, c_sz, C.c_int32, C.c_int32, C.c_int32, c_vp]),
    "dllm_sample_seeded_dev": (C.c_int32, [c_vp, c_vp, c_vp, C.c_uint64, c_sz, c_sz, c_sz, C.c_int32, C.c_int32, C.c_int32]),
    "dllm_progressive_bits": (C.c_uint8, [c_sz, c_sz, C.c_uint8, C.c_uint8, C.POINTER(C.c_int32)]),
    "dllm_kv_quantize": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, c_sz, c_sz, C.c_uint8, C.c_int32, C.POINTER(c_vp)]),
    "dllm_kv_quantize_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, c_sz, c_sz, C.c_uint8, C.c_int32,
                                         C.POINTER(c_vp)]),
    "dllm_kv_quantize_sharded_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_sz, c_sz, c_sz, C.c_uint8, C.c_int32,
                                                 C.POINTER(c_vp)]),
    "dllm_kv_update_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp]),
    "dllm_kv_create": (C.c_int32, [c_vp, c_sz, c_sz, c_sz, C.c_uint8, C.c_int32, C.POINTER(c_vp)]),
    "dllm_kv_append": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz]),
    "dllm_kv_append_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz]),
    "dllm_kv_seq_len": (c_sz, [c_vp]),
    "dllm_kv_dequantize": (C.c_int32, [c_vp, c_vp, c_vp, c_vp]),
    "dllm_kv_dequantize_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp]),
    "dllm_kv_export": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dllm_kv_memory_usage": (c_sz, [c_vp]),
    "dllm_kv_destroy": (None, [c_vp]),
    "dllm_kvcache_create": (C.c_int32, [c_vp, c_sz, c_sz, c_sz, C.c_uint8, C.c_uint8, C.c_int32, C.POINTER(c_vp)]),
    "dllm_kvcache_update_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz]),
    "dllm_kvcache_append_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp, c_sz]),
    "dllm_kvcache_set_phase": (C.c_int32, [c_vp, c_vp, C.c_int32]),
    "dllm_kvcache_set_decode_bits": (C.c_int32, [c_vp, c_vp, C.c_uint8]),
    "dllm_kvcache_get_dev": (C.c_int32, [c_vp, c_vp, c_vp, c_vp]),
    "dllm_kvcache_info": (C.c_int32, [c_vp, C.POINTER(c_sz), C.POINTER(C.c_int32), c_u8p, C.POINTER(c_sz)]),
    "dllm_kvcache_copy": (c_vp, [c_vp, C.c_int32]),
    "dllm_kvcache_destroy": (None, [c_vp]),
    "dllm_tp_unique_id": (C.c_int32, [c_vp]),
    "dllm_tp_init": (C.c_int32, [c_vp, c_vp, C.c_int32, C.c_int32]),
    "dllm_tp_finalize": (C.c_int32, [c_vp]),
    "dllm_tp_configure": (C.c_int32, [c_vp, C.c_int32, C.c_int32, C.c_int32]),
    "dllm_tp_p2p_enable": (C.c_int32, [c_vp, c_sz]),
    "dllm_tp_p2p_status": (C.c_int32, [c_vp, C.POINTER(c_vp), C.POINTER(c_sz), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "dllm_tp_allreduce_dev": (C.c_int32, [c_vp, c_vp, c_sz]),
    "dllm_tp_allgather_cols_dev": (C.c_int32, [c_vp, c_vp, c_sz, c_sz, c_vp]),
    "dllm_model_set_parallel": (C.c_int32, [c_vp, c_vp, c_vp, c_sz]),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {_PKG}` (or __graft_entry__.build()). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError:
                if os.environ.get("DLLM_B200_LIB"):      # an older experiment build: the symbol is simply not callable
                    continue
                raise
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, ctx_handle=None):
    if rc == OK:
        return
    msg = ""
    if ctx_handle:
        raw = lib().dllm_last_error(ctx_handle)
        msg = raw.decode(errors="replace") if raw else ""
    raise _ERR.get(rc, DllmError)(rc, msg)
