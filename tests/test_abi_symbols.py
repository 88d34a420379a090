"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol that
include/dllm_b200.h declares, the ctypes table covers the header, and the library refuses to
run (loudly) without a CUDA device.  No compute calls are made here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dllm_b200.h")
LIB = os.path.join(ROOT, "diffusion-llm-rs_b200", "lib", "libdllm_b200.so")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"DLLM_API\s+[^;(]*?\b(dllm_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "diffusion-llm-rs_b200"), "-j8", "-s"])
    return LIB


def test_header_declares_a_reasonable_surface():
    syms = header_symbols()
    assert len(syms) >= 60
    for must in ("dllm_quantize_tensor", "dllm_dequantize_tensor", "dllm_pack", "dllm_unpack",
                 "dllm_quantize_d_rows", "dllm_qlinear_forward", "dllm_dequant_matmul", "dllm_sample",
                 "dllm_kv_quantize", "dllm_tp_init"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib], text=True)
    exported = set(re.findall(r"\sT\s+(dllm_\w+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = sorted(exported - set(header_symbols()))
    assert not extra, f"exported but not declared in the header: {extra}"


def test_ctypes_table_matches_header(built_lib):
    from dllm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.lib()
    for name in header_symbols():
        assert hasattr(lib, name)


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    # each public block names the reference file it replaces
    for ref in ("diffuse-llm-rs/src/quantization.rs", "quantization/src/quantize.rs", "prefill-kvquant-rs/lib.rs",
                "diffusion_prefill/src/prefill_kv.rs", "diffuse-llm-rs/src/lib.rs", "calibrate.rs"):
        assert ref in src


def test_no_cpu_fallback_without_a_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from dllm_b200 import _lib, Context, NoDevice
    assert _lib.lib().dllm_device_count() == 0
    with pytest.raises(NoDevice):
        Context(0)
    h = ctypes.c_void_p()
    assert _lib.lib().dllm_ctx_create(0, ctypes.byref(h)) == _lib.ERR_NO_DEVICE and not h.value


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "diffusion-llm-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "dllm_oracle" not in txt and "libdllm_oracle" not in txt, f
