"""Writes tests/golden/reference_kat.json: every known-answer vector the reference's own
unit tests hold for the hot path, with the outputs the reference SOURCE
produces for them (derived in float32 by tests/np_restatement.py — the reference itself
cannot be compiled or run here, see oracle/dllm_oracle.h).  `ref_assert` records what the
reference test actually asserts, and whether that assertion holds for the reference's own
code.  Run from the repo root:  python tests/golden/make_reference_kat.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import np_restatement as R  # noqa: E402

F = np.float32


def fl(a):
    return [float(F(v)) for v in np.asarray(a).ravel()]


kat = []

# diffusion_prefill/src/prefill_kv.rs:147-160
v = [0.1, 0.5, 1.0, 0.0]
c, s, z = R.quantize_d_row(v, 4)
kat.append(dict(name="D_prefill_kv_test_quantization", quantizer="D", bits=4, input=v,
                codes=c.tolist(), scale=float(s), zero_point=float(z),
                dequant=fl(R.dequantize_cd(c, s, z)),
                ref="diffusion_prefill/src/prefill_kv.rs:147-160",
                ref_assert="|orig-dec| < 0.1", ref_assert_holds=True))

# diffuse-llm-rs/src/quantization.rs:254-265
v = [1.0, 2.0, 3.0, 4.0]
c, s, z = R.quantize_tensor(v, 4)
kat.append(dict(name="B_test_quantized_tensor", quantizer="B", bits=4, input=v,
                codes=c.tolist(), scale=float(s), zero_point=float(z),
                dequant=fl(R.dequantize_tensor(c, s, z)), compression_ratio=8.0,
                ref="diffuse-llm-rs/src/quantization.rs:254-265",
                ref_assert="len==4 and compression_ratio>4", ref_assert_holds=True))

# diffuse-llm-rs/src/quantization.rs:242-252
v = [1.0, 2.0, 3.0, 4.0, 5.0]
c, s, z = R.quantize_tensor(v, 4)
d = R.dequantize_tensor(c, s, z)
kat.append(dict(name="B_test_quantization", quantizer="B", bits=4, input=v,
                codes=c.tolist(), scale=float(s), zero_point=float(z), dequant=fl(d),
                ref="diffuse-llm-rs/src/quantization.rs:242-252",
                ref_assert="|orig-deq| < 0.1",
                ref_assert_holds=bool(np.all(np.abs(np.asarray(v, F) - d) < 0.1))))

# quantization/src/lib.rs:61-79 (and quantize.rs:222-233: same data, shape-only)
v = [-1.0, 0.0, 1.0, 2.0, 3.0, 4.0]
c = R.quantize_a(v, 0, 1.0, 0)
d = R.dequantize_tensor(c, 1.0, 0.0)
kat.append(dict(name="A_roundtrip_int8", quantizer="A", qtype=0, scale=1.0, zero_point=0,
                input=v, codes=c.tolist(), dequant=fl(d),
                ref="quantization/src/lib.rs:61-79",
                ref_assert="|a-b| < 0.1",
                ref_assert_holds=bool(np.all(np.abs(np.asarray(v, F) - d) < 0.1))))

# quantization/examples/basic.rs:25-33 (example input; prints only)
v = [-1.5, -0.5, 0.5, 1.5, 2.0, 3.0, 4.0, 5.0]
c = R.quantize_a(v, 0, 1.0, 0)
kat.append(dict(name="A_example_basic", quantizer="A", qtype=0, scale=1.0, zero_point=0,
                input=v, codes=c.tolist(), dequant=fl(R.dequantize_tensor(c, 1.0, 0.0)),
                ref="quantization/examples/basic.rs:25-33", ref_assert="none (prints)",
                ref_assert_holds=True))

# quantization/src/calibrate.rs:123-132: data 1..6, 8 bits asymmetric
mn, mx = F(1.0), F(6.0)
scale = F((mx - mn) / F(255.0))
zp = int(R.as_i32(R.round_half_away(F(-mn) / scale)))
kat.append(dict(name="A_calibration_1_to_6", quantizer="calib", bits=8, symmetric=False,
                min=1.0, max=6.0, total_samples=6, scale=float(scale), zero_point=zp,
                ref="quantization/src/calibrate.rs:123-132",
                ref_assert="|scale-0.0235|<1e-3 and zp==-43",
                ref_assert_holds=bool(abs(float(scale) - 0.0235) < 1e-3 and zp == -43)))

# diffusion_prefill/src/fusion_ann.rs:144-165: two rows, bits cycle [4, 8]
rows = [[0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2, 0.1]]
out = []
for i, r in enumerate(rows):
    b = [4, 8][i % 2]
    c, s, z = R.quantize_d_row(r, b)
    out.append(dict(bits=b, codes=c.tolist(), scale=float(s), zero_point=float(z)))
kat.append(dict(name="D_fusion_ann_rows", quantizer="D_rows", bits=[4, 8], input=rows, rows=out,
                ref="diffusion_prefill/src/fusion_ann.rs:144-165",
                ref_assert="len==2, bits==4 and 8", ref_assert_holds=True))

path = os.path.join(os.path.dirname(__file__), "reference_kat.json")
with open(path, "w") as f:
    json.dump(kat, f, indent=1)
print("wrote", path, len(kat), "vectors")
