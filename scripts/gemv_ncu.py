"""Tiny GEMV driver for ncu captures: python scripts/gemv_ncu.py K N bits M [calls]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-llm-rs_b200"))
import torch
import dllm_b200
from dllm_b200 import QWeight, PATH_GEMV

K, N, bits, M = map(int, sys.argv[1:5])
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 6
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
w = torch.randn(K, N, device="cuda") * 0.02
pool = [QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128) for _ in range(3)]
x = torch.randn(M, K, device="cuda")
y = torch.empty(M, N, device="cuda")
torch.cuda.synchronize()
for i in range(calls):
    pool[i % 3].forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_GEMV)
ctx.sync()
print("ok", float(y.abs().sum()))
