"""In-chain timeline of the GEMV kernel (trace build only): capture a CUDA graph of back-to-back calls, replay it, dump the
per-CTA globaltimer stamps of every launch.   DLLM_B200_LIB=<trace build> DLLM_GEMV_TRACE_CHAIN=1 python scripts/gemv_chain_trace.py K N bits M"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-llm-rs_b200"))
import torch
import dllm_b200
from dllm_b200 import QWeight, PATH_GEMV

K, N, bits, M = map(int, sys.argv[1:5])
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
w = torch.randn(K, N, device="cuda") * 0.02
npool = 6
pool = [QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128) for _ in range(npool)]
x = torch.randn(M, K, device="cuda")
y = torch.empty(M, N, device="cuda")
torch.cuda.synchronize()
with torch.cuda.stream(stream):
    for i in range(4):
        pool[i % npool].forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_GEMV)
stream.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=stream):
    for i in range(12):
        pool[i % npool].forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_GEMV)
with torch.cuda.stream(stream):
    for _ in range(3):
        g.replay()
stream.synchronize()
out = sys.argv[5] if len(sys.argv) > 5 else "gpurun_out/gemv_chain.csv"
os.makedirs(os.path.dirname(out), exist_ok=True)
print("dump", ctx._lib.dllm_debug_gemv_trace_dump(out.encode()))
