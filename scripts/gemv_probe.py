"""Time one GEMV shape (per call and kernel-only): python scripts/gemv_probe.py K N bits M [path]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-llm-rs_b200"))
import torch
import dllm_b200
from dllm_b200 import QWeight

K, N, bits, M = map(int, sys.argv[1:5])
path = int(sys.argv[5]) if len(sys.argv) > 5 else dllm_b200.PATH_GEMV
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
w = torch.randn(K, N, device="cuda") * 0.02
npool = max(2, min(8, int(600e6 // (K * N * bits // 8)) + 1))
pool = [QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128) for _ in range(npool)]
x = torch.randn(M, K, device="cuda")
y = torch.empty(M, N, device="cuda")
torch.cuda.synchronize()


def run(i):
    pool[i % npool].forward_dev(x.data_ptr(), M, y.data_ptr(), path)


with torch.cuda.stream(stream):
    for i in range(6):
        run(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record(stream)
    for i in range(60):
        run(i)
    e1.record(stream)
    e1.synchronize()
us = e0.elapsed_time(e1) / 60 * 1e3
ctx._ck(ctx._lib.dllm_profile_begin(ctx.h))
with torch.cuda.stream(stream):
    for i in range(30):
        run(i)
nl, ms, fl, by = C.c_uint64(), C.c_double(), C.c_double(), C.c_double()
ctx._ck(ctx._lib.dllm_profile_end(ctx.h, C.byref(nl), C.byref(ms), C.byref(fl), C.byref(by)))
kus = ms.value / nl.value * 1e3
byts = K * N * bits / 8 + (K // 128) * N * 8 + 4 * M * (K + N)
# CUDA-graph replay of 4 x npool back-to-back calls: what a captured decode loop pays per linear (no host in the way)
g = torch.cuda.CUDAGraph()
ncap = 4 * npool
with torch.cuda.graph(g, stream=stream):
    for i in range(ncap):
        run(i)
with torch.cuda.stream(stream):
    g.replay()
    stream.synchronize()
    e0.record(stream)
    for _ in range(10):
        g.replay()
    e1.record(stream)
    e1.synchronize()
gus = e0.elapsed_time(e1) / (10 * ncap) * 1e3
print(f"graph replay: {gus:.1f} us/call  {byts / gus / 1e3:.0f} GB/s")
print(f"dbg={os.environ.get('DLLM_GEMV_DBG', '0')} K={K} N={N} bits={bits} M={M} path={path}: {us:.1f} us/call  "
      f"kernel-only {kus:.1f} us  {byts / kus / 1e3:.0f} GB/s", flush=True)
