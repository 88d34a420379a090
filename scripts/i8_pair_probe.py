"""Two-layer int8-mode stack [K -> N -> K] at M tokens (layer 1 leaves through the staged bf16 epilogue, layer 2 through the f32 one):
python scripts/i8_pair_probe.py K N M   (DLLM_UMMA_DBG=128 dumps the stage timelines of the launches to gpurun_out/pair2_i8_trace_*.csv)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-llm-rs_b200"))
import torch
import dllm_b200
from dllm_b200 import QWeight, PATH_I8, PATH_UMMA
from dllm_b200.diffuse_llm import DiffusionConfig, QuantizedDiffusionModel

K, N, M = map(int, sys.argv[1:4])
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
lay = []
for (k, n) in ((K, N), (N, K)):
    w = torch.randn(k, n, device="cuda") / k ** 0.5
    torch.cuda.synchronize()
    lay.append(QWeight.quantize_dev(ctx, w.data_ptr(), k, n, 4, 0))
    ctx.sync()
x = torch.randn(M, K, device="cuda")
y = torch.empty(M, K, device="cuda")
torch.cuda.synchronize()
for name, pth in (("int8", PATH_I8), ("bf16", PATH_UMMA)):
    mdl = QuantizedDiffusionModel(lay, K, DiffusionConfig(hidden_size=K), ctx, pth)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 2 if os.environ.get("DLLM_UMMA_DBG") else 20
    with torch.cuda.stream(stream):
        for _ in range(2):
            mdl.forward_dev(x.data_ptr(), 1, M * K, y.data_ptr())
        stream.synchronize()
        e0.record(stream)
        for _ in range(n):
            mdl.forward_dev(x.data_ptr(), 1, M * K, y.data_ptr())
        e1.record(stream)
        e1.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{name}: K={K} N={N} M={M}: {us:.1f} us per two-layer pass, {4.0 * M * K * N / us / 1e6:.0f} T(FL)OP/s incl. casts / quantizers", flush=True)
    mdl.close()
