"""Time the exact int8 tcgen05 linear against the bf16 tcgen05 linear on one shape: python scripts/i8_probe.py K N bits M"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-llm-rs_b200"))
import torch
import dllm_b200
from dllm_b200 import QWeight, PATH_UMMA

K, N, bits, M = map(int, sys.argv[1:5])
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
w = torch.randn(K, N, device="cuda") * 0.02
torch.cuda.synchronize()
qt = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 0)       # per tensor: the int8 path's scheme
qg = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128)
xq = torch.randint(-128, 128, (M, K), device="cuda", dtype=torch.int8)
xf = torch.randn(M, K, device="cuda")
yi = torch.empty(M, N, device="cuda", dtype=torch.int32)
yf = torch.empty(M, N, device="cuda")
torch.cuda.synchronize()


def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        stream.synchronize()
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


us_i8 = timeit(lambda: qt.forward_i8_dev(xq.data_ptr(), M, yi.data_ptr()))
us_bf = timeit(lambda: qg.forward_dev(xf.data_ptr(), M, yf.data_ptr(), PATH_UMMA))
ops = 2.0 * M * K * N
# spot check against torch on a slice (exactness is tested in tests/)
codes, scales, zps = qt.export()
ref = (xq[:8].cpu().to(torch.int64) @ (torch.from_numpy(codes).to(torch.int64) - int(zps.ravel()[0])))
ok = bool(torch.equal(yi[:8].cpu().to(torch.int64), ref))
print(f"K={K} N={N} bits={bits} M={M}: int8 {us_i8:.1f} us ({ops / us_i8 / 1e6:.0f} TOP/s)  bf16 path (incl. f32->bf16 cast) {us_bf:.1f} us "
      f"({ops / us_bf / 1e6:.0f} TFLOP/s)  exact={ok}", flush=True)
