import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "diffusion-llm-rs_b200"))
import torch, dllm_b200, ctypes as C
from dllm_b200 import QWeight, PATH_UMMA
K, N, bits, M = map(int, sys.argv[1:5])
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
w = torch.randn(K, N, device="cuda") * 0.02
qw = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128)
x = torch.randn(M, K, device="cuda"); y = torch.empty(M, N, device="cuda"); torch.cuda.synchronize()
with torch.cuda.stream(stream):
    for i in range(3): qw.forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_UMMA)
    stream.synchronize()
    ctx._ck(ctx._lib.dllm_profile_begin(ctx.h))
    for i in range(20): qw.forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_UMMA)
nl, ms, fl, by = C.c_uint64(), C.c_double(), C.c_double(), C.c_double()
ctx._ck(ctx._lib.dllm_profile_end(ctx.h, C.byref(nl), C.byref(ms), C.byref(fl), C.byref(by)))
print(f"pair={os.environ.get('DLLM_UMMA_PAIR','-')} ntok2={os.environ.get('DLLM_UMMA_NTOK2','-')} dbg={os.environ.get('DLLM_UMMA_DBG','0')} K={K} N={N} bits={bits} M={M}: kernel {ms.value/nl.value*1e3:.1f} us  {fl.value/ms.value/1e9:.0f} TFLOP/s", flush=True)
