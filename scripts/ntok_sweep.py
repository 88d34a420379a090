"""Dense CTA-pair kernel: time per launch for every candidate tile width (tokens per tile) on the layer shapes of the 1B- and
7B-class stacks and their tensor-parallel shards, at the token counts a GPU sees under token split / chunked overlap.
Calibrates pair2_pick_ntok's cost model (umma_gemm.cu).  Output: one JSON line per (shape, M)."""
import sys, os, json
sys.path.insert(0, os.path.join(os.getcwd(), "diffusion-llm-rs_b200"))
os.environ["DLLM_PROBE_BF16_OUT"] = "1"
import torch, dllm_b200, ctypes as C
from dllm_b200 import QWeight, PATH_UMMA
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
shapes = [(2048, 2048), (2048, 8192), (8192, 2048), (4096, 4096), (4096, 14336), (14336, 4096),
          (4096, 2048), (2048, 4096), (4096, 7168), (7168, 4096), (4096, 512), (512, 4096), (4096, 1792), (1792, 4096)]
Ms = [8192, 4096, 2048, 1024]
if len(sys.argv) > 1:
    Ms = [int(v) for v in sys.argv[1].split(",")]
for K, N in shapes:
    w = torch.randn(K, N, device="cuda") * 0.02
    qw = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, 4, 128)
    ctx.sync()
    for M in Ms:
        x = torch.randn(M, K, device="cuda"); y = torch.empty(M, N, device="cuda"); torch.cuda.synchronize()
        res = {}
        for nt in (0, 128, 160, 192, 224, 256):
            if nt: os.environ["DLLM_UMMA_NTOK2"] = str(nt)
            else: os.environ.pop("DLLM_UMMA_NTOK2", None)
            with torch.cuda.stream(stream):
                for i in range(3): qw.forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_UMMA)
                stream.synchronize()
                ctx._ck(ctx._lib.dllm_profile_begin(ctx.h))           # CUDA events around every dense-kernel launch
                for i in range(20): qw.forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_UMMA)
            nl, ms, fl, by = C.c_uint64(), C.c_double(), C.c_double(), C.c_double()
            ctx._ck(ctx._lib.dllm_profile_end(ctx.h, C.byref(nl), C.byref(ms), C.byref(fl), C.byref(by)))
            res["auto" if nt == 0 else str(nt)] = round(ms.value / nl.value * 1e3, 1)
        best = min((v, k) for k, v in res.items() if k != "auto")
        print(json.dumps({"K": K, "N": N, "M": M, "us": res, "best": best[1], "auto_over_best": round(res["auto"] / best[0], 3),
                          "best_TFLOPs": round(2.0 * M * K * N / best[0] / 1e6)}), flush=True)
        del x, y
    qw.close(); del w
