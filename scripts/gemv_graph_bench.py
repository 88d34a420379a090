"""Dequant-GEMV (configs[1]) timed the way bench.py times it: CUDA-graph replay of 16 calls x 10 over a pool of weights
larger than the L2.  One JSON line per (K=N, bits, M).  Usage: python scripts/gemv_graph_bench.py [shapes] [bits] [Ms]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-llm-rs_b200"))
import torch, dllm_b200
from dllm_b200 import QWeight
shapes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "14336,8192,4096").split(",")]
bitss = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "4,2,8").split(",")]
Ms = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "1,4,8,16").split(",")]
try:
    hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6650.0
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
tag = {k: v for k, v in os.environ.items() if k.startswith("DLLM_GEMV") or k.startswith("DLLM_BENCH")}
PATH = int(os.environ.get("DLLM_BENCH_PATH", dllm_b200.PATH_GEMV))
for KN in shapes:
    K = N = KN
    w = torch.randn(K, N, device="cuda") * 0.02
    torch.cuda.synchronize()
    for bits in bitss:
        wbytes = K * N * bits // 8
        npool = int(os.environ["DLLM_GEMV_POOL"]) if os.environ.get("DLLM_GEMV_POOL") else max(4, -(-400_000_000 // wbytes))
        pool = [QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128) for _ in range(npool)]
        ctx.sync()
        for M in Ms:
            x = torch.randn(M, K, device="cuda"); y = torch.empty(M, N, device="cuda")
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                for i in range(npool):
                    pool[i].forward_dev(x.data_ptr(), M, y.data_ptr(), PATH)
                stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for i in range(16):
                    pool[i % npool].forward_dev(x.data_ptr(), M, y.data_ptr(), PATH)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                g.replay(); stream.synchronize()
                e0.record(stream)
                for _ in range(10):
                    g.replay()
                e1.record(stream); e1.synchronize()
            us = e0.elapsed_time(e1) / 160 * 1e3
            byts = wbytes + (K // 128) * N * 8 + 4 * M * K + 4 * M * N
            print(json.dumps({"K": K, "bits": bits, "M": M, "us": round(us, 2), "GBps": round(byts / us / 1e3, 1),
                              "hbm_frac": round(byts / us / 1e3 / hbm, 3), **tag}), flush=True)
            del g, x, y
        for p in pool:
            p.close()
    del w
