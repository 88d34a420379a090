"""Kernel-level roofline sweep (CUDA events on the launching stream, L2-busting working sets).
Prints one JSON line per kernel/shape: algorithmic bytes / time vs MEASURED_PEAKS.
Usage: python scripts/microbench.py [--quick]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-llm-rs_b200"))
import dllm_b200  # noqa: E402
from dllm_b200 import QWeight, PATH_SIMT, PATH_UMMA, PATH_GEMV  # noqa: E402


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["hbm_gbs"], p["bf16_tflops"], "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


def time_fn(stream, fn, iters, warmup=3):
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            fn()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        start.record(stream)
        for _ in range(iters):
            fn()
        end.record(stream)
        end.synchronize()
    return start.elapsed_time(end) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--umma", action="store_true", help="also time the tcgen05 path")
    ap.add_argument("--dense", action="store_true", help="only the dense (M=8192) tcgen05 shapes")
    ap.add_argument("--gemv-only", action="store_true")
    ap.add_argument("--simt", action="store_true", help="also time the f32-faithful SIMT path")
    ap.add_argument("--shapes", default="4096,8192,14336")
    ap.add_argument("--bits", default="4,2,8")
    ap.add_argument("--ms", default="1,4,16")
    args = ap.parse_args()
    hbm, tf, src = peaks()
    stream = torch.cuda.Stream()
    ctx = dllm_b200.Context(0, stream=stream.cuda_stream)
    out = []

    def report(name, secs, bytes_, flops=0, **kw):
        r = {"kernel": name, "us": round(secs * 1e6, 2), "GBps": round(bytes_ / secs / 1e9, 1),
             "hbm_frac": round(bytes_ / secs / 1e9 / hbm, 3), "peak": src}
        if flops:
            r["TFLOPs"] = round(flops / secs / 1e12, 1)
            r["tensor_frac"] = round(flops / secs / 1e12 / tf, 3)
        r.update(kw)
        print(json.dumps(r), flush=True)
        out.append(r)

    if args.dense:
        M = 8192
        for (K, N) in ((2048, 2048), (2048, 8192), (8192, 2048), (4096, 4096), (4096, 11008), (11008, 4096)):
            w = torch.randn(K, N, device="cuda") * 0.02
            for bits in (4, 8, 2):
                qw = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128)
                xin = torch.randn(M, K, device="cuda")
                y = torch.empty(M, N, device="cuda")
                torch.cuda.synchronize()
                t = time_fn(stream, lambda: qw.forward_dev(xin.data_ptr(), M, y.data_ptr(), PATH_UMMA), 10)
                report(f"dense_umma_{bits}b(+f32->bf16 cast)", t, K * N * bits // 8 + 6 * M * K + 4 * M * N, 2.0 * M * K * N, K=K, N=N, M=M)
                qw.close()
        return
    rows, dim = ((1 << 16) if not args.quick else (1 << 14)), 4096
    if args.gemv_only:
        rows = 0
    if rows:
        # ---- KV quantizers: rows x 4096 f32 ----
        n = rows * dim
        x = torch.randn(rows, dim, device="cuda")
        codes = torch.empty(n, dtype=torch.uint8, device="cuda")
        scales, zps = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
        params = torch.empty(4, device="cuda")
        deq = torch.empty_like(x)
        torch.cuda.synchronize()
        for bits in ((8, 4, 2) if rows else ()):
            t = time_fn(stream, lambda: ctx.quantize_d_rows_dev(x.data_ptr(), rows, dim, bits, True, codes.data_ptr(), scales.data_ptr(), zps.data_ptr()), 10)
            report(f"kv_quant_rows_D_{bits}b", t, 4 * n + n * bits // 8 + 8 * rows, rows=rows, dim=dim)
            t = time_fn(stream, lambda: ctx.dequantize_d_rows_dev(codes.data_ptr(), rows, dim, bits, True, scales.data_ptr(), zps.data_ptr(), deq.data_ptr()), 10)
            report(f"kv_dequant_rows_D_{bits}b", t, 4 * n + n * bits // 8 + 8 * rows, rows=rows, dim=dim)
            t = time_fn(stream, lambda: ctx.quantize_tensor_dev(x.data_ptr(), n, bits, True, codes.data_ptr(), params.data_ptr()), 10)
            report(f"quant_tensor_B_{bits}b(2 passes)", t, 8 * n + n * bits // 8, n=n)
            t = time_fn(stream, lambda: ctx.dequantize_tensor_dev(codes.data_ptr(), n, bits, True, params.data_ptr(), deq.data_ptr()), 10)
            report(f"dequant_tensor_B_{bits}b", t, 4 * n + n * bits // 8, n=n)
        if rows:
            ctx.quantize_tensor_dev(x.data_ptr(), n, 8, False, codes.data_ptr(), params.data_ptr())
        packed = torch.empty(max(n, 16), dtype=torch.uint8, device="cuda")
        for bits in ((4, 1) if rows else ()):
            t = time_fn(stream, lambda: ctx.pack_dev(codes.data_ptr(), n, bits, packed.data_ptr()), 10)
            report(f"pack_{bits}b", t, n + n * bits // 8, n=n)
            t = time_fn(stream, lambda: ctx.unpack_dev(packed.data_ptr(), n, bits, codes.data_ptr()), 10)
            report(f"unpack_{bits}b", t, n + n * bits // 8, n=n)
        del x, codes, deq, packed


    # ---- dequant-GEMV sweep (BASELINE configs[1]); weight pool rotated so loads come from HBM ----
    shapes = [int(v) for v in args.shapes.split(",")] if not args.quick else [4096]
    for KN in shapes:
        K = N = KN
        w = torch.randn(K, N, device="cuda") * 0.02
        for bits in [int(v) for v in args.bits.split(",")]:
            wbytes = K * N * bits // 8
            pool_n = max(2, min(8, int(512e6 // wbytes) + 1))
            pool = [QWeight.quantize_dev(ctx, w.data_ptr(), K, N, bits, 128) for _ in range(pool_n)]
            ctx.sync()
            for M in ([int(v) for v in args.ms.split(",")] if not args.quick else (1,)):
                xin = torch.randn(M, K, device="cuda")
                y = torch.empty(M, N, device="cuda")
                torch.cuda.synchronize()
                for path, pname in ((PATH_GEMV, "mma"),) + (((PATH_SIMT, "simt"),) if args.simt else ()) + (((PATH_UMMA, "umma"),) if args.umma else ()):
                    state = {"i": 0}

                    def step():
                        pool[state["i"] % pool_n].forward_dev(xin.data_ptr(), M, y.data_ptr(), path)
                        state["i"] += 1
                    try:
                        t = time_fn(stream, step, 20 * pool_n // 2)
                    except Exception as e:  # noqa: BLE001
                        print(json.dumps({"kernel": f"gemv_{pname}", "error": str(e)[:200]}))
                        continue
                    bytes_ = wbytes + (K // 128) * N * 8 + 4 * M * K + 4 * M * N
                    extra = {}
                    if path != PATH_SIMT:
                        # the dominant kernel alone: every launch bracketed by CUDA events on its stream
                        import ctypes as C
                        ctx._ck(ctx._lib.dllm_profile_begin(ctx.h))
                        with torch.cuda.stream(stream):
                            for _ in range(4 * pool_n):
                                step()
                        nl, ms, fl, by = C.c_uint64(), C.c_double(), C.c_double(), C.c_double()
                        ctx._ck(ctx._lib.dllm_profile_end(ctx.h, C.byref(nl), C.byref(ms), C.byref(fl), C.byref(by)))
                        kus = ms.value / max(nl.value, 1) * 1e3
                        extra = {"kernel_us": round(kus, 2), "kernel_GBps": round(bytes_ / kus / 1e3, 1),
                                 "kernel_hbm_frac": round(bytes_ / kus / 1e3 / hbm, 3)}
                    report(f"gemv_{pname}_{bits}b", t, bytes_, 2.0 * M * K * N, K=K, N=N, M=M, pool=pool_n, **extra)
            for p in pool:
                p.close()
        del w
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
