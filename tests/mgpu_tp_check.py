"""Run under torchrun (one rank per GPU).  Prints `TP_CHECK_OK`.

1. Tensor-parallel stack (NCCL at the layer boundaries) against the ORACLE's f64-accumulated stack built from the oracle's
   own quantization of the unsharded weights (diffuse-llm-rs/src/lib.rs:806-813 composed with quantization.rs:81-85):
   f32 SIMT path, tcgen05 path (bf16 all-reduce), GEMV path; then the tcgen05 path at 2048 tokens with the all-reduces
   overlapped on the communication stream (token chunks) and with the overlap switched off; then the same with the
   library's own peer-to-peer all-reduce kernel over NVLink instead of NCCL.
2. A stack whose LAST layer is column-parallel, through dllm_denoise_step_dev: the all-gather lands in the step's own
   noise_pred buffer (it used to be staged in that same buffer).
3. Row-sharded KV quantize: per-token (D) rows are independent; per-tensor (B) all-reduces min / max — the concatenated
   codes and the parameters must equal the single-GPU quantization bit for bit (quantization.rs:140-157).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-llm-rs_b200"))
import dllm_b200  # noqa: E402
from dllm_b200 import QWeight, parallel as P  # noqa: E402
from dllm_b200.diffuse_llm import DiffusionConfig, QuantizedDiffusionModel  # noqa: E402
from oracle import pyoracle as O  # noqa: E402  (the checker)

F = np.float32


def oracle_stack64(x_tokens, ws, bs):
    h = np.asarray(x_tokens, np.float64)
    for w, b in zip(ws, bs):
        c, s, z = O.quantize_weight_grouped(w, 4, 128)
        wd = O.dequantize_weight_grouped(c, s, z, 128).astype(np.float64)
        h = h @ wd + (b.astype(np.float64) if b is not None else 0.0)
    return h


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = dllm_b200.Context(local)
    lib = ctx._lib
    rng = np.random.default_rng(123)                      # same on every rank
    hidden = 512
    dims = [hidden, 1024, hidden, hidden, hidden]
    shapes = list(zip(dims[:-1], dims[1:]))
    ws = [(rng.standard_normal(s) / np.sqrt(s[0])).astype(F) for s in shapes]
    bs = [(rng.standard_normal(s[1]) * 0.1).astype(F) for s in shapes]
    x_all = rng.standard_normal((32, hidden * 64)).astype(F)   # up to 2048 tokens
    plan = P.tp_plan(shapes, world)
    assert plan == [P.COLUMN, P.ROW, P.COLUMN, P.ROW], plan
    tpg = P.TensorParallelGroup(ctx, rank, world)
    tpg.init_nccl()
    ok = True

    def sharded_model(path):
        layers = []
        for w, b, mode in zip(ws, bs, plan):
            wsh, bsh = P.shard_weight(w, b, mode, rank, world)
            layers.append(QWeight.quantize(ctx, np.ascontiguousarray(wsh), 4, 128, bsh))
        model = QuantizedDiffusionModel(layers, hidden, ctx=ctx, path=path)
        tpg.set_plan(model, plan)
        return model

    # (path, tolerance vs the f64 oracle, batch rows, seq): SIMT f32; tcgen05 with bf16 all-reduces; GEMV (AUTO at <= 16 tokens)
    for path, tol, rows, seq in ((dllm_b200.PATH_SIMT, 1e-5, 4, 64), (dllm_b200.PATH_UMMA, 2e-2, 4, 64), (dllm_b200.PATH_AUTO, 2e-3, 1, 8)):
        x = np.ascontiguousarray(x_all[:rows, :hidden * seq])
        model = sharded_model(path)
        y = model.forward(x)
        ref = oracle_stack64(x.reshape(-1, hidden), ws, bs).reshape(rows, -1)
        rel = np.linalg.norm(y - ref) / np.linalg.norm(ref)
        print(f"rank {rank} path {path} rel-vs-oracle {rel:.3e}", flush=True)
        ok = ok and rel <= tol and np.all(np.isfinite(y))
        model.close()

    # 2048 tokens on the tcgen05 path: overlapped all-reduces (2 and 4 token chunks, 8 SMs left to NCCL) and no overlap
    x = x_all
    ref = oracle_stack64(x.reshape(-1, hidden), ws, bs).reshape(32, -1)
    model = sharded_model(dllm_b200.PATH_UMMA)
    for chunks, reserve in ((2, 8), (4, 8), (1, 0)):
        ctx._ck(lib.dllm_tp_configure(ctx.h, chunks, reserve, 0))
        y = model.forward(x)
        rel = np.linalg.norm(y - ref) / np.linalg.norm(ref)
        print(f"rank {rank} overlap chunks={chunks} rel-vs-oracle {rel:.3e}", flush=True)
        ok = ok and rel <= 2e-2 and np.all(np.isfinite(y))
    ctx._ck(lib.dllm_tp_configure(ctx.h, 0, -1, 0))
    model.close()
    # a wider stack with the library's own all-reduce over NVLink peer memory (CUDA-IPC arena, csrc/tp.cu) instead of NCCL:
    # same oracle bound, bit-identical on all ranks (every element is summed by exactly one rank), and identical bits to the
    # NCCL result of the same placement is NOT required (different summation order is allowed) — only the oracle bound is
    if tpg.enable_p2p(2048, 4096):
        # wide enough for the dense CTA-pair kernel, so that the row-parallel layers run with the reduce-scatter fused into the
        # GEMM epilogue (bulk tensor stores into the owners' receive buffers over NVLink) + the reduce / all-gather kernel
        hid2 = 2048
        dims2 = [hid2, 4096, hid2, 4096, hid2]
        shapes2 = list(zip(dims2[:-1], dims2[1:]))
        ws_b = [(rng.standard_normal(s_) / np.sqrt(s_[0])).astype(F) for s_ in shapes2]
        bs_b = [(rng.standard_normal(s_[1]) * 0.1).astype(F) for s_ in shapes2]
        plan_b = P.tp_plan(shapes2, world)
        assert plan_b == [P.COLUMN, P.ROW, P.COLUMN, P.ROW], plan_b
        xb = rng.standard_normal((4, hid2 * 512)).astype(F)          # 2048 tokens
        ref_b = oracle_stack64(xb.reshape(-1, hid2), ws_b, bs_b).reshape(4, -1)
        layers_b = []
        for w_, b_, mode in zip(ws_b, bs_b, plan_b):
            wsh, bsh = P.shard_weight(w_, b_, mode, rank, world)
            layers_b.append(QWeight.quantize(ctx, np.ascontiguousarray(wsh), 4, 128, bsh))
        model_b = QuantizedDiffusionModel(layers_b, hid2, ctx=ctx, path=dllm_b200.PATH_UMMA)
        tpg.set_plan(model_b, plan_b)
        # gated = the all-gather half runs under the NEXT GEMM (its activation loads wait on per-source arrival counters)
        #         (2: the reduce / gather kernel runs on the communication stream under that GEMM; 1: in front of it; 0: closing barrier)
        for fused, gated, chunks, reserve in (("1", "2", 1, 0), ("1", "1", 1, 0), ("1", "0", 1, 0), ("1", "2", 2, 16), ("0", "0", 1, 0), ("0", "0", 2, 16)):
            os.environ["DLLM_TP_FUSED_RS"] = fused
            os.environ["DLLM_TP_GATED"] = gated
            ctx._ck(lib.dllm_tp_configure(ctx.h, chunks, reserve, 0))
            n0 = tpg.p2p_status()["allreduces"]
            y = model_b.forward(xb)
            rel = np.linalg.norm(y - ref_b) / np.linalg.norm(ref_b)
            st = tpg.p2p_status()
            t = torch.from_numpy(y.view(np.int32).astype(np.int64)).sum().reshape(1).cuda()
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            print(f"rank {rank} p2p fused_rs={fused} gated={gated} chunks={chunks} rel-vs-oracle {rel:.3e} kernel calls {st['allreduces'] - n0} same-bits {bool(lo.item() == hi.item())}", flush=True)
            ok = ok and rel <= 2e-2 and np.all(np.isfinite(y)) and st["allreduces"] > n0 and st["timed_out"] == 0 and lo.item() == hi.item()
        ctx._ck(lib.dllm_tp_configure(ctx.h, 0, -1, 0))
        os.environ.pop("DLLM_TP_FUSED_RS", None)
        os.environ.pop("DLLM_TP_GATED", None)
        model_b.close()
        tpg.disable_p2p()
    else:
        print(f"rank {rank} p2p arena unavailable (CUDA IPC): NCCL path only", flush=True)

    # a stack that ENDS with a column-parallel layer, through dllm_denoise_step_dev: the all-gather's destination is the
    # step's noise_pred buffer
    ws2 = [ws[0], ws[1], (rng.standard_normal((hidden, hidden)) / np.sqrt(hidden)).astype(F)]
    plan2 = [P.COLUMN, P.ROW, P.COLUMN]
    layers = []
    for w, mode in zip(ws2, plan2):
        wsh, _ = P.shard_weight(w, None, mode, rank, world)
        layers.append(QWeight.quantize(ctx, np.ascontiguousarray(wsh), 4, 128, None))
    cfg = DiffusionConfig(num_timesteps=50, hidden_size=hidden, use_kv_cache=False)
    model = QuantizedDiffusionModel(layers, hidden, cfg, ctx, dllm_b200.PATH_SIMT)
    tpg.set_plan(model, plan2)
    xs = np.ascontiguousarray(x_all[:3, :hidden * 4])
    zs = rng.standard_normal(xs.shape).astype(F)
    n = xs.size
    dx, dz = ctx.malloc(n * 4), ctx.malloc(n * 4)
    ctx.h2d(dx, xs)
    ctx.h2d(dz, zs)
    model.denoise_step_dev(dx, dz, 7, 3, hidden * 4)
    got = ctx.d2h(dx, xs.shape, F)
    pred = oracle_stack64(xs.reshape(-1, hidden), ws2, [None] * 3).reshape(3, -1).astype(F)
    exp = O.p_sample(xs, pred, zs, [7, 7, 7], O.beta_schedule(O.BETA_LINEAR, 50), True)
    rel = np.linalg.norm(got - exp) / np.linalg.norm(exp)
    print(f"rank {rank} trailing-column denoise step rel-vs-oracle {rel:.3e}", flush=True)
    ok = ok and rel <= 1e-4
    ctx.free(dx)
    ctx.free(dz)
    model.close()

    # row-sharded KV quantize: concatenation over the ranks == single-GPU result, bit for bit
    Lk, S, Hd = 3, 64 * world, 256
    keys = rng.standard_normal((Lk, S, Hd)).astype(F)
    vals = (rng.standard_normal((Lk, S, Hd)) * 3 + 1).astype(F)
    s0, s1 = P.dp_partition(S, rank, world)
    kl, vl = np.ascontiguousarray(keys[:, s0:s1]), np.ascontiguousarray(vals[:, s0:s1])
    nloc = kl.size
    dk, dv = ctx.malloc(nloc * 4), ctx.malloc(nloc * 4)
    ctx.h2d(dk, kl)
    ctx.h2d(dv, vl)
    for scheme in (dllm_b200.KV_ROW_D, dllm_b200.KV_TENSOR_B):
        for bits in (8, 4):
            h = C.c_void_p()
            ctx._ck(lib.dllm_kv_quantize_sharded_dev(ctx.h, dk, dv, Lk, s1 - s0, Hd, bits, scheme, C.byref(h)))
            rows = Lk * (s1 - s0)
            nsc = rows if scheme == dllm_b200.KV_ROW_D else 1
            kc, vc = np.empty(nloc, np.uint8), np.empty(nloc, np.uint8)
            ksc, kzp, vsc, vzp = (np.empty(nsc, F) for _ in range(4))
            ctx._ck(lib.dllm_kv_export(ctx.h, h, kc.ctypes.data, vc.ctypes.data, ksc.ctypes.data, kzp.ctypes.data,
                                       vsc.ctypes.data, vzp.ctypes.data))
            lib.dllm_kv_destroy(h)
            for full, codes, sc, zp in ((keys, kc, ksc, kzp), (vals, vc, vsc, vzp)):
                if scheme == dllm_b200.KV_TENSOR_B:
                    ec, es, ez = O.quantize_tensor(full.ravel(), bits)          # ONE scale / zero-point for the whole tensor
                    ec = ec.reshape(full.shape)[:, s0:s1].ravel()
                    good = np.array_equal(codes, ec) and sc[0] == es and zp[0] == ez
                else:
                    ec, es, ez = O.quantize_d_rows(full[:, s0:s1].reshape(-1, Hd), [bits])
                    good = np.array_equal(codes, ec.ravel()) and np.array_equal(sc, es) and np.array_equal(zp, ez)
                if not good:
                    print(f"rank {rank} KV scheme {scheme} bits {bits}: MISMATCH", flush=True)
                ok = ok and good
    ctx.free(dk)
    ctx.free(dv)

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    tpg.close()
    dist.destroy_process_group()
    if rank == 0:
        print("TP_CHECK_OK" if flag.item() == 1 else "TP_CHECK_FAILED", flush=True)
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
