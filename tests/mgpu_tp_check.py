"""Run under torchrun (one rank per GPU): tensor-parallel stack (NCCL at the layer boundaries) against
the unsharded stack on rank 0, plus the data-parallel split of a KV quantize.  Prints `TP_CHECK_OK`."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-llm-rs_b200"))
import dllm_b200  # noqa: E402
from dllm_b200 import QWeight, parallel as P  # noqa: E402
from dllm_b200.diffuse_llm import QuantizedDiffusionModel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = dllm_b200.Context(local)
    rng = np.random.default_rng(123)                      # same on every rank
    hidden = 512
    dims = [hidden, 1024, hidden, hidden, hidden]
    shapes = list(zip(dims[:-1], dims[1:]))
    ws = [(rng.standard_normal(s) / np.sqrt(s[0])).astype(np.float32) for s in shapes]
    bs = [(rng.standard_normal(s[1]) * 0.1).astype(np.float32) for s in shapes]
    x = rng.standard_normal((4, hidden * 64)).astype(np.float32)   # 256 tokens
    plan = P.tp_plan(shapes, world)
    assert plan == [P.COLUMN, P.ROW, P.COLUMN, P.ROW], plan
    tpg = P.TensorParallelGroup(ctx, rank, world)
    tpg.init_nccl()
    ok = True
    x_full = x
    # (path, tolerance, tokens): f32 SIMT; tcgen05 with bf16 all-reduce at the row-parallel boundaries; and the
    # HBM-bound GEMV kernel (AUTO at <= 16 tokens) with f32 all-reduce
    for path, tol, rows in ((dllm_b200.PATH_SIMT, 1e-5, 4), (dllm_b200.PATH_UMMA, 2e-2, 4), (dllm_b200.PATH_AUTO, 2e-3, 1)):
        x = x_full[:rows, :hidden * (64 if rows == 4 else 8)]            # 256 tokens / 8 tokens
        layers = []
        for w, b, mode in zip(ws, bs, plan):
            wsh, bsh = P.shard_weight(w, b, mode, rank, world)
            layers.append(QWeight.quantize(ctx, np.ascontiguousarray(wsh), 4, 128, bsh))
        model = QuantizedDiffusionModel(layers, hidden, ctx=ctx, path=path)
        tpg.set_plan(model, plan)
        y = model.forward(x)
        full = QuantizedDiffusionModel([QWeight.quantize(ctx, w, 4, 128, b) for w, b in zip(ws, bs)], hidden, ctx=ctx,
                                       path=dllm_b200.PATH_SIMT)
        ref = full.forward(x)
        rel = np.linalg.norm(y - ref) / np.linalg.norm(ref)
        print(f"rank {rank} path {path} rel {rel:.3e}", flush=True)
        ok = ok and rel <= tol and np.all(np.isfinite(y))
        model.close()
        full.close()
    # data parallel: every rank quantizes its own rows of a KV tensor; concatenation == single-GPU result
    kv = rng.standard_normal((64, 256)).astype(np.float32)
    b0, b1 = P.dp_partition(64, rank, world)
    c, s, z = ctx.quantize_d_rows(kv[b0:b1], [4])
    cf, sf, zf = ctx.quantize_d_rows(kv, [4])
    ok = ok and np.array_equal(c, cf[b0:b1]) and np.array_equal(s, sf[b0:b1]) and np.array_equal(z, zf[b0:b1])
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    tpg.close()
    dist.destroy_process_group()
    if rank == 0:
        print("TP_CHECK_OK" if flag.item() == 1 else "TP_CHECK_FAILED", flush=True)
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
