"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous, NCCL over
NVLink 5 / NVSwitch for the data path.

Two ways the path shards:
  * data parallel — independent denoising batches / KV rows per rank, no data-path collective
    (`dp_partition`);
  * tensor parallel — linears alternate column-parallel (split N) and row-parallel (split K); a
    column->row pair needs exactly one collective (the all-reduce of the row-parallel partial sums)
    at the layer boundary (`tp_plan`, `shard_weight`, `TensorParallelGroup`).

Everything in this file except TensorParallelGroup.init_nccl is host logic and runs on CPU (the
gloo tests in tests/test_parallel_cpu.py drive it with world_size 2).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

REPLICATED, COLUMN, ROW = 0, 1, 2


def dp_partition(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the independent units (batches, KV rows) owned by `rank`."""
    return n_items * rank // world, n_items * (rank + 1) // world


def tp_plan(shapes: Sequence[Tuple[int, int]], world: int, group: int = 128) -> List[int]:
    """Pair consecutive linears [K,N],[N,N2] as column->row when both shard on a group boundary;
    anything left over stays replicated.  Returns one of REPLICATED/COLUMN/ROW per layer."""
    plan = [REPLICATED] * len(shapes)
    if world <= 1:
        return plan
    i = 0
    while i + 1 < len(shapes):
        (k0, n0), (k1, n1) = shapes[i], shapes[i + 1]
        ok = (n0 == k1 and n0 % world == 0 and (n0 // world) % max(group, 64) == 0)
        if ok:
            plan[i], plan[i + 1] = COLUMN, ROW
            i += 2
        else:
            i += 1
    return plan


def shard_weight(w: np.ndarray, bias: Optional[np.ndarray], mode: int, rank: int, world: int):
    """The [K,N] slice (and bias) this rank holds.  Row-parallel partial sums are all-reduced, so the
    bias is kept on rank 0 only; quantization groups never straddle a shard (tp_plan checks)."""
    K, N = w.shape
    if mode == COLUMN:
        n0, n1 = N * rank // world, N * (rank + 1) // world
        return w[:, n0:n1], (bias[n0:n1] if bias is not None else None)
    if mode == ROW:
        k0, k1 = K * rank // world, K * (rank + 1) // world
        b = bias if (bias is not None and rank == 0) else (np.zeros_like(bias) if bias is not None else None)
        return w[k0:k1, :], b
    return w, bias


def simulate_tp_forward(x: np.ndarray, weights, biases, plan, rank: int, world: int, all_reduce, all_gather):
    """f64 numpy model of the sharded stack, used by the CPU (gloo) tests: the same collective
    placement as model_forward_tokens in csrc/api.cu."""
    h = np.asarray(x, np.float64)
    for l, (w, b) in enumerate(zip(weights, biases)):
        ws, bs = shard_weight(np.asarray(w, np.float64), None if b is None else np.asarray(b, np.float64), plan[l], rank, world)
        y = h @ ws + (bs if bs is not None else 0.0)
        if plan[l] == ROW:
            y = all_reduce(y)
        elif plan[l] == COLUMN and (l + 1 == len(weights) or plan[l + 1] != ROW):
            y = all_gather(y)
        h = y
    return h


def broadcast_unique_id(make_id, rank: int, world: int) -> bytes:
    """Rank 0 creates the 128-byte NCCL unique id; torch.distributed (any backend) broadcasts it."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    if world > 1:
        dev = None
        if dist.get_backend() == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            buf = buf.to(dev)
        dist.broadcast(buf, src=0)
        buf = buf.cpu()
    return bytes(buf.tolist())


class TensorParallelGroup:
    """Owns the library-side NCCL communicator of one rank (dllm_tp_init)."""

    def __init__(self, ctx, rank: int, world: int):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.ready = False

    def init_nccl(self):
        lib = self.ctx._lib

        def make_id():
            raw = (C.c_uint8 * 128)()
            self.ctx._ck(lib.dllm_tp_unique_id(raw))
            return bytes(raw)

        uid = broadcast_unique_id(make_id, self.rank, self.world)
        arr = (C.c_uint8 * 128).from_buffer_copy(uid)
        self.ctx._ck(lib.dllm_tp_init(self.ctx.h, arr, self.rank, self.world))
        self.ready = True

    def set_plan(self, model, plan: Sequence[int]):
        arr = (C.c_int32 * len(plan))(*plan)
        self.ctx._ck(self.ctx._lib.dllm_model_set_parallel(self.ctx.h, model.h, arr, len(plan)))

    def enable_p2p(self, max_tokens: int, max_width: int) -> bool:
        """Switch the row-parallel all-reduces of tcgen05 stacks to the library's own NVLink kernel: a symmetric arena for
        the two bf16 activation buffers [max_tokens, max_width] (+ slack for chunk rounding).  Collective.  False (on every
        rank) when CUDA IPC / peer access is unavailable — the NCCL path then stays in use."""
        from . import _lib as L
        # three regions: two activation buffers + the receive buffer of the fused reduce-scatter
        need = 3 * ((max_tokens + 512) * max_width * 2 + 4096)
        try:
            self.ctx._ck(self.ctx._lib.dllm_tp_p2p_enable(self.ctx.h, need))
        except L.DllmError:
            return False
        return self.world > 1

    def disable_p2p(self):
        """Collective: release the arena; the all-reduces go back to NCCL."""
        self.ctx._ck(self.ctx._lib.dllm_tp_p2p_enable(self.ctx.h, 0))

    def p2p_status(self):
        a, b, n, t = C.c_void_p(), C.c_size_t(), C.c_uint64(), C.c_uint32()
        self.ctx._ck(self.ctx._lib.dllm_tp_p2p_status(self.ctx.h, C.byref(a), C.byref(b), C.byref(n), C.byref(t)))
        return {"arena": a.value or 0, "arena_bytes": b.value, "allreduces": n.value, "timed_out": t.value}

    def allreduce_dev(self, buf_dev: int, n: int):
        self.ctx._ck(self.ctx._lib.dllm_tp_allreduce_dev(self.ctx.h, buf_dev, n))

    def close(self):
        if self.ready:
            self.ctx._lib.dllm_tp_finalize(self.ctx.h)
            self.ready = False
