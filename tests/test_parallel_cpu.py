"""Host-side multi-GPU logic on CPU: world_size-2 `gloo` processes drive the same sharding plan,
collective placement and unique-id broadcast the NCCL path uses (dllm_b200/parallel.py)."""
import os
import socket

import numpy as np
import pytest

from dllm_b200 import parallel as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_dp_partition_covers_everything_once():
    for n in (0, 1, 7, 32, 1048576):
        for world in (1, 2, 3, 8):
            spans = [P.dp_partition(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_tp_plan_pairs_and_group_boundaries():
    H, F = 4096, 14336
    shapes = [(H, H)] * 4 + [(H, F), (F, H)]
    for world in (2, 4, 8):
        plan = P.tp_plan(shapes, world)
        assert plan == [P.COLUMN, P.ROW, P.COLUMN, P.ROW, P.COLUMN, P.ROW]
    # K in {4096, 8192, 14336} shards on a 128-group boundary at p in {2,4,8}
    for K in (4096, 8192, 14336):
        for world in (2, 4, 8):
            assert P.tp_plan([(K, K), (K, K)], world) == [P.COLUMN, P.ROW]
    # F = 11008: 11008/2 = 43*128 shards cleanly, 11008/4 = 2752 would split a 128-group -> stays replicated
    assert P.tp_plan([(H, 11008), (11008, H)], 2) == [P.COLUMN, P.ROW]
    assert P.tp_plan([(H, 11008), (11008, H)], 4) == [P.REPLICATED] * 2
    assert P.tp_plan([(H, 11008), (11008, H)], 8) == [P.REPLICATED] * 2
    assert P.tp_plan(shapes, 1) == [P.REPLICATED] * 6
    assert P.tp_plan([(256, 200), (200, 256)], 2) == [P.REPLICATED] * 2   # 100 is not a group multiple


def test_shard_weight_shapes_and_bias_placement():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((256, 512))
    b = rng.standard_normal(512)
    cols = [P.shard_weight(w, b, P.COLUMN, r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate([c[0] for c in cols], axis=1), w)
    assert np.array_equal(np.concatenate([c[1] for c in cols]), b)
    rows = [P.shard_weight(w, b, P.ROW, r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate([r_[0] for r_ in rows], axis=0), w)
    assert np.array_equal(sum(r_[1] for r_ in rows), b)          # bias only on rank 0


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. unique-id broadcast: every rank ends up with rank 0's 128 bytes
        uid = P.broadcast_unique_id(lambda: bytes(range(128)), rank, world)
        assert uid == bytes(range(128))

        # 2. sharded forward == unsharded forward (same collective placement as csrc/api.cu)
        rng = np.random.default_rng(7)        # same seed on every rank: replicated inputs
        dims = [256, 512, 256, 256, 256, 384]  # 384 leaves a trailing column-parallel layer to gather
        shapes = list(zip(dims[:-1], dims[1:]))
        ws = [rng.standard_normal(s) / np.sqrt(s[0]) for s in shapes]
        bs = [rng.standard_normal(s[1]) * 0.1 for s in shapes]
        x = rng.standard_normal((16, 256))
        plan = P.tp_plan(shapes, world) if world > 1 else [0] * len(shapes)
        plan[-1] = P.COLUMN if (dims[-1] // world) % 64 == 0 else plan[-1]

        def all_reduce(y):
            t = torch.from_numpy(np.ascontiguousarray(y))
            dist.all_reduce(t)
            return t.numpy()

        def all_gather(y):
            t = torch.from_numpy(np.ascontiguousarray(y))
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return np.concatenate([o.numpy() for o in outs], axis=1)

        y = P.simulate_tp_forward(x, ws, bs, plan, rank, world, all_reduce, all_gather)
        ref = x
        for w, b in zip(ws, bs):
            ref = ref @ w + b
        assert y.shape == ref.shape and np.allclose(y, ref, rtol=1e-10, atol=1e-10)

        # 3. data-parallel bookkeeping: whole-job throughput = sum of per-rank units / max time
        b0, b1 = P.dp_partition(32, rank, world)
        t = torch.tensor([float(b1 - b0), 1.0 + rank])
        units = t[:1].clone()
        dist.all_reduce(units)
        tmax = t[1:].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        assert units.item() == 32 and tmax.item() == float(world)
        q.put((rank, "ok", plan))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"fail: {e!r}", None))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] == "ok" for r in res), res
    assert res[0][2][:4] == [P.COLUMN, P.ROW, P.COLUMN, P.ROW]
