"""Tensor-parallel probe (run under torchrun, one rank per GPU): the library's own NVLink all-reduce against NCCL
(correctness on random data, 64 MiB timing), then the 7B-class step under every combination of
{NCCL, peer-to-peer} x {serial, token chunks 2 / 4 overlapped} with the exposed collective time.  Rank 0 prints JSON lines."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diffusion-llm-rs_b200"))
import torch, torch.distributed as dist
import dllm_b200
from dllm_b200 import QWeight, parallel as PAR
from dllm_b200.diffuse_llm import DiffusionConfig, QuantizedDiffusionModel
from bench import Timer, layer_shapes, BATCH, CANVAS

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(local, stream=stream.cuda_stream)
lib = ctx._lib
tm = Timer(torch, stream, world)
tpg = PAR.TensorParallelGroup(ctx, rank, world); tpg.init_nccl()
def say(**kw):
    if rank == 0: print(json.dumps(kw), flush=True)

model_name = sys.argv[1] if len(sys.argv) > 1 else "7b"
H, shapes = layer_shapes(model_name)
feat, tokens = CANVAS * H, BATCH * CANVAS
plan = PAR.tp_plan(shapes, world)
maxw = max(max((n // world if p == PAR.COLUMN else n), (k // world if p == PAR.ROW else k)) for (k, n), p in zip(shapes, plan))

# ---- NCCL all-reduce timing first (plain cudaMalloc buffer) ----
buf = torch.zeros(16 << 20, device="cuda")
ms = tm.run(lambda i: tpg.allreduce_dev(buf.data_ptr(), buf.numel()), 20, 3)
say(what="nccl_allreduce_64MiB_f32", ms=ms, algbw_GBps=buf.numel() * 4 / ms / 1e6)

ok = tpg.enable_p2p(tokens, max(maxw, H))
st = tpg.p2p_status()
say(what="p2p_enable", ok=ok, status=st, last_error=(lib.dllm_last_error(ctx.h) or b"").decode())
if ok:
    # the kernel alone on 64 MiB (f32 elements; the same bytes as one [8192, 4096] bf16 boundary tensor), then correctness
    n = 16 << 20
    ms = tm.run(lambda i: tpg.allreduce_dev(st["arena"], n), 20, 3)
    say(what="p2p_allreduce_64MiB", ms=ms, algbw_GBps=n * 4 / ms / 1e6, link_GBps_per_direction=n * 4 * 2 * (world - 1) / world / ms / 1e6)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    mine = torch.randn(1 << 20, device="cuda", generator=g)
    ref = mine.clone(); dist.all_reduce(ref)
    import numpy as np
    host = mine.cpu().numpy()
    ctx._ck(lib.dllm_memcpy_h2d(ctx.h, st["arena"], host.ctypes.data, host.nbytes))
    ctx.sync()
    tm.barrier()
    with torch.cuda.stream(stream):
        tpg.allreduce_dev(st["arena"], mine.numel()); stream.synchronize()
    out = np.empty_like(host)
    ctx._ck(lib.dllm_memcpy_d2h(ctx.h, out.ctypes.data, st["arena"], host.nbytes))
    ctx.sync()
    got = torch.from_numpy(out).cuda()
    tm.barrier()
    say(what="p2p_vs_nccl_f32", max_abs_diff=float((got - ref).abs().max()), rel=float((got - ref).norm() / ref.norm()))

# ---- the sharded 7B-class stack ----
wgen = torch.Generator(device="cuda").manual_seed(4242)
full, shard = [], []
for (K, N), mode in zip(shapes, plan):
    w = torch.randn(K, N, device="cuda", generator=wgen) * (1.0 / K ** 0.5)
    ws = w[:, N * rank // world: N * (rank + 1) // world].contiguous() if mode == PAR.COLUMN else \
         w[K * rank // world: K * (rank + 1) // world, :].contiguous() if mode == PAR.ROW else w
    torch.cuda.synchronize()
    shard.append(QWeight.quantize_dev(ctx, ws.data_ptr(), ws.shape[0], ws.shape[1], 4, 128))
    if rank == 0: full.append(QWeight.quantize_dev(ctx, w.data_ptr(), K, N, 4, 128))
    ctx.sync(); del w, ws
cfg = DiffusionConfig(num_timesteps=1000, hidden_size=H, use_kv_cache=False)
m_tp = QuantizedDiffusionModel(shard, H, cfg, ctx, dllm_b200.PATH_AUTO)
tpg.set_plan(m_tp, plan)
gen = torch.Generator(device="cuda").manual_seed(777)
x0 = torch.randn(BATCH, feat, device="cuda", generator=gen)
z = torch.randn(BATCH, feat, device="cuda", generator=gen)
x = x0.clone()
step = lambda i: m_tp.denoise_step_dev(x.data_ptr(), z.data_ptr(), 999 - (i % 900), BATCH, feat)

ms_single = 0.0
if rank == 0:
    m_full = QuantizedDiffusionModel(full, H, cfg, ctx, dllm_b200.PATH_AUTO)
    xs = x0.clone()
    ms_single = Timer(torch, stream, 1).run(lambda i: m_full.denoise_step_dev(xs.data_ptr(), z.data_ptr(), 999 - i, BATCH, feat), 3, 1)
tm.barrier()
ms_single = tm.max_over_ranks(ms_single)
say(what="single_gpu_step", ms=ms_single)

# correctness of the p2p path: forward with p2p on vs the unsharded stack on rank 0
def fwd(model, out):
    with torch.cuda.stream(stream):
        model.forward_dev(x0.data_ptr(), BATCH, feat, out.data_ptr()); stream.synchronize()
p_tp = torch.empty_like(x0); fwd(m_tp, p_tp)
if rank == 0:
    p_full = torch.empty_like(x0); fwd(m_full, p_full)
    say(what="tp_vs_unsharded_rel_err", rel=float((p_tp - p_full).norm() / p_full.norm()), p2p=tpg.p2p_status())
tm.barrier()
# all ranks must hold identical bits
chk = p_tp.view(torch.int32).to(torch.int64).sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
say(what="ranks_bit_identical", same=bool(lo.item() == hi.item()))

for fused, gated, chunks, reserve in (("1", "2", 1, 0), ("1", "1", 1, 0), ("1", "0", 1, 0), ("0", "0", 1, 0), ("1", "2", 1, 0)):
    if tokens < chunks * 512: continue
    os.environ["DLLM_TP_FUSED_RS"] = fused
    os.environ["DLLM_TP_GATED"] = gated
    ctx._ck(lib.dllm_tp_configure(ctx.h, chunks, reserve if chunks > 1 else 0, 0))
    x.copy_(x0); ms_c = tm.run(step, 4, 2)
    ctx._ck(lib.dllm_tp_configure(ctx.h, chunks, reserve if chunks > 1 else 0, 1))
    ms_n = tm.run(step, 3, 1)
    say(what="tp_step", world=world, p2p=ok, fused_rs=fused, gated=gated, chunks=chunks, reserve=reserve, ms=ms_c, gemm_only_ms=ms_n, exposed_ms=ms_c - ms_n,
        efficiency=ms_single / (world * ms_c), finite=bool(torch.isfinite(x).all()))
ctx._ck(lib.dllm_tp_configure(ctx.h, 0, -1, 0))
say(what="p2p_status_end", status=tpg.p2p_status())
m_tp.close()
tpg.close()
dist.barrier(); dist.destroy_process_group()
