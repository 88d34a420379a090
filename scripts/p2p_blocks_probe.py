"""Own all-reduce over NVLink peer memory, 64 MiB, as a function of the block count (DLLM_P2P_BLOCKS, read once per process):
how many SMs keep the links busy.  Run under torchrun; rank 0 prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diffusion-llm-rs_b200"))
import torch, torch.distributed as dist
import dllm_b200
from dllm_b200 import parallel as PAR
from bench import Timer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
ctx = dllm_b200.Context(local, stream=stream.cuda_stream)
tm = Timer(torch, stream, world)
tpg = PAR.TensorParallelGroup(ctx, rank, world); tpg.init_nccl()
ok = tpg.enable_p2p(8192, 4096)
st = tpg.p2p_status()
n = 16 << 20
ms = tm.run(lambda i: tpg.allreduce_dev(st["arena"], n), 30, 5)
if rank == 0:
    print(json.dumps({"blocks": os.environ.get("DLLM_P2P_BLOCKS", "default"), "world": world, "ms": ms,
                      "link_GBps_per_direction": n * 4 * 2 * (world - 1) / world / ms / 1e6}), flush=True)
tpg.close(); dist.barrier(); dist.destroy_process_group()
