#!/bin/bash
# round 2, GPU call W (2 GPUs): how many blocks the own all-reduce needs to keep NVLink busy
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2w.log 2>&1
for b in 8 16 32 64 128 256; do
DLLM_P2P_BLOCKS=$b timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 296$b scripts/p2p_blocks_probe.py 2>/dev/null | grep blocks
done
echo "== done"
