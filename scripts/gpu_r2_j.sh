#!/bin/bash
# round 2, GPU call J (2 GPUs): peer-to-peer all-reduce probe (correctness, timing, 7B-class TP step in every mode)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2j.log 2>&1
nvidia-smi topo -m 2>&1 | head -8
echo "== tp probe"
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/tp_probe.py 7b > gpurun_out/r2j_tp_probe.jsonl 2> gpurun_out/r2j_tp_probe.err; echo "rc=$?"
cat gpurun_out/r2j_tp_probe.jsonl; tail -c 1500 gpurun_out/r2j_tp_probe.err
echo "== done"
