#!/bin/bash
# round 2, GPU call T (4 GPUs): reduce / all-gather with bulk-copy peer stores against plain stores at 4 ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2t.log 2>&1
for bulk in 1 0; do
echo "== tp probe, 4 ranks, DLLM_P2P_BULK=$bulk"
DLLM_P2P_BULK=$bulk timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2957$bulk scripts/tp_probe.py 7b > gpurun_out/r2t_tp_probe_bulk$bulk.jsonl 2> gpurun_out/r2t_tp_probe.err; echo "rc=$?"
grep -E "allreduce_64|tp_step|timed|p2p_status_end|single" gpurun_out/r2t_tp_probe_bulk$bulk.jsonl | cut -c1-420; tail -c 300 gpurun_out/r2t_tp_probe.err
done
echo "== done"
