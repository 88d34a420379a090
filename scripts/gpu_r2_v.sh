#!/bin/bash
# round 2, GPU call V (4 GPUs): probe with token-tile rotation + gated all-gather
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2v.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29593 scripts/tp_probe.py 7b > gpurun_out/r2v_tp_probe.jsonl 2> gpurun_out/r2v_tp_probe.err; echo "rc=$?"
grep -E "tp_step|timed|p2p_status_end|single|rel_err|identical" gpurun_out/r2v_tp_probe.jsonl | cut -c1-420; tail -c 400 gpurun_out/r2v_tp_probe.err
echo "== done"
