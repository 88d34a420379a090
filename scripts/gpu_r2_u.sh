#!/bin/bash
# round 2, GPU call U (2 GPUs): all-gather half under the consuming GEMM (gated loads): check, probe — tight time-outs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2u.log 2>&1
echo "== tp probe"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29585 scripts/tp_probe.py 7b > gpurun_out/r2u_tp_probe.jsonl 2> gpurun_out/r2u_tp_probe.err; echo "rc=$?"
grep -E "tp_step|timed|p2p_status_end|single|rel_err|identical" gpurun_out/r2u_tp_probe.jsonl | cut -c1-420; tail -c 300 gpurun_out/r2u_tp_probe.err
echo "== multi-GPU test"
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30
echo "== done"
