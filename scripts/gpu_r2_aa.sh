#!/bin/bash
# round 2, GPU call AA (2 GPUs): exchange kernels with weak data accesses: check against the oracle, all-reduce alone, TP steps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2aa.log 2>&1
echo "== multi-GPU test"
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -12
echo "== all-reduce alone"
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 scripts/p2p_blocks_probe.py 2>/dev/null | grep blocks
echo "== tp probe"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 scripts/tp_probe.py 7b > gpurun_out/r2aa_tp_probe.jsonl 2> gpurun_out/r2aa_tp_probe.err; echo "rc=$?"
grep -E "allreduce_64|p2p_vs|tp_step|timed|p2p_status_end|single|rel_err|identical" gpurun_out/r2aa_tp_probe.jsonl | cut -c1-420; tail -c 300 gpurun_out/r2aa_tp_probe.err
echo "== done"
