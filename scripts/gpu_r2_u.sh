#!/bin/bash
# round 2, GPU call U (2 GPUs): all-gather half under the consuming GEMM (gated loads) + token-tile rotation: check, probe, headline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2u.log 2>&1
echo "== multi-GPU test"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30
echo "== tp probe"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 scripts/tp_probe.py 7b > gpurun_out/r2u_tp_probe.jsonl 2> gpurun_out/r2u_tp_probe.err; echo "rc=$?"
grep -E "tp_step|timed|p2p_status_end|single|rel_err|identical" gpurun_out/r2u_tp_probe.jsonl | cut -c1-420; tail -c 600 gpurun_out/r2u_tp_probe.err
echo "== headline (1 GPU, short) + dense tests"
timeout 600 python bench.py --no-cpu --no-secondary --no-tp7b --no-kv32k | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -x -q -k "umma or pair or benchmark or stack or sample" 2>&1 | tail -3
echo "== done"
