#!/bin/bash
# round 2, GPU call L (2 GPUs): fused GEMM + reduce-scatter: multi-GPU check, TP probe, model tests, headline regression check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2l.log 2>&1
echo "== model tests"
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -8
echo "== multi-GPU test"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30
echo "== tp probe"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/tp_probe.py 7b > gpurun_out/r2l_tp_probe.jsonl 2> gpurun_out/r2l_tp_probe.err; echo "rc=$?"
grep -v "^$" gpurun_out/r2l_tp_probe.jsonl | cut -c1-420; tail -c 800 gpurun_out/r2l_tp_probe.err
echo "== headline (1 GPU, short)"
timeout 600 python bench.py --no-cpu --no-secondary --no-tp7b --no-kv32k | cut -c1-400
echo "== done"
