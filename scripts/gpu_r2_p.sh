#!/bin/bash
# round 2, GPU call P: GEMV with the parameter stream under the evict-first policy: graph-replay sweep + GEMV tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2p.log 2>&1
timeout 600 python scripts/gemv_graph_bench.py 14336,8192,4096 4,2,8 1,4,16 > gpurun_out/r2p_gemv.jsonl 2>&1; cat gpurun_out/r2p_gemv.jsonl
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "gemv" 2>&1 | tail -3
echo "== done"
