#!/bin/bash
# round 2, GPU call Q: GEMV floor: activations prepared by a separate kernel and riding the ring (no preparation code in the
# main kernel) against the resident form; small weight pools (L2-resident) against pools larger than L2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2q.log 2>&1
echo "== resident activations (default)"
timeout 300 python scripts/gemv_graph_bench.py 14336,8192,4096 4 1 2>&1 | grep -v "^$"
echo "== activations prepared by their own kernel, riding the ring"
DLLM_GEMV_XR=0 timeout 300 python scripts/gemv_graph_bench.py 14336,8192,4096 4 1 2>&1 | grep -v "^$"
echo "== default, pool of 2 weights (8192: 67 MB, 4096: 17 MB: L2-resident)"
DLLM_GEMV_POOL=2 timeout 300 python scripts/gemv_graph_bench.py 8192,4096 4 1 2>&1 | grep -v "^$"
echo "== done"
