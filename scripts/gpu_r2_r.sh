#!/bin/bash
# round 2, GPU call R: SIMT path per call (small shapes), then the whole GPU suite on the current build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2r.log 2>&1
echo "== f32 SIMT path, graph replay"
DLLM_BENCH_PATH=1 timeout 300 python scripts/gemv_graph_bench.py 8192,4096 4,8 1,4 2>&1 | grep -v "^$"
echo "== pytest gpu (all)"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== done"
