#!/bin/bash
# round 2, GPU call Y: GEMV: shifts on the FMA pipe (IMAD.HI) + address translations prefetched before the dependency wait
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2y.log 2>&1
echo "== both changes"
timeout 400 python scripts/gemv_graph_bench.py 14336,8192,4096 4,2,8 1,16 2>&1 | grep -v "^$"
echo "== shifts on the FMA pipe only (no prefetch before the wait)"
DLLM_B200_LIB=$PWD/diffusion-llm-rs_b200/lib_exp/libdllm_b200_nopf.so timeout 400 python scripts/gemv_graph_bench.py 14336,8192,4096 4,2 1 2>&1 | grep -v "^$"
echo "== tests"
timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "gemv" 2>&1 | tail -3
echo "== done"
