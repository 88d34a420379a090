#!/bin/bash
# round 2, GPU call I: GEMV with aligned stages + int32-chained pairs: parity tests, then KBS 2 vs 4 under graph replay
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2i.log 2>&1
echo "== GEMV tests"
timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "gemv or auto_path or paths" 2>&1 | tail -8
echo "== GEMV tests with KBS=4 forced"
DLLM_GEMV_KBS=4 timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "gemv" 2>&1 | tail -5
echo "== GEMV tests with KBS=2 forced"
DLLM_GEMV_KBS=2 timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "gemv" 2>&1 | tail -5
echo "== graph-replay sweep, KBS=2"
DLLM_GEMV_KBS=2 timeout 600 python scripts/gemv_graph_bench.py 14336,8192,4096 4,2,8 1,4,16 > gpurun_out/r2i_gemv_kbs2.jsonl 2>&1; cat gpurun_out/r2i_gemv_kbs2.jsonl
echo "== graph-replay sweep, KBS=4"
DLLM_GEMV_KBS=4 timeout 600 python scripts/gemv_graph_bench.py 14336,8192,4096 4,2,8 1,4,16 > gpurun_out/r2i_gemv_kbs4.jsonl 2>&1; cat gpurun_out/r2i_gemv_kbs4.jsonl
echo "== model tests (GEMV path inside the stack), dense tests"
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_linear.py -m gpu -x -q 2>&1 | tail -5
echo "== done"
