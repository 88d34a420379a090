#!/bin/bash
# round 2, GPU call O: where the time between the dependency wait and the activation loads goes (finer stamps)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2o.log 2>&1
DLLM_B200_LIB=$PWD/diffusion-llm-rs_b200/lib_exp/libdllm_b200_trace.so timeout 300 python scripts/gemv_probe.py 14336 14336 4 1 2>&1 | tail -2
cp gpurun_out/gemv_trace.csv gpurun_out/r2o_gemv_trace.csv
DLLM_B200_LIB=$PWD/diffusion-llm-rs_b200/lib_exp/libdllm_b200_trace.so timeout 300 python scripts/gemv_probe.py 4096 4096 4 1 2>&1 | tail -2
cp gpurun_out/gemv_trace.csv gpurun_out/r2o_gemv_trace_4096.csv
echo "== done"
