#!/bin/bash
# round 2, GPU call N: GEMV per-call floor: how many ring stages may be requested before the activations are resident
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2n.log 2>&1
for pf in default 0 3 6 12; do
  echo "== prefill $pf"
  if [ $pf = default ]; then unset DLLM_GEMV_PREFILL; else export DLLM_GEMV_PREFILL=$pf; fi
  timeout 300 python scripts/gemv_graph_bench.py 14336,8192,4096 4,2 1 2>&1 | grep -v "^$"
done
echo "== timelines (trace build), 4-bit 14336 M=1"
for pf in default 0 3; do
  if [ $pf = default ]; then unset DLLM_GEMV_PREFILL; else export DLLM_GEMV_PREFILL=$pf; fi
  DLLM_B200_LIB=$PWD/diffusion-llm-rs_b200/lib_exp/libdllm_b200_trace.so timeout 300 python scripts/gemv_probe.py 14336 14336 4 1 2>&1 | tail -2
  cp gpurun_out/gemv_trace.csv gpurun_out/r2n_gemv_trace_pf$pf.csv
done
echo "== done"
