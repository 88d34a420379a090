#!/bin/bash
# round 2, GPU call H: tile-width sweep of the dense kernel (calibrates the width picker); GEMV stage timelines (2- and 4-bit)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2h.log 2>&1
echo "== ntok sweep"
timeout 900 python scripts/ntok_sweep.py > gpurun_out/r2h_ntok_sweep.jsonl 2> gpurun_out/r2h_ntok_sweep.err; echo "rc=$?"; tail -3 gpurun_out/r2h_ntok_sweep.err
cat gpurun_out/r2h_ntok_sweep.jsonl
echo "== GEMV timelines (trace build)"
for bits in 2 4 8; do
  DLLM_B200_LIB=$PWD/diffusion-llm-rs_b200/lib_exp/libdllm_b200_trace.so timeout 300 python scripts/gemv_probe.py 14336 14336 $bits 1 2>&1 | tail -3
  cp gpurun_out/gemv_trace.csv gpurun_out/r2h_gemv_trace_b$bits.csv; cp gpurun_out/gemv_stages.csv gpurun_out/r2h_gemv_stages_b$bits.csv
done
echo "== done"
