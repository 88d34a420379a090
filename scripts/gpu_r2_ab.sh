#!/bin/bash
# round 2, GPU call AB: per-stage / per-tile clock stamps of the dense CTA-pair kernel on the [2048,2048] x 8192 linear (tile-boundary gaps)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ab.log 2>&1
export DLLM_PROBE_BF16_OUT=1
DLLM_UMMA_DBG=128 timeout 120 python scripts/dense_probe.py 2048 2048 4 8192
cp gpurun_out/pair2_trace.csv gpurun_out/r2ab_pair2_trace_2048.csv
DLLM_UMMA_DBG=128 timeout 120 python scripts/dense_probe.py 8192 2048 4 8192
cp gpurun_out/pair2_trace.csv gpurun_out/r2ab_pair2_trace_8192.csv
timeout 120 python scripts/dense_probe.py 2048 2048 4 8192
timeout 120 python scripts/dense_probe.py 8192 2048 4 8192
echo "== done"
