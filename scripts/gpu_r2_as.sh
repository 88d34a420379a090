#!/bin/bash
# round 2, GPU call AS: int8 CTA-pair kernel, stage timeline of the empty pipeline (no MMAs / unpack / activation loads / stores)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2as.log 2>&1
for dbg in 203 128; do
  DLLM_UMMA_DBG=$dbg timeout 120 python scripts/i8_pair_probe.py 2048 8192 8192 2>&1 | grep "^int8"
  mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2as_trace_dbg$dbg.csv
  rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
done
echo "== done"
