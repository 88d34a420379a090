#!/bin/bash
# round 2, GPU call AN: stage timeline of the bf16 CTA-pair kernel (final build) on the step's shapes, and its tile widths
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2an.log 2>&1
for shp in "2048 2048" "2048 8192"; do
  DLLM_PROBE_BF16_OUT=1 timeout 120 python scripts/dense_probe.py $shp 4 8192
  for nt in 224 256; do DLLM_UMMA_NTOK2=$nt DLLM_PROBE_BF16_OUT=1 timeout 120 python scripts/dense_probe.py $shp 4 8192; done
  DLLM_UMMA_DBG=128 DLLM_PROBE_BF16_OUT=1 timeout 120 python scripts/dense_probe.py $shp 4 8192
  mv gpurun_out/pair2_trace.csv "gpurun_out/r2an_trace_bf16_${shp// /x}.csv"
done
echo "== done"
