#!/bin/bash
# round 2, GPU call AR: int8 CTA-pair kernel, instruction-removal runs (which resource bounds the stage period)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ar.log 2>&1
for dbg in 0 1 2 8 64 10 74 75; do
  echo "== DLLM_UMMA_DBG=$dbg"
  DLLM_UMMA_DBG=$dbg timeout 120 python scripts/i8_pair_probe.py 2048 8192 8192 2>&1 | grep "^int8"
done
echo "== done"
