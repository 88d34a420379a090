#!/bin/bash
# round 2, GPU call AF: int8 variant of the 256-token CTA-pair kernel (kind::i8 cta_group::2) + single-pass activation quantizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2af.log 2>&1
echo "== pytest gpu (int8)"
timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8" 2>&1 | tail -25
rc=${PIPESTATUS[0]}
echo "rc=$rc"
if [ "$rc" != "0" ]; then
  echo "== the same with the 1-CTA int8 kernel (DLLM_I8_PAIR=0)"
  DLLM_I8_PAIR=0 timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -k "i8" 2>&1 | tail -15
fi
echo "== int8 stack"
timeout 300 python bench.py --only-int8-stack 2>&1 | tail -3
echo "== int8 stack, 1-CTA int8 kernel"
DLLM_I8_PAIR=0 timeout 300 python bench.py --only-int8-stack 2>&1 | tail -3
echo "== done"
