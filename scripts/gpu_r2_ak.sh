#!/bin/bash
# round 2, GPU call AK: int8 CTA-pair kernel: half 0 issued 1 / 2 / 3 stages ahead of half 1; early against late hand-back of the buffers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ak.log 2>&1
echo "== pytest gpu (int8)"
timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8" 2>&1 | tail -5
L=$PWD/diffusion-llm-rs_b200/lib_exp
for v in main i8pre1 i8pre3 i8late; do
  echo "== int8 stack: $v"
  if [ $v = main ]; then timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c330-560
  else DLLM_B200_LIB=$L/libdllm_b200_$v.so timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c330-560; fi
done
echo "== timeline (main)"
DLLM_UMMA_DBG=128 timeout 120 python scripts/i8_pair_probe.py 2048 2048 8192
mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2ak_trace_2048x2048_bf16out.csv
rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
echo "== done"
