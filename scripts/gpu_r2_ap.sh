#!/bin/bash
# round 2, GPU call AP: int8 CTA-pair kernel, both k-blocks of a stage unpacked in one straight-line block: A ring in shared memory
# (4 accumulator buffers) against A ring in tensor memory (2 buffers, 6 slots, half 0 two stages ahead)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ap.log 2>&1
L=$PWD/diffusion-llm-rs_b200/lib_exp
echo "== pytest gpu (int8), main"
timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8" 2>&1 | tail -3
echo "== pytest gpu (int8), tensor-memory A ring"
DLLM_B200_LIB=$L/libdllm_b200_i8tmem.so timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8" 2>&1 | tail -3
for v in main i8tmem; do
  echo "== int8 stack: $v"
  if [ $v = main ]; then timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470
  else DLLM_B200_LIB=$L/libdllm_b200_$v.so timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470; fi
done
echo "== timeline (main)"
DLLM_UMMA_DBG=128 timeout 120 python scripts/i8_pair_probe.py 2048 2048 8192
mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2ap_trace_main.csv
rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
echo "== done"
