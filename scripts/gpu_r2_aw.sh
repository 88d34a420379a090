#!/bin/bash
# round 2, GPU call AW: int8 CTA-pair kernel with seven A / X slots (no parameter loads in the weight ring)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2aw.log 2>&1
echo "== pytest gpu (int8)"
timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8" 2>&1 | tail -4
echo "== int8 stack (7 slots, half 0 two stages ahead)"
timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470
echo "== int8 stack (7 slots, three stages ahead)"
DLLM_B200_LIB=$PWD/diffusion-llm-rs_b200/lib_exp/libdllm_b200_i8pre3.so timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470
echo "== timeline"
DLLM_UMMA_DBG=128 timeout 120 python scripts/i8_pair_probe.py 2048 8192 8192 | grep "^int8"
mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2aw_trace.csv
rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
echo "== done"
