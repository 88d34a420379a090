#!/bin/bash
# round 2, GPU call AH: int8 CTA-pair kernel with three accumulator buffers + early hand-back in the drain; warp-per-row activation
# quantizer with programmatic dependent launch; against the two-buffer build and a 224-token tile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ah.log 2>&1
echo "== pytest gpu (int8)"
timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8" 2>&1 | tail -25
echo "== int8 stack (3 buffers)"
timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1
echo "== int8 stack (2 buffers, 6 slots)"
DLLM_B200_LIB=$PWD/diffusion-llm-rs_b200/lib_exp/libdllm_b200_i8nbuf2.so timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1
echo "== int8 stack (3 buffers, 224-token tiles)"
DLLM_UMMA_NTOK2=224 timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1
echo "== timeline"
DLLM_UMMA_DBG=128 timeout 120 python scripts/i8_pair_probe.py 2048 2048 8192
mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2ah_trace_2048x2048_bf16out.csv
rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
echo "== done"
