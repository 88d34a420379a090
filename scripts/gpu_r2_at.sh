#!/bin/bash
# round 2, GPU call AT: int8 CTA-pair kernel with 16 epilogue warps (1024 threads, 64 registers)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2at.log 2>&1
echo "== pytest gpu (int8)"
timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8" 2>&1 | tail -12
echo "== int8 stack"
timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470
echo "== timeline"
DLLM_UMMA_DBG=128 timeout 120 python scripts/i8_pair_probe.py 2048 2048 8192
mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2at_trace.csv
rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
echo "== done"
