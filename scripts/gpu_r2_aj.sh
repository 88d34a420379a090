#!/bin/bash
# round 2, GPU call AJ: bf16 CTA-pair kernel: accumulator half handed back before its last 32 columns are converted (A/B against the
# build that hands it back after the conversion); int8 stack back on the two-batch drain
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2aj.log 2>&1
echo "== pytest gpu (dense kernels)"
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8 or umma or pair" 2>&1 | tail -8
SHORT="python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --no-tp7b --no-kv32k"
for rep in 1 2; do
echo "== headline, early hand-back (main build), run $rep"
timeout 300 $SHORT 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'], d['clocks'])"
echo "== headline, late hand-back, run $rep"
DLLM_B200_LIB=$PWD/diffusion-llm-rs_b200/lib_exp/libdllm_b200_bf16late.so timeout 300 $SHORT 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'], d['clocks'])"
done
echo "== int8 stack"
timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1
echo "== done"
