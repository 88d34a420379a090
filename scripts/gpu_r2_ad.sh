#!/bin/bash
# round 2, GPU call AD: int8 denoise mode (DLLM_PATH_I8): new parity tests, then the bench line with the int8_stack block
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ad.log 2>&1
echo "== pytest gpu (int8 mode + neighbours)"
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -k "i8 or umma" 2>&1 | tail -15
echo "== bench (N=1, secondary on)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-tp7b --no-kv32k > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2ad_bench.err
echo "== done"
