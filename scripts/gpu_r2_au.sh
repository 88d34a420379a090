#!/bin/bash
# round 2, GPU call AU: final int8 kernel build (8 epilogue warps, generalized epilogue indexing, timing switches): parity + stack
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2au.log 2>&1
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py tests/test_gpu_quantizers.py -m gpu -q -x -k "i8 or umma or pair or adaptive" 2>&1 | tail -4
timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470
echo "== done"
