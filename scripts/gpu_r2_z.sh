#!/bin/bash
# round 2, GPU call Z: compute-sanitizer memcheck over the code paths that are new this round (GEMV aligned stages / pair chaining,
# HBM-resident KV cache entry, seeded loop, dense CTA-pair kernel on a small dense shape)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2z.log 2>&1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "kv_cache or cached or seeded or add_noise" 2>&1 | tail -25
echo "rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "stage_shapes or gemv_shapes or pair" 2>&1 | tail -25
echo "rc=$?"
echo "== done"
