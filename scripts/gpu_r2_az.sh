#!/bin/bash
# round 2, GPU call AZ: seeded loop in the int8 mode, CUDA graph against eager launches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2az.log 2>&1
timeout 60 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "int8_mode_graph" 2>&1 | tail -15
echo "== done"
