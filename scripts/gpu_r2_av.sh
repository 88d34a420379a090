#!/bin/bash
# round 2, GPU call AV: C++ host mirror harness with the int8-mode check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2av.log 2>&1
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "host" 2>&1 | tail -6
echo "== done"
