#!/bin/bash
# round 2, GPU call AY: the whole GPU suite on the final commit
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ay.log 2>&1
timeout 170 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== done"
