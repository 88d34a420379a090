#!/bin/bash
# round 2, GPU call BA: int8 CTA-pair kernel, tensor-memory hand-offs arrive on the leader's barriers without the cluster-scope release
# (MEMBAR.ALL.GPU + ERRBAR in front of every such arrive before)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ba.log 2>&1
timeout 40 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8 or int8" 2>&1 | tail -4
timeout 40 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470
echo "== done"
