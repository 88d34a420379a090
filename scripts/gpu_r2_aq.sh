#!/bin/bash
# round 2, GPU call AQ: AdaptiveQuantizer against the sketch restatement; bench line with the K = 4096 GEMV cases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2aq.log 2>&1
timeout 300 python -m pytest tests/test_gpu_quantizers.py -m gpu -q -x -k "adaptive" 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-tp7b --no-kv32k > gpurun_out/r2aq_bench.json 2> gpurun_out/r2aq_bench.err; echo "rc=$?"; tail -c 600 gpurun_out/r2aq_bench.err
echo "== done"
