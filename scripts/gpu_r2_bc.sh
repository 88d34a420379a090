#!/bin/bash
# round 2, GPU call BC (the last seconds of the budget): the benchmark-shape parity test of the bf16 CTA-pair kernel against the f64 oracle
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2bc.log 2>&1
timeout 16 python -m pytest tests/test_gpu_linear.py -m gpu -q -x -k "umma_benchmark_shapes" 2>&1 | tail -4
echo "== done"
