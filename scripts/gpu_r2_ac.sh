#!/bin/bash
# round 2, GPU call AC: dense CTA-pair kernel with both accumulator halves drained before the stores: parity, timings, headline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ac.log 2>&1
echo "== dense / pair / model tests"
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -x -q -k "umma or pair or benchmark or stack or sample or dense" 2>&1 | tail -4
export DLLM_PROBE_BF16_OUT=1
for shape in "2048 2048" "2048 8192" "8192 2048" "4096 4096" "4096 14336" "14336 4096"; do timeout 120 python scripts/dense_probe.py $shape 4 8192; done
unset DLLM_PROBE_BF16_OUT
echo "== headline"
timeout 600 python bench.py --no-cpu --no-secondary --no-kv32k | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['tp7b']['single_gpu'])"
echo "== done"
