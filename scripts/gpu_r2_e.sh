#!/bin/bash
# round 2, GPU call E: pair2 with the faster drain: tests, bf16 probes, timeline, container + seeded tests, short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2e.log 2>&1
echo "== tests: pair, container, noise, seeded, add_noise"
timeout 1200 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -x -q -k "pair or container or noise or seeded or benchmark_shapes" 2>&1 | tail -5
echo "== dense probe, bf16 output"
export DLLM_PROBE_BF16_OUT=1
for shape in "2048 2048" "2048 8192" "8192 2048" "4096 4096" "4096 14336" "14336 4096"; do
  DLLM_UMMA_PAIR=2 timeout 120 python scripts/dense_probe.py $shape 4 8192
done
for dbg in 1 64 3 75; do DLLM_UMMA_PAIR=2 DLLM_UMMA_DBG=$dbg timeout 120 python scripts/dense_probe.py 2048 8192 4 8192; done
DLLM_UMMA_PAIR=2 DLLM_UMMA_DBG=128 timeout 120 python scripts/dense_probe.py 2048 8192 4 8192
unset DLLM_PROBE_BF16_OUT
echo "== bench (short)"
timeout 900 python bench.py --no-cpu --no-secondary --no-kv32k > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "rc=$?"; tail -c 800 gpurun_out/r2e_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline")}, indent=1)[:2500])
    print(json.dumps(d.get("tp7b"), indent=1)[:1500])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== done"
