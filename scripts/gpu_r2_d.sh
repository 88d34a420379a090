#!/bin/bash
# round 2, GPU call D: bf16-output probes of the pair2 kernel, new bench.py (N=1), pair / model tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2d.log 2>&1
echo "== pair checks (single linear f32 out + two-layer stack bf16 intermediate)"
timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "pair" 2>&1 | tail -5
echo "== dense probe, bf16 output (as inside the stack)"
export DLLM_PROBE_BF16_OUT=1
for shape in "2048 2048" "2048 8192" "8192 2048"; do
  DLLM_UMMA_PAIR=0 timeout 120 python scripts/dense_probe.py $shape 4 8192
  DLLM_UMMA_PAIR=2 timeout 120 python scripts/dense_probe.py $shape 4 8192
  DLLM_UMMA_PAIR=2 DLLM_UMMA_DBG=512 timeout 120 python scripts/dense_probe.py $shape 4 8192
done
for nt in 192 224 256; do DLLM_UMMA_PAIR=2 DLLM_UMMA_NTOK2=$nt timeout 120 python scripts/dense_probe.py 2048 2048 4 8192; done
echo "== pair2 instruction-removal runs (2048x8192, bf16 out)"
for dbg in 1 2 8 64 3 75; do DLLM_UMMA_PAIR=2 DLLM_UMMA_DBG=$dbg timeout 120 python scripts/dense_probe.py 2048 8192 4 8192; done
echo "== timeline (bf16 out)"
DLLM_UMMA_PAIR=2 DLLM_UMMA_DBG=128 timeout 120 python scripts/dense_probe.py 2048 8192 4 8192
unset DLLM_PROBE_BF16_OUT
echo "== bench (full, N=1)"
timeout 1200 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2d_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline", "clocks")}, indent=1))
    print(json.dumps(d.get("tp7b"), indent=1)); print(json.dumps(d.get("kv32k"), indent=1)); print(json.dumps(d.get("gemv"), indent=1))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-1200
echo "== done"
