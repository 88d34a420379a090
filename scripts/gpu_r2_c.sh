#!/bin/bash
# round 2, GPU call C: pair2 with the staged bulk-tensor-store epilogue; seeded loop tests; whole GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2c.log 2>&1
echo "== pair2 quick check"
DLLM_UMMA_PAIR=0 timeout 300 python tests/umma_pair_check.py gpurun_out/y0.npy
DLLM_UMMA_PAIR=2 timeout 300 python tests/umma_pair_check.py gpurun_out/y2.npy
python - <<'PY'
import numpy as np
a, b = np.load("gpurun_out/y0.npy"), np.load("gpurun_out/y2.npy")
print("PAIR2 bit-identical:", np.array_equal(a.view(np.uint32), b.view(np.uint32)), "max abs diff", float(np.abs(a - b).max()))
PY
rm -f gpurun_out/y0.npy gpurun_out/y2.npy
echo "== dense probe"
for shape in "2048 2048" "2048 8192" "8192 2048"; do
  DLLM_UMMA_PAIR=0 timeout 120 python scripts/dense_probe.py $shape 4 8192
  DLLM_UMMA_PAIR=2 timeout 120 python scripts/dense_probe.py $shape 4 8192
  for nt in 192 224 256; do DLLM_UMMA_PAIR=2 DLLM_UMMA_NTOK2=$nt timeout 120 python scripts/dense_probe.py $shape 4 8192; done
done
echo "== pair2 instruction-removal runs (2048x8192)"
for dbg in 1 2 8 64 3 75; do DLLM_UMMA_PAIR=2 DLLM_UMMA_DBG=$dbg timeout 120 python scripts/dense_probe.py 2048 8192 4 8192; done
echo "== timeline"
DLLM_UMMA_PAIR=2 DLLM_UMMA_DBG=128 timeout 120 python scripts/dense_probe.py 2048 8192 4 8192
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2c_bench_pair2.json 2> gpurun_out/r2c_bench_pair2.err; tail -c 900 gpurun_out/r2c_bench_pair2.json
echo "== done"
