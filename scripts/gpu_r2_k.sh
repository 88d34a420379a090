#!/bin/bash
# round 2, GPU call K (2 GPUs): device KV cache tests (1 GPU), multi-GPU check incl. the peer-to-peer all-reduce, bench N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2k.log 2>&1
echo "== model tests"
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15
echo "== multi-GPU test"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -25
echo "== bench N=2"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-kv32k > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo "rc=$?"; tail -c 1200 gpurun_out/r2k_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2k_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, indent=1))
    print(json.dumps(d.get("tp7b"), indent=1))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== done"
