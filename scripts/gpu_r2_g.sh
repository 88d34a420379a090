#!/bin/bash
# round 2, GPU call G (2 GPUs): multi-GPU check against the oracle, then the bench line at N = 2 as the driver launches it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2g.log 2>&1
nvidia-smi -L
echo "== multi-GPU test"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
echo "== bench N=2"
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2g_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2g_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline", "clocks")}, indent=1)[:2500])
    print(json.dumps(d.get("tp7b"), indent=1)); print(json.dumps(d.get("kv32k"), indent=1)[:3000])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== done"
