#!/bin/bash
# round 2, GPU call M (N GPUs, N = $1): the bench line exactly as the driver launches it
N=${1:-4}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2m_n$N.log 2>&1
nvidia-smi -L | wc -l
echo "== bench N=$N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2m_bench_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2m_bench_n{N}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "e2e")}, indent=1)[:1200])
    print(json.dumps(d.get("tp7b"), indent=1))
    kv = d.get("kv32k") or {}
    for k, v in kv.items():
        if isinstance(v, dict): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
except Exception as e:
    print("bench parse failed", e)
PY
echo "== done"
