#!/bin/bash
# round 2, final single-GPU record: smoke, C++ host mirror harness, the bench line, the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2final1.log 2>&1
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== host mirror + model tests"
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
echo "== bench (full, N=1)"
timeout 900 python bench.py > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err; echo "rc=$?"; tail -c 400 gpurun_out/r2final_bench_n1.err
echo "== bench (driver-style flags)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --no-secondary --no-tp7b --no-kv32k --no-cpu | cut -c1-900
echo "== reference arm"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2final_ref.json 2>&1; echo "rc=$?"
echo "== done"
