#!/bin/bash
# round 2, GPU call AE (2 GPUs): the tensor-parallel check against the oracle and the N = 2 bench line on the final build
# (the dense kernel's epilogue — shared with the fused reduce-scatter — changed after the last 2-GPU run)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ae.log 2>&1
echo "== multi-GPU check"
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15
echo "== bench N=2"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ae_bench_n2.json 2> gpurun_out/r2ae_bench_n2.err; echo "rc=$?"; tail -c 800 gpurun_out/r2ae_bench_n2.err
echo "== done"
