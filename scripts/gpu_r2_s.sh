#!/bin/bash
# round 2, GPU call S (2 GPUs): reduce / all-gather kernel with bulk-copy peer stores: multi-GPU check, probe (bulk vs plain stores)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2s.log 2>&1
echo "== f32 SIMT path, graph replay (1 GPU)"
DLLM_BENCH_PATH=1 timeout 300 python scripts/gemv_graph_bench.py 8192,4096 4,8 1,4 2>&1 | grep -v "^$"
echo "== multi-GPU test"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30
for bulk in 1 0; do
echo "== tp probe, DLLM_P2P_BULK=$bulk"
DLLM_P2P_BULK=$bulk timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$bulk scripts/tp_probe.py 7b > gpurun_out/r2s_tp_probe_bulk$bulk.jsonl 2> gpurun_out/r2s_tp_probe.err; echo "rc=$?"
grep -E "tp_step|timed|p2p_status_end" gpurun_out/r2s_tp_probe_bulk$bulk.jsonl | cut -c1-420; tail -c 300 gpurun_out/r2s_tp_probe.err
done
echo "== done"
