#!/bin/bash
# round 2, GPU call AL: int8 CTA-pair kernel with two unpack groups (640 threads, 96 registers), with and without the one-round-trip drain
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2al.log 2>&1
L=$PWD/diffusion-llm-rs_b200/lib_exp
for v in main i8ndq2 i8ndq2d64; do
  echo "== int8 stack: $v"
  if [ $v = main ]; then timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470
  else DLLM_B200_LIB=$L/libdllm_b200_$v.so timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470; fi
done
for v in i8ndq2 i8ndq2d64; do
echo "== timeline ($v)"
DLLM_B200_LIB=$L/libdllm_b200_$v.so DLLM_UMMA_DBG=128 timeout 120 python scripts/i8_pair_probe.py 2048 2048 8192
mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2al_trace_${v}.csv
rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
done
echo "== done"
