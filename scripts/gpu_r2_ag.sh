#!/bin/bash
# round 2, GPU call AG: where the int8 CTA-pair kernel spends its time: stage timelines + ncu launch list of the int8 stack
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ag.log 2>&1
for shp in "2048 2048" "2048 8192"; do
  timeout 120 python scripts/i8_pair_probe.py $shp 8192
done
DLLM_UMMA_DBG=128 timeout 120 python scripts/i8_pair_probe.py 2048 2048 8192
mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2ag_trace_2048x2048_bf16out.csv; mv gpurun_out/pair2_i8_trace_1.csv gpurun_out/r2ag_trace_2048x2048_f32out.csv
rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
DLLM_UMMA_DBG=128 timeout 120 python scripts/i8_pair_probe.py 2048 8192 8192
mv gpurun_out/pair2_i8_trace_0.csv gpurun_out/r2ag_trace_2048x8192_bf16out.csv; mv gpurun_out/pair2_i8_trace_1.csv gpurun_out/r2ag_trace_8192x2048_f32out.csv
rm -f gpurun_out/pair2_i8_trace_*.csv gpurun_out/pair2_trace.csv
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pair2|rowquant|f32_to_bf16|p_sample" -c 700 --csv --log-file gpurun_out/r2ag_launches.csv python bench.py --only-int8-stack > gpurun_out/r2ag_ncu.log 2>&1
echo "rc=$?"
echo "== done"
