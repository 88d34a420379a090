#!/bin/bash
# round 2, GPU call F: whole GPU suite, full bench line (N=1), reference arm, ncu launch list + full page of the dense kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2f.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== pytest gpu (all)"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== bench (full, N=1)"
timeout 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2f_bench.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_ref.json 2>&1; echo "rc=$?"; cut -c1-600 gpurun_out/r2f_ref.json
echo "== short bench, plain then under ncu (launch list)"
SHORT="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --no-tp7b --no-kv32k"
$SHORT > gpurun_out/r2f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"umma|f32_to_bf16|p_sample|noise|fixup|bf16" -c 500 --csv --log-file gpurun_out/r2f_launches.csv $SHORT > gpurun_out/r2f_ncu1.log 2>&1
echo "launch list rc=$?"
echo "== ncu full page of the dense kernel (six linears of one layer, second step)"
ncu --set full --clock-control none --import-source on -k regex:pair2 -s 126 -c 6 -o gpurun_out/r2f_pair2 $SHORT > gpurun_out/r2f_ncu2.log 2>&1
echo "full rc=$?"
ls -la gpurun_out
echo "== done"
