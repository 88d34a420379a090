#!/bin/bash
# round 2, final evidence of the build on one GPU: whole suite, smoke, full bench line, reference arm, ncu launch list of the step,
# ncu full pages (dense bf16 CTA-pair kernel inside the step; int8 CTA-pair kernel inside the int8 stack; GEMV 4-bit and 2-bit M = 1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2fin.log 2>&1
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (full, N=1)"
timeout 900 python bench.py > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err; echo "rc=$?"; tail -c 600 gpurun_out/r2fin_bench.err
echo "== reference arm"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2fin_ref.json 2>&1; echo "rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --no-tp7b --no-kv32k"
timeout 300 $SHORT > gpurun_out/r2fin_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"umma|f32_to_bf16|p_sample|noise|fixup|bf16" -c 500 --csv --log-file gpurun_out/r2fin_launches.csv $SHORT > gpurun_out/r2fin_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair2 -s 126 -c 6 -o gpurun_out/r2fin_pair2 $SHORT > gpurun_out/r2fin_ncu2.log 2>&1
echo "pair2 full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair2 -s 12 -c 6 -o gpurun_out/r2fin_pair2_i8 python bench.py --only-int8-stack > gpurun_out/r2fin_ncu3.log 2>&1
echo "pair2 int8 full rc=$?"
ls -la gpurun_out | tail -12
echo "== done"
