#!/bin/bash
# round 2, GPU call X: final evidence of the current build on one GPU: whole suite, full bench line, reference arm, ncu launch list,
# ncu full pages (dense CTA-pair kernel inside the step; GEMV 4-bit and 2-bit M = 1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2x.log 2>&1
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== bench (full, N=1)"
timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "rc=$?"; tail -c 600 gpurun_out/r2x_bench.err
echo "== reference arm"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2x_ref.json 2>&1; echo "rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --no-tp7b --no-kv32k"
timeout 300 $SHORT > gpurun_out/r2x_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"umma|f32_to_bf16|p_sample|noise|fixup|bf16" -c 500 --csv --log-file gpurun_out/r2x_launches.csv $SHORT > gpurun_out/r2x_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair2 -s 126 -c 6 -o gpurun_out/r2x_pair2 $SHORT > gpurun_out/r2x_ncu2.log 2>&1
echo "pair2 full rc=$?"
for bits in 4 2; do
  timeout 120 python scripts/gemv_ncu.py 14336 14336 $bits 1 > gpurun_out/r2x_gemv_plain_$bits.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_mma -s 3 -c 2 -o gpurun_out/r2x_gemv_b$bits python scripts/gemv_ncu.py 14336 14336 $bits 1 > gpurun_out/r2x_ncu_gemv_$bits.log 2>&1
  echo "gemv $bits-bit full rc=$?"
done
ls -la gpurun_out | tail -20
echo "== done"
