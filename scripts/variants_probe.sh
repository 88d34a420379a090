#!/bin/bash
# time the GEMV headline shape with the normal build and every experiment build present in lib_exp/
cd "$(dirname "$0")/.."
for lib in "" diffusion-llm-rs_b200/lib_exp/*.so; do
  echo "== ${lib:-default}"
  for a in "$@"; do DLLM_B200_LIB=${lib:+$PWD/$lib} timeout 120 python scripts/gemv_probe.py $a 2>&1 | tail -2; done
done
