#!/bin/bash
# round 2, GPU call AX: final int8 kernel configuration (7 slots, half 0 three stages ahead): parity of everything that touches the dense
# kernels + the int8 stack
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2ax.log 2>&1
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_model.py -m gpu -q -x -k "i8 or umma or pair or host" 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --only-int8-stack 2>&1 | tail -1 | cut -c290-470
echo "== done"
