#!/bin/bash
# round 2, GPU call BB: bf16 CTA-pair kernel with the same plain arrives for its tensor-memory hand-offs: bit-identity against the
# 1-CTA kernel (whose barriers are CTA-local) on the dense ragged shape, then the headline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
exec > gpurun_out/r2bb.log 2>&1
DLLM_UMMA_PAIR=0 timeout 40 python tests/umma_pair_check.py /tmp/a.npy > /tmp/a.log 2>&1 &
DLLM_UMMA_PAIR=2 timeout 40 python tests/umma_pair_check.py /tmp/b.npy > /tmp/b.log 2>&1 &
wait
tail -1 /tmp/a.log; tail -1 /tmp/b.log
cmp /tmp/a.npy /tmp/b.npy && echo "BIT_IDENTICAL"
timeout 40 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary --no-tp7b --no-kv32k 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'finite', d['output_finite'], 'e2e', d['e2e']['value'])"
echo "== done"
