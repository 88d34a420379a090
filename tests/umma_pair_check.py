"""Parity of the CTA-pair (cta_group::2) variants of the tcgen05 kernel on a dense shape (DLLM_UMMA_PAIR=1: 128-token pair
tiles; =2: the 256-token pair kernel of the denoise step, DLLM_UMMA_NTOK2 forces its tile width): the output must equal
the 1-CTA kernel's (DLLM_UMMA_PAIR=0) bit for bit (same k order, same accumulation) and match the f64 reference.
usage: umma_pair_check.py out.npy [K N M [group]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-llm-rs_b200"))
import numpy as np
import torch
import dllm_b200
from dllm_b200 import QWeight, PATH_UMMA

ctx = dllm_b200.Context(0)
g = torch.Generator(device="cuda").manual_seed(7)
K, N, M = 384, 7000, 1300            # 55 column tiles (odd: the last pair is half empty; ragged N), 11 token tiles (ragged): dense mode
GROUP = 128
if len(sys.argv) > 4:
    K, N, M = (int(v) for v in sys.argv[2:5])
if len(sys.argv) > 5:
    GROUP = int(sys.argv[5])
w = torch.randn(K, N, device="cuda", generator=g) * 0.02
x = torch.randn(M, K, device="cuda", generator=g)
y = torch.empty(M, N, device="cuda")
torch.cuda.synchronize()
qw = QWeight.quantize_dev(ctx, w.data_ptr(), K, N, 4, GROUP)
qw.forward_dev(x.data_ptr(), M, y.data_ptr(), PATH_UMMA)
ctx.sync()
codes, scales, zps = qw.export()
wd = (torch.from_numpy(codes.astype(np.float64)) - torch.from_numpy(np.repeat(zps, GROUP, axis=0).astype(np.float64))) * \
    torch.from_numpy(np.repeat(scales, GROUP, axis=0).astype(np.float64))
ref = x.cpu().double() @ wd
err = (y.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
assert err <= 1e-2, err
out = sys.argv[1] if len(sys.argv) > 1 else None
if out:
    np.save(out, y.cpu().numpy())
print(f"PAIR_CHECK_OK pair={os.environ.get('DLLM_UMMA_PAIR', '0')} rel_err={err:.2e}")
