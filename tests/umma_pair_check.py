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
# the same shape as the first layer of a two-layer stack [K -> N -> K]: the intermediate activation leaves the first kernel as
# bf16 through the staged bulk-tensor-store epilogue (ragged tokens and columns are clipped by the TMA unit) and is the
# second kernel's input; the second kernel writes f32
from dllm_b200.diffuse_llm import QuantizedDiffusionModel
w2 = torch.randn(N, K, device="cuda", generator=g) * (1.0 / N ** 0.5)
torch.cuda.synchronize()
qw2 = QWeight.quantize_dev(ctx, w2.data_ptr(), N, K, 4, GROUP if N % GROUP == 0 else 0)
model = QuantizedDiffusionModel([qw, qw2], K, ctx=ctx, path=PATH_UMMA)
y2 = torch.empty(M, K, device="cuda")
torch.cuda.synchronize()
model.forward_dev(x.data_ptr(), M, K, y2.data_ptr())
ctx.sync()
c2, s2, z2 = qw2.export()
if s2.shape[0] == 1 and s2.shape[1] == 1:
    wd2 = (torch.from_numpy(c2.astype(np.float64)) - float(z2[0, 0])) * float(s2[0, 0])
else:
    wd2 = (torch.from_numpy(c2.astype(np.float64)) - torch.from_numpy(np.repeat(z2, GROUP, axis=0).astype(np.float64))) * \
        torch.from_numpy(np.repeat(s2, GROUP, axis=0).astype(np.float64))
ref2 = ref @ wd2
err2 = (y2.cpu().double() - ref2).abs().max().item() / ref2.abs().max().item()
assert err2 <= 2e-2, err2
out = sys.argv[1] if len(sys.argv) > 1 else None
if out:
    np.save(out, np.concatenate([y.cpu().numpy().ravel(), y2.cpu().numpy().ravel()]))
print(f"PAIR_CHECK_OK pair={os.environ.get('DLLM_UMMA_PAIR', '0')} rel_err={err:.2e} stack_rel_err={err2:.2e}")
