"""One plain launch of every kernel that is > 5 % of the training step, at its CIFAR B=128 shape, for
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 python tools/profile_kernels.py
(the summaries under profiles/r2_ncu_*.txt are extracted from those captures with tools/ncu_summary.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticdeepdiffusionmodels_b200 import functional as F  # noqa: E402

bf16, f32 = torch.bfloat16, torch.float32
dev = torch.device("cuda")
B = 128
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, dtype=bf16):
    return torch.randn(shape, generator=g, device=dev).to(dtype)


T3, T1 = F.taps_3x3(), F.taps_1x1()
# conv_fwd_kernel: 3x3 16x16 256->256 + bias + residual (bulk-tensor epilogue), and the 1x1 256->768 qkv layer
x16, w = rnd(B, 16, 16, 256), rnd(256, 256, 3, 3, dtype=f32) * 0.02
bias = torch.zeros(256, device=dev)
res = rnd(B, 16, 16, 256)
F.tap_gemm(x16, F.pack_weight(w, 0), T3, B, 16, 16, bias=bias, residual=res)
wq = rnd(768, 256, 1, dtype=f32) * 0.02
F.tap_gemm(x16, F.pack_weight(wq, 0), T1, B, 16, 16, bias=torch.zeros(768, device=dev))
# conv_fwd_swap_kernel: 3x3 32x32 128->128 + residual
x32, w32 = rnd(B, 32, 32, 128), rnd(128, 128, 3, 3, dtype=f32) * 0.02
F.tap_gemm(x32, F.pack_weight(w32, 0), T3, B, 32, 32, bias=torch.zeros(128, device=dev), residual=rnd(B, 32, 32, 128))
# conv_wgrad_kernel + wgrad_reduce_kernel: 3x3 16x16 256->256
F.tap_wgrad(x16, rnd(B, 16, 16, 256), T3, B, 16, 16, 256, 256, (256, 256, 3, 3))
# GroupNorm pipe kernels: 32x32 C=128
gamma, beta = torch.ones(128, device=dev), torch.zeros(128, device=dev)
y, mean, rstd = F.gn_silu_fwd(x32, gamma, beta)
F.gn_silu_bwd(x32, rnd(B, 32, 32, 128), gamma, beta, mean, rstd, want_colsum=True)
# attention: T=256, 4 heads x 64
qkv = rnd(B, 256, 768)
o, lse = F.attn_fwd(qkv, 4)
F.attn_bwd(qkv, o, rnd(B, 256, 256), lse, 4)
torch.cuda.synchronize()
print("profiled one launch of each major kernel")
