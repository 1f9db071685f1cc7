"""Calibration: what plain torch elementwise kernels (copy / add / silu) reach on GroupNorm-sized bf16 tensors in
the same graph-replay harness as tools/bench_gn.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_gn import timed
dev = torch.device('cuda')
for B, HW, C in [(128, 1024, 128), (128, 1024, 256), (128, 256, 256)]:
    nbuf = 6
    xs = [torch.randn(B, HW, C, device=dev).bfloat16() for _ in range(nbuf)]
    ys = [torch.empty_like(xs[0]) for _ in range(nbuf)]
    t = timed(lambda i: ys[i].copy_(xs[i]), nbuf)
    n = B * HW * C
    print(f"copy  {B}x{HW}x{C}: {t:.1f} us  {4*n/t*1e-3:.0f} GB/s")
    t = timed(lambda i: torch.add(xs[i], xs[(i+1)%nbuf], out=ys[i]), nbuf)
    print(f"add   {B}x{HW}x{C}: {t:.1f} us  {6*n/t*1e-3:.0f} GB/s")
    t = timed(lambda i: torch.nn.functional.silu(xs[i]), nbuf)
    print(f"silu  {B}x{HW}x{C}: {t:.1f} us  {4*n/t*1e-3:.0f} GB/s")
