"""One shape of tools/bench_gn.py (forward and backward), for knob sweeps:  python tools/bench_gn_one.py B HW C"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticdeepdiffusionmodels_b200 import functional as F  # noqa: E402
from bench_gn import timed  # noqa: E402

B, HW, C = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda")
nbuf = max(2, min(10, int(400e6 / (B * HW * C * 2 * 3)) + 1))
xs = [torch.randn(B, HW, C, device=dev).bfloat16() for _ in range(nbuf)]
dys = [torch.randn(B, HW, C, device=dev).bfloat16() for _ in range(nbuf)]
gamma = torch.randn(C, device=dev) * 0.1 + 1
beta = torch.randn(C, device=dev) * 0.1
y, mean, rstd = F.gn_silu_fwd(xs[0], gamma, beta)
tf = timed(lambda i: F.gn_silu_fwd(xs[i], gamma, beta), nbuf)
tb = timed(lambda i: F.gn_silu_bwd(xs[i], dys[i], gamma, beta, mean, rstd, want_colsum=True), nbuf)
n = B * HW * C
print(f"{os.environ.get('PDDM_GN_DBG', '-'):>3} cc={os.environ.get('PDDM_GN_CC', '-'):>3} B={B} HW={HW} C={C} fwd {tf:7.1f} us "
      f"{4 * n / tf * 1e-3:6.0f} GB/s  bwd {tb:7.1f} us {6 * n / tb * 1e-3:6.0f} GB/s", flush=True)
