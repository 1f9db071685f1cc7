"""Micro-benchmark of the GroupNorm+SiLU kernels on the shapes of the CIFAR UNet (B=128): each case is 10 launches
captured in a CUDA graph; inputs of successive launches rotate over enough buffers to exceed the 126 MB L2.
Prints time per launch and algorithmic bandwidth (fwd 4 B/element, bwd 6 B/element).
    python tools/bench_gn.py            # PDDM_GN_STREAM=1 selects the register-streaming kernels"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticdeepdiffusionmodels_b200 import functional as F  # noqa: E402

SHAPES = [(128, 32 * 32, 128), (128, 32 * 32, 256), (128, 32 * 32, 384), (128, 16 * 16, 256), (128, 16 * 16, 512),
          (128, 8 * 8, 256), (128, 8 * 8, 512), (128, 4 * 4, 256), (128, 4 * 4, 512)]


def timed(fn, nbuf):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(nbuf):
            fn(i)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(10):
                fn(i % nbuf)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(3):
            g.replay()
        e1.record(s)
    e1.synchronize()
    return e0.elapsed_time(e1) / 30 * 1e3


def main():
    dev = torch.device("cuda")
    if "--profile" in sys.argv:  # one plain launch of each kernel on two shapes, for ncu
        for B, HW, C in SHAPES[:2]:
            x = torch.randn(B, HW, C, device=dev).bfloat16()
            dy = torch.randn(B, HW, C, device=dev).bfloat16()
            gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            y, mean, rstd = F.gn_silu_fwd(x, gamma, beta)
            F.gn_silu_bwd(x, dy, gamma, beta, mean, rstd, want_colsum=True)
        torch.cuda.synchronize()
        return
    tot_f = tot_b = 0.0
    for B, HW, C in SHAPES:
        nbuf = max(2, min(10, int(400e6 / (B * HW * C * 2 * 3)) + 1))
        xs = [torch.randn(B, HW, C, device=dev).bfloat16() for _ in range(nbuf)]
        dys = [torch.randn(B, HW, C, device=dev).bfloat16() for _ in range(nbuf)]
        gamma = torch.randn(C, device=dev) * 0.1 + 1
        beta = torch.randn(C, device=dev) * 0.1
        y, mean, rstd = F.gn_silu_fwd(xs[0], gamma, beta)
        tf = timed(lambda i: F.gn_silu_fwd(xs[i], gamma, beta), nbuf)
        tb = timed(lambda i: F.gn_silu_bwd(xs[i], dys[i], gamma, beta, mean, rstd, want_colsum=True), nbuf)
        n = B * HW * C
        tot_f += tf
        tot_b += tb
        extra = ""
        if F.gn_pipe_slots(B, HW, C, 32, 0, 3) >= 2 and not os.environ.get("PDDM_GN_NOPIPE"):
            # backward with the residual-branch gradient fused in (8 B/element) and per-sample partial sums
            pg = torch.empty(B, C, device=dev)
            pb = torch.empty(B, C, device=dev)
            tg = timed(lambda i: F.gn_silu_bwd(xs[i], dys[i], gamma, beta, mean, rstd, want_colsum=True,
                                               gres=dys[(i + 1) % nbuf], part_dgamma=pg, part_dbeta=pb), nbuf)
            extra = f"   bwd+gres {tg:7.1f} us {8 * n / tg * 1e-3:7.0f} GB/s"
        print(f"B={B} HW={HW:5d} C={C:4d}  slots {F.gn_pipe_slots(B, HW, C, 32, 0, 1)}/{F.gn_pipe_slots(B, HW, C, 32, 0, 2)}"
              f"  fwd {tf:7.1f} us {4 * n / tf * 1e-3:7.0f} GB/s   "
              f"bwd {tb:7.1f} us {6 * n / tb * 1e-3:7.0f} GB/s{extra}", flush=True)
    print(f"sum fwd {tot_f:.1f} us, bwd {tot_b:.1f} us")


if __name__ == "__main__":
    main()
