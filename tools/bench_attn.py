"""Micro-benchmark of the fused attention kernels on the shapes of the CIFAR / CelebA UNets (B=128, 4 heads): each
case is 10 launches captured in a CUDA graph, inputs rotating over enough buffers to exceed the 126 MB L2.  Prints
time per launch and algorithmic TFLOP/s (fwd 4*T^2*d, bwd 10*T^2*d per sample and head).
    python tools/bench_attn.py             # timing table
    python tools/bench_attn.py --profile   # one plain launch per kernel at T=256 d=64 (for ncu)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticdeepdiffusionmodels_b200 import functional as F  # noqa: E402

# (B, T, heads, d)
SHAPES = [(128, 256, 4, 64), (128, 64, 4, 64), (128, 16, 4, 64), (64, 256, 4, 96), (64, 64, 4, 128)]


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
    shapes = SHAPES[:1] if "--profile" in sys.argv else SHAPES
    for B, T, heads, d in shapes:
        Cc = heads * d
        nbuf = 1 if "--profile" in sys.argv else max(2, min(8, int(300e6 / (B * T * Cc * 2 * 5)) + 1))
        qkv = [torch.randn(B, T, 3 * Cc, device=dev).bfloat16() for _ in range(nbuf)]
        dout = [torch.randn(B, T, Cc, device=dev).bfloat16() for _ in range(nbuf)]
        out, lse = F.attn_fwd(qkv[0], heads)
        if "--profile" in sys.argv:
            F.attn_bwd(qkv[0], out, dout[0], lse, heads)
            torch.cuda.synchronize()
            continue
        tf = timed(lambda i: F.attn_fwd(qkv[i], heads), nbuf)
        tb = timed(lambda i: F.attn_bwd(qkv[i], out, dout[i], lse, heads), nbuf)
        fl = B * heads * T * T * d
        print(f"B={B} T={T} heads={heads} d={d}: fwd {tf:7.1f} us ({4 * fl / tf * 1e-6:6.1f} TF/s)   "
              f"bwd {tb:7.1f} us ({10 * fl / tb * 1e-6:6.1f} TF/s)", flush=True)


if __name__ == "__main__":
    main()
