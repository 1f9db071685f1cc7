"""Comparator, not a test (pytest does not collect it): the reference's algorithm as stock PyTorch ops (cuDNN /
cuBLAS / native GroupNorm / bmm-softmax attention) on the SAME GPU -- SURVEY.md section 8(d) "existing Blackwell
library kernels".  It runs the oracle's plain-torch restatement of the UNet (oracle/unet_ref.py) on cuda:0 for the
bench workload (CIFAR UNet, B=128, learned variance head, L_simple on the eps half, Adam) in fp32 with TF32 enabled
and under bf16 autocast, forward-only and forward+backward+Adam, and prints one JSON line.

    python tests/stock_torch_comparator.py [--batch 128] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params, unet_forward  # noqa: E402


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = MODEL_CONFIGS["unet"]
    arch = arch_from_config(32, **{k: v for k, v in cfg.items() if k != "name"}, learn_sigma=True)
    P = {k: v.to(dev).requires_grad_(True) for k, v in make_params(arch, seed=1).items()}
    opt = torch.optim.Adam(list(P.values()), lr=1e-4, fused=True)
    B = a.batch
    x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    res = {}
    for name, tf32, autocast in [("fp32_tf32", True, False), ("bf16_autocast", True, True)]:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True

        def fwd():
            t = torch.randint(1, 1001, (B,), device=dev)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                return unet_forward(P, arch, x, t)

        def train():
            noise = torch.randn_like(x)
            opt.zero_grad(set_to_none=True)
            out = fwd()
            loss = ((out[:, :3].float() - noise) ** 2).mean()
            loss.backward()
            opt.step()

        with torch.no_grad():
            ms_f = timed(fwd, a.steps)
        ms_t = timed(train, a.steps)
        res[name] = {"fwd_ms": ms_f, "fwd_img_s": B / ms_f * 1e3, "train_ms": ms_t, "train_img_s": B / ms_t * 1e3}
    print(json.dumps({"what": "stock torch ops on the same GPU (oracle restatement, eager)", "batch": B,
                      "torch": torch.__version__, **res}))


if __name__ == "__main__":
    main()
