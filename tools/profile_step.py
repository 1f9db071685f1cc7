"""Run exactly one eager training step and one eager reverse step of the bench workload between
cudaProfilerStart/Stop, for `ncu --profile-from-start off` (launch list / per-kernel capture)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticdeepdiffusionmodels_b200 import Engine, ops  # noqa: E402
from probabilisticdeepdiffusionmodels_b200.configs import MODEL_CONFIGS, synthetic_init_  # noqa: E402

B = int(os.environ.get("PDDM_B", "128"))
what = sys.argv[1] if len(sys.argv) > 1 else "both"
cfg = MODEL_CONFIGS["unet"]
eng = Engine(dict(cfg), {"lr": 1e-4}, mode="cosine", resolution=32, clip_while_generating=True, learn_sigma=True,
             log_loss_per_t=False)
synthetic_init_(eng.model, seed=1)
eng = eng.cuda()
opt = torch.optim.Adam(eng.model.parameters(), lr=1e-4, fused=True)
x = torch.rand(B, 3, 32, 32, device="cuda") * 2 - 1


def train_step():
    t = torch.randint(1, 1001, (B,), device="cuda")
    noise = torch.randn_like(x)
    opt.zero_grad(set_to_none=True)
    loss, _ = eng.loss_on(x, t, noise)
    loss.backward()
    opt.step()


def sample_step(xt):
    with torch.no_grad(), ops.frozen_weights():
        return eng.denoising_step(xt, 500)


for _ in range(3):
    train_step()
    sample_step(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if what in ("both", "train"):
    train_step()
if what in ("both", "sample"):
    sample_step(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", what)
