"""Replay the captured training step (and optionally the captured reverse step) once between
cudaProfilerStart/Stop so that `ncu --graph-profiling node --cache-control none --profile-from-start off`
lists every kernel NODE of the graph with warm caches, i.e. close to what a replay really costs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticdeepdiffusionmodels_b200 import Engine  # noqa: E402
from probabilisticdeepdiffusionmodels_b200.configs import MODEL_CONFIGS, synthetic_init_  # noqa: E402

B = int(os.environ.get("PDDM_B", "128"))
cfg = MODEL_CONFIGS["unet"]
eng = Engine(dict(cfg), {"lr": 1e-4}, mode="cosine", resolution=32, clip_while_generating=True, learn_sigma=True,
             log_loss_per_t=False)
synthetic_init_(eng.model, seed=1)
eng = eng.cuda()
x = torch.rand(B, 3, 32, 32, device="cuda") * 2 - 1
step = eng.capture_train_step((B, 3, 32, 32))
for _ in range(3):
    step(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one replay of the captured train step")
