"""Oracle (test infrastructure): import shims so the UNMODIFIED reference can be imported in the build container.

The reference imports ``pytorch_lightning`` and ``matplotlib`` at module top level
(src/engine.py:9,23; src/utils.py:6) and neither is installed here.  ``install()`` puts
minimal stand-ins into ``sys.modules`` and ``/root/reference`` on ``sys.path``.  Only
``oracle/gen_golden.py`` (fixture generation, build container only) uses this; nothing on
the GPU box reads ``/root/reference``.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PDDM_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def install():
    if not available():
        raise RuntimeError(f"reference sources not found at {REFERENCE_ROOT}")
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def log(self, *a, **k):
                pass

            def save_hyperparameters(self, *a, **k):
                pass

            def optimizer_step(self, *a, **k):
                pass

        class Callback:
            pass

        pl.LightningModule = LightningModule
        pl.Callback = Callback
        sys.modules["pytorch_lightning"] = pl
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            sys.modules["wandb"] = types.ModuleType("wandb")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
