"""Exponential moving average of the model weights: same public surface as src/modules/ema.py:8-36
(``Ema(model, decay).module``, ``.update(model)``, ``.set(model)``; the shadow is a ``deepcopy`` in eval mode with
``requires_grad_(False)``, and its tensors live in the state_dict under ``ema.module.*``).

The update runs over ``state_dict()`` values like the reference (so buffers are averaged too) but as two
multi-tensor launches instead of ~3 launches per tensor."""
import copy

import torch
from torch import nn


class Ema(nn.Module):
    def __init__(self, model, decay=0.9999, device=None):
        super().__init__()
        self.module = copy.deepcopy(model).eval().requires_grad_(False)
        self.decay = decay
        self.device = device
        if device is not None:
            self.module.to(device=device)

    def _pairs(self, model):
        mine, theirs = self.module.state_dict(), model.state_dict()
        for k, e in mine.items():
            m = theirs[k]
            yield e, (m if self.device is None else m.to(device=self.device))

    @torch.no_grad()
    def update(self, model):
        floats_e, floats_m = [], []
        for e, m in self._pairs(model):
            if e.is_floating_point():
                floats_e.append(e)
                floats_m.append(m.detach())
            else:
                e.copy_(self.decay * e + (1.0 - self.decay) * m)
        if floats_e:
            torch._foreach_mul_(floats_e, self.decay)
            torch._foreach_add_(floats_e, floats_m, alpha=1.0 - self.decay)

    @torch.no_grad()
    def set(self, model):
        for e, m in self._pairs(model):
            e.copy_(m)
