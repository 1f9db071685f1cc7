"""Oracle (test infrastructure): the reference Engine's train / sample / NLL drivers, restated on
top of ``oracle.unet_ref.unet_forward`` and ``oracle.diffusion_ref.DiffusionRef`` (CPU fp32).
"""
import torch

from .diffusion_ref import DiffusionRef, mean_flat
from .unet_ref import unet_forward


def train_loss(P, arch, diff: DiffusionRef, x0, t, noise, weights=None, learn_sigma=False):
    """training_step math with injected t / noise  (src/engine.py:279-289)."""
    x_t = diff.q_sample(x0, noise, t)
    out = unet_forward(P, arch, x_t, t)
    if learn_sigma:
        return diff.loss_hybrid(x0, x_t, t, noise, out, weights)
    return diff.loss_simple(out, noise, weights)


def denoising_step(P, arch, diff, x_t, t_step, z, clip, mean_only=False, learn_sigma=False):
    """src/engine.py:385-397: float32 ``t`` vector, mean via clip / no-clip path, x = mean - sigma*z."""
    t = t_step * torch.ones(x_t.shape[0])
    out = unet_forward(P, arch, x_t, t)
    if learn_sigma:
        eps, v = out.chunk(2, dim=1)
        return diff.p_sample_step_learned(x_t, t_step, eps, v, z, clip=clip, mean_only=mean_only)
    return diff.p_sample_step(x_t, t_step, out, z, clip=clip, mean_only=mean_only)


@torch.no_grad()
def sample_chain(P, arch, diff, x_t, t_start, zs, steps_to_return=(1,), clip=True, mean_only=False,
                 learn_sigma=False):
    """sample_and_return_steps (src/engine.py:510-554) with the per-step noise ``zs[k]`` injected
    (``zs[k]`` is used at step t = t_start - k; the t == 1 step uses none). Returns [B, S, C, H, W]."""
    out = torch.zeros((x_t.shape[0], len(steps_to_return)) + tuple(x_t.shape[1:]))
    idx = 0
    for k, t in enumerate(range(t_start, 0, -1)):
        z = zs[k] if (t > 1 and not mean_only) else 0
        x_t = denoising_step(P, arch, diff, x_t, t, z, clip, mean_only, learn_sigma)
        if t in steps_to_return:
            out[:, idx] = x_t
            idx += 1
    return out


@torch.no_grad()
def calculate_likelihood(P, arch, diff, x0):
    """calculate_likelihood (src/engine.py:417-506) drawing noise from the global torch RNG in the
    reference's order: L_0's noise first, then one draw per t = 2..T."""
    b = x0.shape[0]
    one = torch.ones(b, dtype=torch.int64)
    noise = torch.randn_like(x0)
    x1 = diff.q_sample(x0, noise, one)
    L0 = diff.L_0(x0, x1, unet_forward(P, arch, x1, one))
    Lint = []
    for t_step in range(2, diff.T + 1):
        t = one * t_step
        noise = torch.randn_like(x0)
        x_t = diff.q_sample(x0, noise, t)
        Lint.append(diff.L_t(x0, x_t, t_step, unet_forward(P, arch, x_t, t)))
    LT = diff.L_T(x0)
    Lsum = torch.sum(torch.stack(Lint), dim=0)
    return {"L_0": L0.mean(0), "L_T": LT.mean(0), "L_intermediate": Lsum,
            "nll": torch.mean(L0 + Lsum + LT, dim=0), "L_intermediate_list": Lint}
