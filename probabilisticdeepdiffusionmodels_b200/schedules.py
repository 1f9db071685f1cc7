"""Noise schedules and the Engine's fp32 coefficient tables (host side, init only).

Same functions, argument names and numerics as src/engine.py:26-76 and :121-150 -- e.g. the cosine betas are
python-float64 values rounded to fp32 by ``torch.tensor`` and alpha-bar is an fp32 ``cumprod`` -- because the
tables are the constants every kernel reads (bit-exact against tests/golden/schedules.npz).
"""
import math

import numpy as np
import torch

TABLE_NAMES = (
    "betas", "alphas", "alphas_sqrt", "alphas_hat", "alphas_hat_sqrt", "one_min_alphas_hat_sqrt", "alphas_hat_prev",
    "alphas_hat_next", "posterior_variance", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
    "posterior_mean_coef1", "posterior_mean_coef2", "denoising_coef")


def get_linear_alphas_bar(diffusion_steps):
    """src/engine.py:26-30"""
    return torch.cumprod(1 - get_betas(None, None, diffusion_steps, "linear"), 0)


def cosine_alpha_bar(t):
    """src/engine.py:33-34"""
    return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2


def betas_for_alpha_bar(alpha_bar, diffusion_steps, max_beta):
    """src/engine.py:37-43"""
    return [min(1 - alpha_bar((i + 1) / diffusion_steps) / alpha_bar(i / diffusion_steps), max_beta)
            for i in range(diffusion_steps)]


def mixed_alpha_bar(diffusion_steps):
    """src/engine.py:46-52"""
    lin = get_linear_alphas_bar(diffusion_steps)
    lin = torch.cat([lin, torch.tensor([1]) * (2 * lin[-1] - lin[-2])])
    cos = torch.tensor([cosine_alpha_bar(t / diffusion_steps) for t in range(diffusion_steps + 1)])
    return 0.5 * lin + 0.5 * cos


def get_betas(beta_start=None, beta_end=None, diffusion_steps=1000, mode="linear", max_beta=0.999,
              custom_alpha_bar=None):
    """src/engine.py:55-76"""
    if mode == "linear":
        if beta_start is None or beta_end is None:
            scale = 1000 / diffusion_steps
            beta_start, beta_end = scale * 0.0001, scale * 0.02
        return torch.linspace(beta_start, beta_end, diffusion_steps)
    elif mode == "cosine":
        return torch.tensor(betas_for_alpha_bar(cosine_alpha_bar, diffusion_steps, max_beta))
    elif mode == "mixed":
        alpha_bar = mixed_alpha_bar(diffusion_steps)
        return torch.tensor(betas_for_alpha_bar(lambda t: alpha_bar[int(t * diffusion_steps)], diffusion_steps,
                                                max_beta))
    elif mode == "custom":
        return torch.tensor(betas_for_alpha_bar(custom_alpha_bar, diffusion_steps, max_beta))
    raise ValueError(f"Wrong beta mode: {mode}")


def make_tables(betas):
    """The coefficient tables of Engine.__init__ (src/engine.py:121-150), all fp32 on the CPU."""
    t = {"betas": betas}
    t["alphas"] = 1 - betas
    t["alphas_sqrt"] = torch.sqrt(t["alphas"])
    t["alphas_hat"] = torch.cumprod(t["alphas"], 0)
    t["alphas_hat_sqrt"] = torch.sqrt(t["alphas_hat"])
    t["one_min_alphas_hat_sqrt"] = torch.sqrt(1 - t["alphas_hat"])
    t["alphas_hat_prev"] = torch.Tensor(np.append(1.0, t["alphas_hat"][:-1].numpy()))
    t["alphas_hat_next"] = torch.Tensor(np.append(t["alphas_hat"][1:].numpy(), 0.0))
    t["posterior_variance"] = betas * (1.0 - t["alphas_hat_prev"]) / (1.0 - t["alphas_hat"])
    t["sqrt_recip_alphas_cumprod"] = torch.sqrt(1.0 / t["alphas_hat"])
    t["sqrt_recipm1_alphas_cumprod"] = torch.sqrt(1.0 / t["alphas_hat"] - 1)
    t["posterior_mean_coef1"] = betas * torch.sqrt(t["alphas_hat_prev"]) / (1.0 - t["alphas_hat"])
    t["posterior_mean_coef2"] = (1.0 - t["alphas_hat_prev"]) * t["alphas_sqrt"] / (1.0 - t["alphas_hat"])
    t["denoising_coef"] = betas / t["one_min_alphas_hat_sqrt"]
    return t
