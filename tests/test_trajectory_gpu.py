"""1000-step reverse-diffusion trajectories against the UNMODIFIED reference (src/engine.py:399-403, 510-554).

tests/golden/traj1000.npz (``python -m oracle.gen_golden traj``) holds full fp32 chains of the reference from a
fixed x_T with the per-step z drawn from a seeded CPU generator -- BASELINE configs[0]'s small model (batch 2, linear
and cosine schedule) and configs[1]'s CIFAR UNet (batch 1, cosine, where sqrt_recip_alphas_cumprod[999] = 20291
amplifies the first steps) -- at t = 900 ... 1, and the reference's OWN drift when its network runs under bf16
autocast with everything else unchanged: the noise floor of a bf16-operand implementation of the same chain.

Stated tolerance (north_star: "1000-step sample trajectories must agree within a stated per-pixel tolerance"), on
pixel values in [-1, 1]: at every recorded step  max |delta| <= TOL_MAX  and  mean |delta| <= TOL_MEAN  (1.5 x what
the CUDA path measured on B200: max 3.1e-2 / 5.2e-2 / 5.0e-2, mean 1.9e-3 / 1.9e-3 / 2.2e-3), and the final image's
mean error is additionally held to 1.5 x the reference's own bf16 drift (measured ratio: 0.98 ... 1.00 -- the chain
error is dominated by rounding x_t and the activations to bf16, which both sides do)."""
import numpy as np
import pytest
import torch

from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params
from _parity import within

pytestmark = pytest.mark.gpu

CASES = {  # tag -> (config, resolution, schedule, parameter seed, TOL_MAX, TOL_MEAN)
    "small_linear": ("unet_small_grey", 28, "linear", 21, 4.7e-2, 2.9e-3),
    "small_cosine": ("unet_small_grey", 28, "cosine", 21, 7.7e-2, 2.9e-3),
    "cifar_cosine": ("unet", 32, "cosine", 11, 7.5e-2, 3.4e-3),
}


@pytest.mark.parametrize("tag", list(CASES))
def test_1000_step_trajectory_matches_reference(golden, tag):
    from probabilisticdeepdiffusionmodels_b200 import Engine
    g = golden["traj1000"]
    name, res, mode, pseed, tol_max, tol_mean = CASES[tag]
    cfg = MODEL_CONFIGS[name]
    eng = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=1000, mode=mode, resolution=res, clip_while_generating=True,
                 sigma_mode="beta")
    arch = arch_from_config(res, **{k: v for k, v in cfg.items() if k != "name"})
    eng.model.load_state_dict(make_params(arch, seed=pseed))
    eng = eng.to("cuda")
    xT = torch.from_numpy(g[f"{tag}_xT"])
    steps = tuple(int(s) for s in g["steps"])
    # the reference draws z ~ N(0, I) of x_T's shape from ONE generator for t = 1000 ... 2 (src/engine.py:388-393)
    gen = torch.Generator().manual_seed(int(g[f"{tag}_zseed"]))
    zs = torch.stack([torch.randn(tuple(xT.shape), generator=gen) for _ in range(999)]).cuda()
    out = eng.sample_and_return_steps(xT.cuda(), t_start=1000, steps_to_return=steps, fixed_noise=zs)
    ref = g[f"{tag}_chain"]
    assert tuple(out.shape) == ref.shape
    err = np.abs(out.numpy() - ref)  # [B, steps, C, H, W]
    emax, emean = err.max(axis=(0, 2, 3, 4)), err.mean(axis=(0, 2, 3, 4))
    fmax, fmean = g[f"{tag}_floor_max"], g[f"{tag}_floor_mean"]
    for k, s in enumerate(steps):
        print(f"[parity] traj1000[{tag}] t={s:4d}: max |d| {emax[k]:.3e} (reference bf16 drift {fmax[k]:.3e})   "
              f"mean |d| {emean[k]:.3e} (reference bf16 drift {fmean[k]:.3e})", flush=True)
    assert np.isfinite(out.numpy()).all() and np.abs(out.numpy()[:, -1]).max() <= 1.0 + 1e-6  # clipped chain
    within(f"traj1000[{tag}] max |delta| per pixel over all recorded steps", float(emax.max()), tol_max)
    within(f"traj1000[{tag}] mean |delta| per pixel over all recorded steps", float(emean.max()), tol_mean)
    within(f"traj1000[{tag}] final image mean |delta| vs 1.5 x reference bf16 drift", float(emean[-1]),
           1.5 * float(fmean[-1]))
