"""API-compatible helpers of src/utils.py (mean_flat :13-17, get_generator_if_specified :41-47, normal_kl :50-77,
approx_standard_normal_cdf :80-85, discretized_gaussian_log_likelihood :88-115).

These broadcasting torch expressions exist for callers that want the individual terms on arbitrary tensors; the
training / sampling / NLL hot loops never call them -- they use the fused ``pddm_vlb_terms`` kernel, which the
GPU tests check against exactly these formulas (through the oracle)."""
import numpy as np
import torch

_SQRT_2_OVER_PI = float(np.sqrt(2.0 / np.pi))


def mean_flat(tensor):
    return tensor.mean(dim=list(range(1, tensor.dim())))


def get_generator_if_specified(seed=None, device="cpu"):
    if seed is None:
        return None
    return torch.Generator(device=device).manual_seed(seed)


def normal_kl(mean1, logvar1, mean2, logvar2):
    """KL(N(mean1, e^logvar1) || N(mean2, e^logvar2)) in nats; scalars are promoted next to the first tensor."""
    ref = next((o for o in (mean1, logvar1, mean2, logvar2) if isinstance(o, torch.Tensor)), None)
    assert ref is not None, "at least one argument must be a Tensor"
    lv1, lv2 = (v if isinstance(v, torch.Tensor) else torch.tensor(v).to(ref) for v in (logvar1, logvar2))
    return 0.5 * (-1.0 + lv2 - lv1 + torch.exp(lv1 - lv2) + ((mean1 - mean2) ** 2) * torch.exp(-lv2))


def approx_standard_normal_cdf(x):
    return 0.5 * (1.0 + torch.tanh(_SQRT_2_OVER_PI * (x + 0.044715 * torch.pow(x, 3))))


def discretized_gaussian_log_likelihood(x, means, log_scales):
    """log-probability (nats) of uint8-quantised data x in [-1, 1] under N(means, e^{2 log_scales}), bins of 2/255."""
    assert x.shape == means.shape == log_scales.shape
    centered, inv_std = x - means, torch.exp(-log_scales)
    cdf_hi = approx_standard_normal_cdf(inv_std * (centered + 1.0 / 255.0))
    cdf_lo = approx_standard_normal_cdf(inv_std * (centered - 1.0 / 255.0))
    log_hi = torch.log(cdf_hi.clamp(min=1e-12))
    log_one_minus_lo = torch.log((1.0 - cdf_lo).clamp(min=1e-12))
    mid = torch.log((cdf_hi - cdf_lo).clamp(min=1e-12))
    return torch.where(x < -0.999, log_hi, torch.where(x > 0.999, log_one_minus_lo, mid))
