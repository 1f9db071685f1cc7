"""Oracle (test infrastructure): fp32 CPU restatement of the reference's diffusion math.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  All arithmetic is torch fp32 on CPU, mirroring the
reference's dtype choices (e.g. the fp32 cumprod, python-float64 cosine betas
rounded to fp32).  Timesteps are 1-indexed; tables are read at ``[t-1]``.
"""
import math

import numpy as np
import torch

LN2 = float(np.log(2.0))


# ----------------------------------------------------------------------------------------------
# noise schedules  (src/engine.py:26-76)
# ----------------------------------------------------------------------------------------------
def cosine_alpha_bar(u):
    """src/engine.py:33-34"""
    return math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2


def _betas_from_alpha_bar(fn, steps, max_beta):
    """src/engine.py:37-43 (python float64 arithmetic, later rounded to fp32)."""
    out = []
    for i in range(steps):
        lo, hi = i / steps, (i + 1) / steps
        out.append(min(1 - fn(hi) / fn(lo), max_beta))
    return out


def get_betas(beta_start=None, beta_end=None, diffusion_steps=1000, mode="linear", max_beta=0.999,
              custom_alpha_bar=None):
    """src/engine.py:55-76"""
    if mode == "linear":
        if beta_start is None or beta_end is None:
            s = 1000 / diffusion_steps
            beta_start, beta_end = s * 0.0001, s * 0.02
        return torch.linspace(beta_start, beta_end, diffusion_steps)
    if mode == "cosine":
        return torch.tensor(_betas_from_alpha_bar(cosine_alpha_bar, diffusion_steps, max_beta))
    if mode == "mixed":
        # src/engine.py:46-52: average of the linear alpha-bar (extrapolated by one entry) and cosine
        lin = torch.cumprod(1 - get_betas(None, None, diffusion_steps, "linear"), 0)
        last = 2 * lin[-1] - lin[-2]
        lin = torch.cat([lin, torch.tensor([1]) * last])
        cos = torch.tensor([cosine_alpha_bar(k / diffusion_steps) for k in range(diffusion_steps + 1)])
        mixed = 0.5 * lin + 0.5 * cos
        return torch.tensor(_betas_from_alpha_bar(lambda u: mixed[int(u * diffusion_steps)],
                                                  diffusion_steps, max_beta))
    if mode == "custom":
        return torch.tensor(_betas_from_alpha_bar(custom_alpha_bar, diffusion_steps, max_beta))
    raise ValueError(f"Wrong beta mode: {mode}")


TABLE_NAMES = (
    "betas", "alphas", "alphas_sqrt", "alphas_hat", "alphas_hat_sqrt", "one_min_alphas_hat_sqrt",
    "alphas_hat_prev", "alphas_hat_next", "posterior_variance", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2", "denoising_coef",
)


def make_tables(betas):
    """The fp32 coefficient tables of Engine.__init__  (src/engine.py:121-150)."""
    t = {}
    t["betas"] = betas
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


def _g(table, t):
    """per-sample gather ``table[t-1].view(-1,1,1,1)`` (src/engine.py:255-256)."""
    t = torch.as_tensor(t).long().reshape(-1)
    return table[t - 1].view(-1, 1, 1, 1)


# ----------------------------------------------------------------------------------------------
# helpers from src/modules/nn.py and src/utils.py
# ----------------------------------------------------------------------------------------------
def mean_flat(x):
    """src/utils.py:13-17"""
    return x.mean(dim=list(range(1, x.dim())))


def timestep_embedding(timesteps, dim, max_period=10000):
    """src/modules/nn.py:104-122 -- [cos | sin], zero pad if dim is odd."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None].to(timesteps.device)  # (.to(device): src/modules/nn.py:116)
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def normal_kl(mean1, logvar1, mean2, logvar2):
    """src/utils.py:50-77 (nats)."""
    ref = next(o for o in (mean1, logvar1, mean2, logvar2) if isinstance(o, torch.Tensor))
    logvar1, logvar2 = [v if isinstance(v, torch.Tensor) else torch.tensor(v).to(ref) for v in (logvar1, logvar2)]
    return 0.5 * (-1.0 + logvar2 - logvar1 + torch.exp(logvar1 - logvar2)
                  + ((mean1 - mean2) ** 2) * torch.exp(-logvar2))


def approx_standard_normal_cdf(x):
    """src/utils.py:80-85"""
    return 0.5 * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * torch.pow(x, 3))))


def discretized_gaussian_log_likelihood(x, means, log_scales):
    """src/utils.py:88-115 (nats, bins of +-1/255, clamp 1e-12, edge cases at |x|>0.999)."""
    centered = x - means
    inv_stdv = torch.exp(-log_scales)
    cdf_plus = approx_standard_normal_cdf(inv_stdv * (centered + 1.0 / 255.0))
    cdf_min = approx_standard_normal_cdf(inv_stdv * (centered - 1.0 / 255.0))
    log_cdf_plus = torch.log(cdf_plus.clamp(min=1e-12))
    log_one_minus_cdf_min = torch.log((1.0 - cdf_min).clamp(min=1e-12))
    delta = cdf_plus - cdf_min
    return torch.where(x < -0.999, log_cdf_plus,
                       torch.where(x > 0.999, log_one_minus_cdf_min, torch.log(delta.clamp(min=1e-12))))


# ----------------------------------------------------------------------------------------------
# Engine math (src/engine.py:251-397, 437-506)
# ----------------------------------------------------------------------------------------------
class DiffusionRef:
    """Schedule tables + the Engine's elementwise math, without the network."""

    def __init__(self, diffusion_steps=1000, beta_start=None, beta_end=None, mode="linear", max_beta=0.999,
                 sigma_mode="beta"):
        self.T = diffusion_steps
        self.sigma_mode = sigma_mode
        self.tables = make_tables(get_betas(beta_start, beta_end, diffusion_steps, mode, max_beta))
        for k, v in self.tables.items():
            setattr(self, k, v)

    # src/engine.py:251-261
    def q_sample(self, x0, noise, t):
        return x0 * _g(self.alphas_hat_sqrt, t) + noise * _g(self.one_min_alphas_hat_sqrt, t)

    # src/engine.py:263-277
    def loss_simple(self, eps_pred, noise, weights=None):
        per = mean_flat(torch.square(noise - eps_pred))
        total = torch.sum(weights * per) if weights is not None else torch.mean(per)
        return total, per

    # src/engine.py:354-361
    def sigma(self, t_step):
        tab = self.betas if self.sigma_mode == "beta" else self.posterior_variance
        if self.sigma_mode not in ("beta", "beta_tilde"):
            raise ValueError(f"Wrong sigma mode: {self.sigma_mode}")
        return torch.sqrt(tab[t_step - 1])

    # src/engine.py:363-368
    def xstart_from_eps(self, x_t, t, eps, clip=False):
        x = _g(self.sqrt_recip_alphas_cumprod, t) * x_t - _g(self.sqrt_recipm1_alphas_cumprod, t) * eps
        return x.clamp(-1, 1) if clip else x

    # src/engine.py:477-490
    def q_posterior(self, t, x0, x_t):
        mean = x0 * _g(self.posterior_mean_coef1, t) + x_t * _g(self.posterior_mean_coef2, t)
        return mean, _g(self.posterior_variance, t)

    # src/engine.py:370-381
    def model_mean(self, x_t, t, eps, clip=False):
        if clip:
            return self.q_posterior(t, self.xstart_from_eps(x_t, t, eps, clip=True), x_t)[0]
        return (x_t - eps * _g(self.denoising_coef, t)) / _g(self.alphas_sqrt, t)

    # src/engine.py:385-397  (note the MINUS and z = 0 at t == 1)
    def p_sample_step(self, x_t, t_step, eps, z, clip=False, mean_only=False):
        mean = self.model_mean(x_t, t_step, eps, clip=clip)
        if mean_only or t_step <= 1:
            return mean
        return mean - self.sigma(t_step) * z

    # ---- NLL terms (fixed variance), bits/dim ------------------------------------------------
    # src/engine.py:437-444
    def L_T(self, x0):
        mean, std = x0 * _g(self.alphas_hat_sqrt, self.T), _g(self.one_min_alphas_hat_sqrt, self.T)
        return mean_flat(normal_kl(mean, 2 * torch.log(std), 0.0, 0.0)) / LN2

    # src/engine.py:446-475 (one t)
    def L_t(self, x0, x_t, t_step, eps):
        t = torch.full((x0.shape[0],), t_step, dtype=torch.int64)
        mean_t, var_t = self.q_posterior(t, x0, x_t)
        pmean = self.model_mean(x_t, t_step, eps, clip=False)
        plogvar = 2 * torch.log(self.sigma(t_step))
        kl = normal_kl(mean_t, torch.log(var_t) * torch.ones_like(mean_t), pmean, plogvar * torch.ones_like(mean_t))
        return mean_flat(kl) / LN2

    # src/engine.py:492-506
    def L_0(self, x0, x_1, eps):
        pmean = self.model_mean(x_1, 1, eps, clip=False)
        logscale = torch.log(self.sigma(1)) * torch.ones_like(x0)
        return mean_flat(-discretized_gaussian_log_likelihood(x0, pmean, logscale)) / LN2

    # ---- learned variance / L_hybrid: SURVEY.md Appendix C (no reference implementation) -----
    def posterior_log_variance_clipped(self):
        pv = self.posterior_variance
        return torch.log(torch.cat([pv[1:2], pv[1:]]))

    def model_logvar(self, v, t):
        min_log = _g(self.posterior_log_variance_clipped(), t)
        max_log = _g(torch.log(self.betas), t)
        frac = (v + 1) / 2
        return frac * max_log + (1 - frac) * min_log

    def vb_term(self, x0, x_t, t, eps, v):
        """bits/dim per sample; KL for t>1, discretised decoder NLL at t==1 (App. C step 5)."""
        t = torch.as_tensor(t).long().reshape(-1)
        true_mean, _ = self.q_posterior(t, x0, x_t)
        true_logvar = _g(self.posterior_log_variance_clipped(), t)
        pmean = self.q_posterior(t, self.xstart_from_eps(x_t, t, eps.detach(), clip=False), x_t)[0]
        logvar = self.model_logvar(v, t)
        kl = mean_flat(normal_kl(true_mean, true_logvar, pmean, logvar)) / LN2
        nll0 = mean_flat(-discretized_gaussian_log_likelihood(x0, pmean, 0.5 * logvar)) / LN2
        return torch.where(t == 1, nll0, kl)

    def loss_hybrid(self, x0, x_t, t, noise, model_out, weights=None):
        """L_simple + (T/1000) * vb  (App. C step 6). model_out = [eps | v] on dim 1."""
        eps, v = model_out.chunk(2, dim=1)
        per = mean_flat(torch.square(noise - eps)) + (self.T / 1000.0) * self.vb_term(x0, x_t, t, eps, v)
        total = torch.sum(weights * per) if weights is not None else torch.mean(per)
        return total, per

    def p_sample_step_learned(self, x_t, t_step, eps, v, z, clip=False, mean_only=False):
        """App. C step 7: x_{t-1} = mean - exp(0.5*logvar) * z, z = 0 at t == 1."""
        mean = self.model_mean(x_t, t_step, eps, clip=clip)
        if mean_only or t_step <= 1:
            return mean
        t = torch.full((x_t.shape[0],), t_step, dtype=torch.int64)
        return mean - torch.exp(0.5 * self.model_logvar(v, t)) * z
