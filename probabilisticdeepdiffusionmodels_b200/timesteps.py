"""Timestep selection for training: the uniform and loss-aware ("importance") samplers and the per-timestep loss
history they feed on.  Behavioural mirror of src/sampling/uniform_sampler.py:4-10,
src/sampling/importance_sampler.py:7-41 and src/modules/stepwise_log.py:4-37 (host-side bookkeeping; the only
device work is one ``randint``).  Timesteps are 1-indexed.
"""
import numpy as np
import torch


class StepwiseLog:
    """Per-timestep metric history with running mean / RMS / count arrays (index = t - 1).

    ``max_keep``: the reference truncates a timestep's history to the last ``max_keep`` entries whenever the
    *number of timesteps* exceeds ``max_keep`` (it tests ``len(self.metric_per_t)``, stepwise_log.py:19-20); that
    condition is reproduced so that the importance weights evolve identically."""

    def __init__(self, diffusion_steps, max_keep=None):
        self.diffusion_steps = int(diffusion_steps)
        self.max_keep = max_keep
        self.reset()

    def reset(self):
        T = self.diffusion_steps
        self.metric_per_t = {t: [] for t in range(1, T + 1)}
        self.avg_per_step, self.avg_sq_per_step, self.n_per_step = np.zeros(T), np.zeros(T), np.zeros(T)

    def _truncating(self):
        return self.max_keep is not None and self.diffusion_steps > self.max_keep

    def update(self, t, metric):
        if not np.isfinite(metric):
            return
        t = int(t)
        hist = self.metric_per_t[t]
        hist.append(metric)
        if self._truncating() and len(hist) > self.max_keep:
            del hist[:-self.max_keep]
        arr = np.asarray(hist, dtype=np.float64)
        self.avg_per_step[t - 1] = arr.mean()
        self.avg_sq_per_step[t - 1] = np.sqrt(np.mean(arr * arr))
        self.n_per_step[t - 1] += 1

    def update_multiple(self, ts, metrics):
        for t, m in zip(ts, metrics):
            self.update(t, m)

    def get_avg_in_range(self, t0, t1):
        return np.concatenate([self.metric_per_t[t] for t in range(t0, t1)]).mean()

    def __getitem__(self, t):
        return self.metric_per_t[t]


class UniformSampler:
    """``t, weights = sampler(batch_size, device)`` with t ~ U{1..T} (int64) and ``weights is None``."""

    def __init__(self, diffusion_steps):
        self.diffusion_steps = diffusion_steps

    def __call__(self, batch_size, device):
        return torch.randint(1, self.diffusion_steps + 1, (batch_size,), device=device), None


class ImportanceSampler(UniformSampler):
    """Samples t proportionally to the running RMS loss per timestep once every t has ``min_counts`` samples;
    uniform before that.  Returns float64 weights 1/(p_t * B), which make the batch loss a weighted sum
    (src/engine.py:274-275)."""

    def __init__(self, diffusion_steps, loss_per_t: StepwiseLog, min_counts=10):
        super().__init__(diffusion_steps)
        self.loss_per_t = loss_per_t
        self.min_counts = min_counts
        self._ready = False

    def is_ready(self):
        if not self._ready and bool((self.loss_per_t.n_per_step >= self.min_counts).all()):
            print("ImportanceSampler is warmed up now")
            self._ready = True
        return self._ready

    def __call__(self, batch_size, device):
        if not self.is_ready():
            return super().__call__(batch_size, device)
        p = self.loss_per_t.avg_sq_per_step + 1e-6
        p = p / p.sum()
        idx = np.random.choice(self.diffusion_steps, size=(batch_size,), p=p)
        t = torch.from_numpy(idx).long().to(device) + 1
        return t, torch.from_numpy(1 / (p[idx] * batch_size)).to(device)
