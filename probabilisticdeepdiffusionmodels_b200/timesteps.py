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


class DeviceStepwiseLog:
    """``StepwiseLog`` with its state in device tensors and batched, sync-free updates: the
    reference pulls the per-sample losses and timesteps to the host every training step (src/engine.py:267-271) and
    walks them in a Python loop.  Same statistics -- per-timestep mean, RMS and count over the retained history,
    including the reference's ``max_keep`` truncation rule -- computed with a handful of tensor ops, so the loss
    bookkeeping can live inside a captured CUDA graph.  ``avg_per_step`` / ``avg_sq_per_step`` / ``n_per_step`` are
    float64 tensors of length T (index = t - 1)."""

    def __init__(self, diffusion_steps, max_keep=None, device="cpu"):
        self.diffusion_steps = int(diffusion_steps)
        self.max_keep = max_keep
        self.device = torch.device(device)
        self.reset()

    def _truncating(self):
        return self.max_keep is not None and self.diffusion_steps > self.max_keep

    def reset(self):
        T, dev = self.diffusion_steps, self.device
        self.n_per_step = torch.zeros(T, dtype=torch.float64, device=dev)
        self.avg_per_step = torch.zeros(T, dtype=torch.float64, device=dev)
        self.avg_sq_per_step = torch.zeros(T, dtype=torch.float64, device=dev)
        if self._truncating():
            # ring of the last K entries per timestep; row T is a scratch row for dropped samples
            self.hist = torch.zeros((T + 1, self.max_keep), dtype=torch.float64, device=dev)
        else:
            self.s1 = torch.zeros(T, dtype=torch.float64, device=dev)
            self.s2 = torch.zeros(T, dtype=torch.float64, device=dev)

    def to(self, device):
        self.device = torch.device(device)
        for k, v in list(vars(self).items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(self.device))
        return self

    def get_avg_in_range(self, t0, t1):
        """mean of every retained entry with t0 <= t < t1 (src/modules/stepwise_log.py:33-34) -> 0-dim tensor"""
        if self._truncating():
            valid = torch.clamp(self.n_per_step, max=float(self.max_keep))[t0 - 1: t1 - 1]
            return (self.avg_per_step[t0 - 1: t1 - 1] * valid).sum() / valid.sum()
        return self.s1[t0 - 1: t1 - 1].sum() / self.n_per_step[t0 - 1: t1 - 1].sum()

    @torch.no_grad()
    def update_multiple(self, ts, metrics):
        """ts: int tensor [B] (1-indexed), metrics: float tensor [B]; processed in batch order like the reference."""
        T = self.diffusion_steps
        t0 = ts.to(device=self.device, dtype=torch.int64) - 1
        m = metrics.detach().to(device=self.device, dtype=torch.float64)
        ok = torch.isfinite(m)  # the reference skips non-finite metrics entirely
        okf = ok.to(torch.float64)
        m = torch.where(ok, m, torch.zeros_like(m))
        if not self._truncating():
            self.s1.index_add_(0, t0, m * okf)
            self.s2.index_add_(0, t0, m * m * okf)
            self.n_per_step.index_add_(0, t0, okf)
            n = self.n_per_step.clamp(min=1.0)
            seen = self.n_per_step > 0
            # (all state is updated IN PLACE: under CUDA-graph replay the tensors must keep their addresses)
            self.avg_per_step.copy_(torch.where(seen, self.s1 / n, self.avg_per_step))
            self.avg_sq_per_step.copy_(torch.where(seen, torch.sqrt(self.s2 / n), self.avg_sq_per_step))
            return
        K, B = self.max_keep, t0.shape[0]
        # rank of each sample among the earlier VALID samples with the same timestep (stable order = batch order);
        # invalid samples are routed to an out-of-range key so that they take no slot
        key = torch.where(ok, t0, torch.full_like(t0, T))
        order = torch.argsort(key, stable=True)
        ks = key[order]
        idx = torch.arange(B, device=self.device)
        is_start = torch.ones(B, dtype=torch.bool, device=self.device)
        is_start[1:] = ks[1:] != ks[:-1]
        start = torch.cummax(torch.where(is_start, idx, torch.zeros_like(idx)), 0).values
        rank = torch.empty_like(idx)
        rank[order] = idx - start
        # valid samples per timestep in this batch (index_add, not bincount: no host sync, CUDA-graph capturable)
        cnt = torch.zeros(T + 1, dtype=torch.int64, device=self.device).index_add_(0, key, torch.ones_like(key))[:T]
        # only the last K samples of a timestep can survive; earlier ones would be overwritten anyway
        keep = ok & (rank >= (cnt[t0.clamp(max=T - 1)] - K))
        slot = (self.n_per_step[t0.clamp(max=T - 1)].to(torch.int64) + rank) % K
        rows = torch.where(keep, t0, torch.full_like(t0, T))  # dropped samples -> scratch row
        self.hist.index_put_((rows, slot), m)
        self.n_per_step.add_(cnt.to(torch.float64))
        valid = torch.clamp(self.n_per_step, max=float(K))
        col = torch.arange(K, device=self.device)[None, :].to(torch.float64)
        mask = (col < valid[:, None]).to(torch.float64)  # a ring holds its first min(n, K) slots
        h = self.hist[:T] * mask
        seen = valid > 0
        v = valid.clamp(min=1.0)
        self.avg_per_step.copy_(torch.where(seen, h.sum(1) / v, self.avg_per_step))
        self.avg_sq_per_step.copy_(torch.where(seen, torch.sqrt((h * h).sum(1) / v), self.avg_sq_per_step))


class DeviceImportanceSampler:
    """``ImportanceSampler`` over a ``DeviceStepwiseLog`` without host round trips: the warm-up test, the probability
    vector, the draw (``torch.multinomial``) and the weights are tensor ops on the log's device.  Returns
    ``(t, weights, ready)`` with ``ready`` a 0-dim bool tensor; while it is False ``t`` is uniform and the caller must
    average instead of weighting (``Engine`` does: loss = where(ready, sum(w * L), mean(L)))."""

    def __init__(self, diffusion_steps, loss_per_t: DeviceStepwiseLog, min_counts=10):
        self.diffusion_steps = diffusion_steps
        self.loss_per_t = loss_per_t
        self.min_counts = min_counts
        self._ready = torch.zeros((), dtype=torch.bool, device=loss_per_t.device)

    def probabilities(self):
        p = self.loss_per_t.avg_sq_per_step + 1e-6
        return p / p.sum()

    def __call__(self, batch_size, device=None, generator=None):
        log = self.loss_per_t
        if self._ready.device != log.device:
            self._ready = self._ready.to(log.device)
        self._ready.logical_or_((log.n_per_step >= self.min_counts).all())  # latches like the reference (in place)
        p = self.probabilities()
        idx = torch.multinomial(p.to(torch.float32), batch_size, replacement=True, generator=generator)
        t_imp = idx + 1
        t_uni = torch.randint(1, self.diffusion_steps + 1, (batch_size,), device=log.device, generator=generator)
        t = torch.where(self._ready, t_imp, t_uni)
        weights = 1.0 / (p[t - 1] * batch_size)
        return t, weights, self._ready
