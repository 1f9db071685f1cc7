"""``Engine``: the reference's LightningModule surface (src/engine.py:79-657) on top of the sm_100a kernels.

Same constructor arguments, attributes (the 14 fp32 coefficient tables, ``model``, ``ema``, samplers, logs) and
methods as the reference, so ``scripts/train.py`` / ``sample.py`` / ``eval.py``, the visualisation callback and
the FID tooling drive it unchanged.  What changed underneath:

* the coefficient tables also live on the device (``DeviceTables``); ``q_sample``, the squared-error loss, the
  reverse ``p_sample`` step and the variational-bound terms are single fused kernels that gather their
  per-sample / per-step coefficients on the device (no CPU gather + H2D copy per call);
* the reverse chain replays ONE CUDA graph per step (UNet forward + p_sample + step bookkeeping) whose step
  index lives in device memory; the per-step noise still comes from ``torch.randn(generator=...)`` so a seed
  produces the same stream as the reference on the same device;
* ``capture_train_step`` runs forward + backward + Adam (+EMA) as one CUDA graph.

Extensions (off by default = reference behaviour): ``learn_sigma`` (variance channels + L_hybrid, SURVEY.md
Appendix C) and ``log_loss_per_t=False`` (skips the reference's per-step device->host sync in ``get_loss``).
"""
from contextlib import contextmanager
from typing import List, Tuple

import numpy as np
import os

import torch

from . import functional as F
from . import ops
from .mathutils import get_generator_if_specified, mean_flat
from .modules import get_model
from .schedules import TABLE_NAMES, get_betas, make_tables
from .timesteps import (DeviceImportanceSampler, DeviceStepwiseLog, ImportanceSampler, StepwiseLog,
                        UniformSampler)
from .weight_average import Ema

try:  # the reference subclasses pl.LightningModule; fall back to a minimal stand-in when Lightning is absent
    import pytorch_lightning as pl

    _Base = pl.LightningModule
except Exception:  # pragma: no cover - depends on the environment

    class _Base(torch.nn.Module):
        """Just enough of LightningModule for Engine to be driven by a hand-written loop."""

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, name, value, **k):
            self.__dict__.setdefault("logged", {})[name] = value

        def optimizer_step(self, epoch=None, batch_idx=None, optimizer=None, optimizer_idx=None,
                           optimizer_closure=None, **k):
            if optimizer is not None:
                optimizer.step(closure=optimizer_closure) if optimizer_closure is not None else optimizer.step()


LN2 = float(np.log(2.0))


# src/datasets/data.py:24-28
NORMALIZATIONS = {"cifar": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), "mnist": ((0.5,), (0.5,)),
                  "oneone": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))}


class Engine(_Base):
    def __init__(self, model_config, optimizer_config, diffusion_steps=1000, beta_start=None, beta_end=None,
                 mode="linear", max_beta=0.999, sigma_mode="beta", resolution=32, clip_while_generating=False,
                 sampling="uniform", ema=None, scheduler_name=None, scheduler_kwargs=None, learn_sigma=False,
                 log_loss_per_t=True):
        super().__init__()
        self.save_hyperparameters()
        self.clip_while_generating = clip_while_generating
        self.learn_sigma = bool(learn_sigma)
        self.log_loss_per_t = log_loss_per_t
        cfg = dict(model_config)
        if self.learn_sigma:
            cfg["learn_sigma"] = True
        self.model = get_model(resolution, cfg)
        if ema is not None:
            self.ema = Ema(self.model, decay=ema)
            self.ema.set(self.model)
        else:
            self.ema = None
        self.optimizer_config = optimizer_config
        self.diffusion_steps = diffusion_steps
        self.resolution = resolution
        if sigma_mode not in ("beta", "beta_tilde"):
            raise ValueError(f"Wrong sigma mode: {sigma_mode}")
        self.sigma_mode = sigma_mode

        # fp32 CPU tables under the reference's attribute names (src/engine.py:121-150)
        self._tables = make_tables(get_betas(beta_start, beta_end, diffusion_steps, mode, max_beta=max_beta))
        for name in TABLE_NAMES:
            setattr(self, name, self._tables[name])
        self._dev_tables = {}

        if log_loss_per_t == "device":  # extension: statistics live on the device, no per-step host sync
            self.loss_per_t = DeviceStepwiseLog(diffusion_steps, 10)
            self.loss_per_t_epoch = DeviceStepwiseLog(diffusion_steps)
        else:
            self.loss_per_t = StepwiseLog(diffusion_steps, 10)
            self.loss_per_t_epoch = StepwiseLog(diffusion_steps)
        self.sampling_name = sampling
        if sampling == "uniform":
            self.sampler = UniformSampler(diffusion_steps=diffusion_steps)
        elif sampling == "importance" and log_loss_per_t == "device":
            self.sampler = DeviceImportanceSampler(diffusion_steps, self.loss_per_t, min_counts=10)
        elif sampling == "importance":
            self.sampler = ImportanceSampler(diffusion_steps=diffusion_steps, loss_per_t=self.loss_per_t, min_counts=10)
        else:
            raise ValueError(f'Unknown sampling option: "{sampling}"')
        self.val_sampler = UniformSampler(diffusion_steps=diffusion_steps)
        self.scheduler_name = scheduler_name
        self.scheduler_kwargs = scheduler_kwargs
        self._chain_graphs = {}
        self._train_graph = None

    # ------------------------------------------------------------------ plumbing
    def tabs(self, device=None) -> F.DeviceTables:
        device = torch.device(device if device is not None else self.device)
        key = (device.type, device.index)
        if key not in self._dev_tables:
            self._dev_tables[key] = F.DeviceTables(self._tables, device)
        return self._dev_tables[key]

    def _gather(self, name, t, like):
        """``table[t-1].view(-1,1,1,1)`` on ``like``'s device (the reference's per-call CPU gather, on device)."""
        tab = self.tabs(like.device).t[name] if like.is_cuda else self._tables[name]
        idx = torch.as_tensor(t, device=tab.device).long().reshape(-1) - 1
        return tab[idx].view(-1, 1, 1, 1)

    @contextmanager
    def ema_on(self):
        """src/engine.py:171-182"""
        if self.ema is None:
            yield
        else:
            try:
                self.original_model = self.model
                self.model = self.ema.module
                yield
            finally:
                self.model = self.original_model
                self.original_model = None

    def on_epoch_end(self) -> None:
        """src/engine.py:184-215 without the matplotlib / wandb plots (out of scope)."""
        for i in range(4):
            lo, hi = max(1, int(i * self.diffusion_steps / 4)), int((i + 1) * self.diffusion_steps / 4)
            try:
                self.log(f"loss_q{i + 1}", self.loss_per_t_epoch.get_avg_in_range(lo, hi), on_step=False,
                         on_epoch=True, prog_bar=False)
            except ValueError:
                pass
        self.loss_per_t_epoch.reset()

    def optimizer_step(self, *args, **kwargs):
        """src/engine.py:217-224"""
        super().optimizer_step(*args, **kwargs)
        if self.ema:
            self.ema.update(self.model)

    def configure_optimizers(self):
        """src/engine.py:238-248"""
        optimizer = torch.optim.Adam(self.parameters(), **self.optimizer_config)
        if self.scheduler_name:
            scheduler_class = getattr(torch.optim.lr_scheduler, self.scheduler_name)
            return {"optimizer": optimizer, "lr_scheduler": scheduler_class(optimizer, **self.scheduler_kwargs)}
        return optimizer

    # ------------------------------------------------------------------ training-side diffusion math
    def q_mean_std(self, x, t):
        """src/engine.py:251-257"""
        return x * self._gather("alphas_hat_sqrt", t, x), self._gather("one_min_alphas_hat_sqrt", t, x)

    def get_q_t(self, x, noise, t):
        """x_t = sqrt(abar_t) x + sqrt(1-abar_t) noise (src/engine.py:259-261): one fused kernel."""
        if not x.is_cuda:
            raise RuntimeError("Engine.get_q_t needs CUDA tensors (sm_100a); there is no CPU fallback")
        x, noise = x.float().contiguous(), noise.float().contiguous()
        if isinstance(t, torch.Tensor) and t.numel() == 1:
            t = int(t.item()) if not torch.cuda.is_current_stream_capturing() else t
        return F.q_sample(x, noise, t, self.tabs(x.device))

    def per_sample_loss(self, model_out, target_noise, x, x_t, t):
        """[B] losses: L_simple, or L_simple + (T/1000) L_vlb with learned variance (Appendix C.6)."""
        if self.learn_sigma:
            per, _ = ops.HybridLoss.apply(model_out, target_noise.contiguous(), x.contiguous(), x_t.contiguous(), t,
                                          self.tabs(x.device), self.diffusion_steps / 1000.0)
            return per
        return torch.ops.pddm.simple_loss(model_out, target_noise)

    def get_loss(self, predicted_noise, target_noise, x, x_t, t, weights=None, update_loss_log=True, ready=None):
        """src/engine.py:263-277.  ``ready`` (0-dim bool tensor, device importance sampler): weight the batch only once
        the sampler is warmed up, average before -- decided on the device."""
        loss = self.per_sample_loss(predicted_noise, target_noise, x, x_t, t)
        if update_loss_log and self.log_loss_per_t == "device":
            for log in (self.loss_per_t, self.loss_per_t_epoch):
                if log.device != loss.device:
                    log.to(loss.device)
                log.update_multiple(t, loss.detach())
        elif update_loss_log and self.log_loss_per_t:
            losses = loss.detach().cpu().numpy().tolist()  # the reference's per-step host sync
            ts = t.detach().cpu().numpy().tolist()
            self.loss_per_t.update_multiple(ts, losses)
            self.loss_per_t_epoch.update_multiple(ts, losses)
        if ready is not None:
            return torch.where(ready, torch.sum(weights * loss), torch.mean(loss).to(weights.dtype))
        if weights is not None:
            return torch.sum(weights * loss)
        return torch.mean(loss)

    def _draw_timesteps(self, batch_size):
        """-> (t, weights | None, ready | None) from either kind of sampler"""
        if isinstance(self.sampler, DeviceImportanceSampler):
            if self.loss_per_t.device != self.device:
                self.loss_per_t.to(self.device)
            return self.sampler(batch_size)
        t, weights = self.sampler(batch_size, self.device)
        return t, weights, None

    def training_step(self, batch, batch_idx):  # pylint: disable=unused-argument
        """src/engine.py:279-307"""
        x, y = batch
        t, weights, ready = self._draw_timesteps(x.shape[0])
        noise = torch.randn_like(x)
        x_t = self.get_q_t(x, noise, t)
        predicted_noise = self.model(x_t, t)
        loss = self.get_loss(predicted_noise, noise, x, x_t, weights=weights, t=t, update_loss_log=True, ready=ready)
        total_norm = self.compute_grad_norm(self.model.parameters())
        self.log("loss", loss, on_step=False, on_epoch=True, prog_bar=True)
        self.log("total_grad_norm_L2", total_norm, on_step=True, on_epoch=False, prog_bar=False)
        return loss

    def validation_step(self, batch, batch_idx):  # pylint: disable=unused-argument
        """src/engine.py:309-330"""
        x, y = batch
        t, weights = self.val_sampler(x.shape[0], self.device)
        noise = torch.randn_like(x)
        x_t = self.get_q_t(x, noise, t)
        loss = self.get_loss(self.model(x_t, t), noise, x, x_t, weights=weights, t=t, update_loss_log=False)
        if self.ema is not None:
            with self.ema_on():
                predicted_noise = self.model(x_t, t)
            loss_ema = self.get_loss(predicted_noise, noise, x, x_t, weights=weights, t=t, update_loss_log=False)
            self.log("val_loss_no_ema", loss, on_step=False, on_epoch=True, prog_bar=False)
            self.log("val_loss", loss_ema, on_step=False, on_epoch=True, prog_bar=True)
        else:
            self.log("val_loss", loss, on_step=False, on_epoch=True, prog_bar=True)

    def compute_grad_norm(self, parameters, norm_type=2):
        """src/engine.py:332-346 (norm of the previous step's gradients), as two multi-tensor launches."""
        if isinstance(parameters, torch.Tensor):
            parameters = [parameters]
        grads = [p.grad.detach() for p in parameters if p.grad is not None]
        if not grads:
            return torch.tensor(0.0)
        return torch.norm(torch.stack(torch._foreach_norm(grads, float(norm_type))), float(norm_type))

    # ------------------------------------------------------------------ reverse-process pieces (API parity)
    def get_sigma(self, t):
        """src/engine.py:354-361 (``t`` is a 0-based index here, like the reference's call ``get_sigma(t_step-1)``)."""
        if self.sigma_mode == "beta":
            return torch.sqrt(self.betas[t])
        elif self.sigma_mode == "beta_tilde":
            return torch.sqrt(self.posterior_variance[t])
        raise ValueError(f"Wrong sigma mode: {self.sigma_mode}")

    def xstart_from_epsilon(self, x_t, t, epsilon, clip=False):
        """src/engine.py:363-368"""
        x = self._gather("sqrt_recip_alphas_cumprod", t, x_t) * x_t \
            - self._gather("sqrt_recipm1_alphas_cumprod", t, x_t) * epsilon
        return x.clamp(-1, 1) if clip else x

    def q_posterior(self, t, x0, x_t):
        """src/engine.py:477-490"""
        mean_t = x0 * self._gather("posterior_mean_coef1", t, x_t) + x_t * self._gather("posterior_mean_coef2", t, x_t)
        return mean_t, self._gather("posterior_variance", t, x_t)

    def model_mean_through_start(self, x_t, t, epsilon, clip=False):
        """src/engine.py:370-373"""
        return self.q_posterior(t, self.xstart_from_epsilon(x_t, t, epsilon, clip=clip), x_t)[0]

    def model_mean_from_epsilon(self, x_t, t, epsilon, clip=False):
        """src/engine.py:375-381"""
        if clip:
            return self.model_mean_through_start(x_t, t, epsilon, clip=True)
        return (x_t - epsilon * self._gather("denoising_coef", t, x_t)) / self._gather("alphas_sqrt", t, x_t)

    def model_mean_std(self, x_t, t, t_step, clip=False):
        """src/engine.py:348-352"""
        out = self.model(x_t, t)
        eps = out[:, : x_t.shape[1]] if self.learn_sigma else out
        if x_t.is_cuda:
            mean = F.p_sample_step(x_t.float().contiguous(), out.contiguous(), None, t_step, self.tabs(x_t.device),
                                   clip, self.sigma_mode)
        else:
            mean = self.model_mean_from_epsilon(x_t, t_step, eps, clip=clip)
        return eps, mean, self.get_sigma(t_step - 1).to(self.device)

    def denoising_step(self, x_t, t_step, mean_only=False, generator=None):
        """x_{t-1} = mean - sigma * z  (src/engine.py:385-397): UNet forward + one fused kernel."""
        t = t_step * torch.ones(x_t.shape[0], device=self.device)
        out = self.model(x_t, t)
        z = None
        if not mean_only and t_step > 1:
            z = torch.randn(x_t.shape, generator=generator, device=self.device, dtype=x_t.dtype)
        return F.p_sample_step(x_t.float().contiguous(), out.contiguous(), z, t_step, self.tabs(x_t.device),
                               self.clip_while_generating, "learned" if self.learn_sigma else self.sigma_mode)

    def sample_from_step(self, x_t, t_start, mean_only=False, generator=None, use_graph=True):
        """src/engine.py:399-403.  With ``use_graph`` the chain replays one captured step (same arithmetic)."""
        if use_graph and x_t.is_cuda and not torch.is_grad_enabled():
            return self._chain(x_t, t_start, (), mean_only, generator)[0]
        for t in range(t_start, 0, -1):
            x_t = self.denoising_step(x_t, t, mean_only=mean_only, generator=generator)
        return x_t

    # ------------------------------------------------------------------ graph-replayed reverse chain
    def _weights_stamp(self):
        """Changes whenever the current model's weights may have changed (in-place torch ops bump ``_version``;
        kernels that write parameters through raw pointers bump the global epochs)."""
        from . import plan as _plan
        return (tuple(p._version for p in self.model.parameters()), ops._EPOCH[0], _plan._EPOCH[0])

    def _chain_graph(self, shape, dtype, device, mean_only):
        """One captured reverse step.  The graph must never read stale weights (optimizer steps, EMA updates and
        ``load_state_dict`` between two sampling calls are the normal case): with a plan, the graph reads its bf16
        operands from the plan's persistent arena, which ``_chain`` refreshes (one launch) before replaying; on the
        op-by-op path the graph hard-codes the addresses of cached packs, so it is re-captured when the stamp moves."""
        from . import plan as _plan
        key = (tuple(shape), str(device), bool(mean_only), id(self.model), self.clip_while_generating,
               self.sigma_mode, self.learn_sigma)
        g = self._chain_graphs.get(key)
        probe = torch.empty(shape, dtype=torch.float32, device=device)
        pl = _plan.plan_for(self.model, probe)
        stamp = None if pl is not None else self._weights_stamp()
        if g is not None and g["plan"] is pl and (pl is not None or g["stamp"] == stamp):
            return g
        st = {"x": torch.zeros(shape, dtype=torch.float32, device=device),
              "z": None if mean_only else torch.zeros(shape, dtype=torch.float32, device=device),
              "t_vec": torch.ones(shape[0], dtype=torch.float32, device=device),
              "t_dev": torch.ones(1, dtype=torch.int32, device=device)}
        sigma = "learned" if self.learn_sigma else self.sigma_mode
        tabs = self.tabs(device)

        def step():
            out = self.model(st["x"], st["t_vec"])
            F.p_sample_step(st["x"], out, st["z"], -1, tabs, self.clip_while_generating, sigma, out=st["x"],
                            t_dev=st["t_dev"])
            F.step_advance(st["t_dev"], st["t_vec"])

        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: weight packs cached, attributes set, allocator primed
                st["t_dev"].fill_(1)
                st["t_vec"].fill_(1.0)
                step()
        torch.cuda.current_stream(device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        st["t_dev"].fill_(1)
        st["t_vec"].fill_(1.0)
        with torch.cuda.graph(graph):
            step()
        st["graph"] = graph
        st["plan"], st["stamp"] = pl, stamp
        self._chain_graphs[key] = st
        return st

    @torch.no_grad()
    def _chain(self, x_t, t_start, steps_to_return, mean_only, generator, return_stds=False, fixed_noise=None):
        was_training = self.model.training
        self.model.eval()
        snaps, stds = [], []
        with ops.frozen_weights():
            st = self._chain_graph(x_t.shape, x_t.dtype, x_t.device, mean_only)
            if st["plan"] is not None:
                st["plan"].refresh_packs()  # the replayed graph reads the arena: make it current (one launch)
            st["x"].copy_(x_t)
            st["t_dev"].fill_(int(t_start))
            st["t_vec"].fill_(float(t_start))
            if return_stds:
                stds.append(torch.std(x_t).detach().cpu().item())
            for t in range(t_start, 0, -1):
                if not mean_only and t > 1:
                    if fixed_noise is not None:  # parity harness: z for step t is fixed_noise[t_start - t]
                        st["z"].copy_(fixed_noise[t_start - t])
                    elif generator is not None:
                        st["z"].normal_(generator=generator)
                    else:
                        st["z"].normal_()
                st["graph"].replay()
                if t in steps_to_return:
                    snaps.append(st["x"].clone())
                if return_stds:
                    stds.append(torch.std(st["x"]).detach().cpu().item())
            out = st["x"].clone()
        self.model.train(was_training)
        return out, snaps, stds

    # ------------------------------------------------------------------ NLL evaluation (src/engine.py:407-506)
    def test_step(self, batch, batch_idx):
        x, _ = batch
        with self.ema_on():
            nll = self.calculate_likelihood(x)
        for k_out, k_in in (("test_L_0", "L_0"), ("test_L_intermediate", "L_intermediate"), ("test_L_T", "L_T"),
                            ("test_nll", "nll"), ("test_mse", "MSE")):
            self.log(k_out, nll[k_in])

    def calculate_likelihood(self, x, t_batch=1, shard=None, group=None):
        """Eq. (5) of DDPM in bits/dim (src/engine.py:417-435).  ``t_batch`` > 1 (extension) evaluates that
        many timesteps per UNet forward by folding t into the batch dimension (the T-1 terms are independent); the
        default reproduces the reference's one-forward-per-t loop and its noise draw order.

        ``shard=(rank, world)`` (extension): the T-1 intermediate terms are dealt round-robin to the
        ranks (``parallel.shard_timesteps``), rank 0 adds L_0 and L_T, and ONE all-reduce of a [3B + 2] vector hands
        every rank the complete result -- the only collective of the evaluation.  ``x`` is the same batch on every
        rank (shard the batch with ``parallel.shard_batch`` first to split both ways)."""
        if shard is None:
            L_0 = self._calculate_L_0(x)
            L_intermediate_list, MSE_list = self._calculate_L_intermediate(x, t_batch)
            L_T = self._calculate_L_T(x)
            L_intermediate = torch.sum(torch.stack(L_intermediate_list), dim=0)
            return {"MSE": torch.mean(torch.stack(MSE_list)), "MSE_list": MSE_list, "L_0": torch.mean(L_0, dim=0),
                    "L_intermediate": L_intermediate, "L_T": torch.mean(L_T, dim=0),
                    "nll": torch.mean(L_0 + L_intermediate + L_T, dim=0), "L_intermediate_list": L_intermediate_list}
        from . import parallel
        rank, world = shard
        steps = parallel.shard_timesteps(self.diffusion_steps, rank, world)
        L_list, MSE_list = self._calculate_L_intermediate(x, t_batch, steps=steps)
        B = x.shape[0]
        zero = torch.zeros(B, dtype=torch.float32, device=self.device)
        parts = {"L_int": torch.sum(torch.stack(L_list), dim=0) if L_list else zero,
                 "L_0": self._calculate_L_0(x) if rank == 0 else zero,
                 "L_T": self._calculate_L_T(x) if rank == 0 else zero,
                 "mse_sum": sum((m.double().sum() for m in MSE_list), torch.zeros((), dtype=torch.float64,
                                                                                 device=self.device)),
                 "mse_count": float(sum(m.numel() for m in MSE_list))}
        tot = parallel.all_reduce_nll(parts, group=group)
        return {"MSE": (tot["mse_sum"] / max(tot["mse_count"], 1.0)).float(), "MSE_list": MSE_list,
                "L_0": torch.mean(tot["L_0"], dim=0), "L_intermediate": tot["L_int"], "L_T": torch.mean(tot["L_T"], dim=0),
                "nll": torch.mean(tot["L_0"] + tot["L_int"] + tot["L_T"], dim=0), "L_intermediate_list": L_list}

    def _vlb(self, x0, x_t, model_out, t, mode):
        out, _ = F.vlb_terms(x0.float().contiguous(), None if x_t is None else x_t.contiguous(),
                             None if model_out is None else model_out.contiguous(), t, self.tabs(x0.device), mode,
                             self.sigma_mode)
        return out

    def _calculate_L_T(self, x):
        """KL(q(x_T|x_0) || N(0, I)) (src/engine.py:437-444)"""
        return self._vlb(x, None, None, None, 2)

    def _calculate_L_intermediate(self, x0, t_batch=1, steps=None) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """sum_t KL(q(x_{t-1}|x_t,x_0) || p(x_{t-1}|x_t)), fixed variance (src/engine.py:446-475); ``steps``: the
        subset of t in [2, T] this rank evaluates (default: all)."""
        L_list, MSE_list = [], []
        ones = torch.ones(x0.shape[0], dtype=torch.int64, device=self.device)
        steps = list(range(2, self.diffusion_steps + 1)) if steps is None else list(steps)
        if t_batch > 1:
            B = x0.shape[0]
            for i in range(0, len(steps), t_batch):
                chunk = steps[i: i + t_batch]
                t = torch.tensor(chunk, dtype=torch.int64, device=self.device).repeat_interleave(B)
                xr = x0.repeat(len(chunk), 1, 1, 1)
                noise = torch.randn_like(xr)
                x_t = self.get_q_t(xr, noise, t)
                out = self.model(x_t, t)
                eps = out[:, : x0.shape[1]].contiguous() if self.learn_sigma else out
                kl = self._vlb(xr, x_t, eps, t, 0).view(len(chunk), B)
                mse = torch.pow(eps - noise, 2).view(len(chunk), B, *x0.shape[1:])
                L_list.extend(kl.unbind(0))
                MSE_list.extend(mse.unbind(0))
            return L_list, MSE_list
        for t_step in steps:
            t = ones * t_step
            noise = torch.randn_like(x0)
            x_t = self.get_q_t(x0, noise, t)
            out = self.model(x_t, t)
            eps = out[:, : x0.shape[1]].contiguous() if self.learn_sigma else out
            L_list.append(self._vlb(x0, x_t, eps, t, 0))
            MSE_list.append(torch.pow(eps - noise, 2))
        return L_list, MSE_list

    def _calculate_L_0(self, x):
        """-log p(x_0 | x_1) with the discretised Gaussian decoder (src/engine.py:492-506)"""
        t = torch.ones(x.shape[0], dtype=torch.int64, device=self.device)
        noise = torch.randn_like(x)
        x_t = self.get_q_t(x, noise, t)
        out = self.model(x_t, t)
        eps = out[:, : x.shape[1]].contiguous() if self.learn_sigma else out
        return self._vlb(x, x_t, eps, t, 0)

    # ------------------------------------------------------------------ generation endpoints (src/engine.py:508-657)
    @torch.no_grad()
    def sample_and_return_steps(self, x_t, t_start=None, steps_to_return=(1,), mean_only=False, generator=None,
                                seed=None, return_stds=False, fixed_noise=None):
        """Returns shape [B, STEPS, C, W, H] (CPU tensor, like the reference).  ``fixed_noise`` ([steps, B, C, H, W],
        extension) injects the per-step z instead of drawing it, so CPU oracle and GPU see identical noise."""
        if t_start is None:
            t_start = self.diffusion_steps
        if generator is None:
            generator = get_generator_if_specified(seed, device=self.device)
        assert all(t < t_start for t in steps_to_return)
        self.eval()
        _, snaps, stds = self._chain(x_t.to(self.device).float(), t_start, tuple(steps_to_return), mean_only,
                                     generator, return_stds, fixed_noise)
        output = torch.zeros((x_t.shape[0], len(steps_to_return)) + tuple(x_t.shape[1:]))
        for i, s in enumerate(snaps):
            output[:, i] = s.cpu()
        return (output, stds) if return_stds else output

    @torch.no_grad()
    def generate_images(self, n=1, minibatch=4, mean_only=False, seed=None):
        self.eval()
        generator = get_generator_if_specified(seed, device=self.device)
        images = []
        for _ in range(int(np.ceil(n / minibatch))):
            x_t = torch.randn((minibatch, self.model.in_channels, self.resolution, self.resolution),
                              generator=generator, device=self.device)
            x_t = self.sample_from_step(x_t, self.diffusion_steps, mean_only=mean_only, generator=generator)
            images.append(x_t.detach().cpu().numpy())
        return np.concatenate(images, axis=0)

    @torch.no_grad()
    def generate_images_uint8(self, n=1, minibatch=4, mean_only=False, seed=None, normalize=None):
        """``generate_images`` followed by the reference's host post-processing (src/modules/fid_score.py:15-27:
        ``unnormalize(img, normalize, clip=True)`` per image, then the writer's ``(255 * A).astype(uint8)``) done on
        the device: each mini-batch is clamped, un-normalised and quantised by one kernel and leaves as ONE uint8 NHWC
        copy (a quarter of the fp32 bytes).  Same RNG stream as ``generate_images``.  ``normalize``: None (the
        reference's call), a (mean, std) pair, or a key of NORMALIZATIONS (src/datasets/data.py:24-28).
        Returns uint8 [N, H, W, C]."""
        self.eval()
        mean = std = None
        if normalize is not None:
            mean, std = NORMALIZATIONS[normalize] if isinstance(normalize, str) else normalize
        generator = get_generator_if_specified(seed, device=self.device)
        images = []
        for _ in range(int(np.ceil(n / minibatch))):
            x_t = torch.randn((minibatch, self.model.in_channels, self.resolution, self.resolution),
                              generator=generator, device=self.device)
            x_t = self.sample_from_step(x_t, self.diffusion_steps, mean_only=mean_only, generator=generator)
            images.append(F.images_to_uint8(x_t.float().contiguous(), mean, std).cpu().numpy())
        return np.concatenate(images, axis=0)

    @torch.no_grad()
    def generate_images_grid(self, steps_to_return, n=1, minibatch=4, mean_only=False, seed=None):
        self.eval()
        generator = get_generator_if_specified(seed, device=self.device)
        starting_noise, images = [], []
        for _ in range(int(np.ceil(n / minibatch))):
            x_t = torch.randn((n, self.model.in_channels, self.resolution, self.resolution), generator=generator,
                              device=self.device)
            starting_noise.append(x_t.detach().cpu().numpy())
            steps = self.sample_and_return_steps(x_t, self.diffusion_steps, steps_to_return=steps_to_return,
                                                 mean_only=mean_only, generator=generator)
            images.append(steps.detach().cpu().numpy())
        return np.concatenate(starting_noise, axis=0), np.concatenate(images, axis=0)

    @torch.no_grad()
    def get_noised_representation(self, x0, t=None, seed=None, generator=None):
        if t is None:
            t = self.diffusion_steps
        if generator is None:
            generator = get_generator_if_specified(seed, device=self.device)
        x0 = x0.to(self.device)
        noise = torch.randn(x0.shape, generator=generator, device=self.device, dtype=x0.dtype)
        return self.get_q_t(x0, noise, t)

    @torch.no_grad()
    def diffuse_and_reconstruct(self, x0, t=None, seed=None):
        self.eval()
        if t is None:
            t = self.diffusion_steps
        generator = get_generator_if_specified(seed, device=self.device)
        x_t = self.get_noised_representation(x0, t, generator=generator)
        return self.sample_from_step(x_t.detach().clone(), t, generator=generator), x_t

    @torch.no_grad()
    def diffuse_and_reconstruct_grid(self, x0, t_start=None, steps_to_return=(1,), seed=None, mean_only=False,
                                     return_stds=False):
        self.eval()
        if t_start is None:
            t_start = self.diffusion_steps
        generator = get_generator_if_specified(seed, device=self.device)
        x0 = x0.to(self.device)
        noise = torch.randn(x0.shape, generator=generator, device=self.device, dtype=x0.dtype)
        x_t = self.get_q_t(x0, noise, t_start)
        return (self.sample_and_return_steps(x_t.detach().clone(), t_start, steps_to_return, generator=generator,
                                             mean_only=mean_only, return_stds=return_stds), x_t)

    # ------------------------------------------------------------------ fused training step (forward+backward+Adam)
    def loss_on(self, x, t, noise, weights=None):
        """training_step's math with injected (t, noise): q_sample -> UNet -> per-sample loss -> reduction."""
        x_t = self.get_q_t(x, noise, t)
        per = self.per_sample_loss(self.model(x_t, t), noise, x, x_t, t)
        return (torch.sum(weights * per) if weights is not None else torch.mean(per)), per

    def loss_and_grad(self, model_out, noise, x, x_t, t, gscale):
        """Per-sample training loss (L_simple, or L_hybrid with learned variance -- SURVEY.md App. C.6) and the gradient
        of ``sum_b gscale[b] * loss[b]`` w.r.t. the model output, from the fused loss kernels (no autograd)."""
        model_out, noise = model_out.contiguous(), noise.contiguous()
        if self.learn_sigma:
            w = self.diffusion_steps / 1000.0
            per_simple, _ = F.sq_err(model_out, noise)
            vb, gv = F.vlb_terms(x.contiguous(), x_t.contiguous(), model_out, t, self.tabs(x.device), mode=1,
                                 want_grad_v=True)
            _, dout = F.sq_err(model_out, noise, gscale, want_grad=True, grad_v_unit=gv, v_scale=w)
            return per_simple + w * vb, dout
        per, dout = F.sq_err(model_out, noise, gscale, want_grad=True)
        return per, dout

    def capture_train_step(self, batch_shape, optimizer=None, grad_hook=None, overlap_wgrad=True):
        """Capture one optimisation step (t ~ U{1..T}, eps ~ N(0,1) drawn inside the graph, loss, backward,
        optional ``grad_hook`` (e.g. the data-parallel all-reduce), Adam, EMA) as a CUDA graph.
        Returns ``step(x) -> loss`` replaying it on a static input buffer."""
        dev = self.device
        fused_ema = False
        if optimizer is None and os.environ.get("PDDM_TORCH_ADAM") == "1":  # A/B switch: torch's fused Adam
            optimizer = torch.optim.Adam(self.model.parameters(), capturable=True, fused=True, **self.optimizer_config)
        if optimizer is None:
            # Adam and (when configured) the EMA of the weights in one launch over the parameter list
            from .optim import FusedAdam
            fused_ema = self.ema is not None and not any(True for _ in self.model.buffers())
            optimizer = FusedAdam(self.model.parameters(), **self.optimizer_config,
                                  ema_params=self.ema.module.parameters() if fused_ema else None,
                                  ema_decay=self.ema.decay if fused_ema else None)
        st = {"x": torch.zeros(batch_shape, dtype=torch.float32, device=dev), "opt": optimizer}
        # Learning rate: the captured Adam launch reads it from device memory, so the reference's LR scheduler
        # (src/engine.py:238-248; Lightning steps it once per epoch) keeps working without re-capturing:
        # ``step.scheduler_step()`` advances the host-side scheduler and refreshes the device scalars.
        scheduler = None
        if hasattr(optimizer, "flush_tables"):  # FusedAdam
            for g in optimizer.param_groups:
                g["lr_dev"] = torch.full((1,), float(g["lr"]), dtype=torch.float32, device=dev)
            if self.scheduler_name:
                scheduler = getattr(torch.optim.lr_scheduler, self.scheduler_name)(optimizer, **self.scheduler_kwargs)
        st["scheduler"] = scheduler
        params = [p for p in self.model.parameters() if p.requires_grad]
        device_log = self.log_loss_per_t == "device"
        if self.sampling_name == "importance" and not device_log:
            raise ValueError(
                'capture_train_step draws timesteps inside the CUDA graph: sampling="importance" needs '
                'Engine(log_loss_per_t="device") (the host-side loss log cannot be updated from a replayed graph)')
        from . import _lib, plan as _plan
        plan = _plan.plan_for(self.model, st["x"]) if all(p.requires_grad for p in self.model.parameters()) else None
        st["plan"] = plan
        if plan is not None:
            plan.comm = None
            if grad_hook is not None and hasattr(grad_hook, "attach"):
                grad_hook.attach(plan)  # overlapped all-reduce: the backward pass reports finished arena ranges
        B = batch_shape[0]

        def draw():
            noise = torch.randn_like(st["x"])
            if device_log:  # timestep draw (uniform or importance) inside the graph
                t, weights, ready = self._draw_timesteps(B)
            else:
                t, weights, ready = torch.randint(1, self.diffusion_steps + 1, (B,), device=dev), None, None
            return noise, t, weights, ready

        def log_losses(t, per):
            if device_log:
                for log in (self.loss_per_t, self.loss_per_t_epoch):
                    if log.device != per.device:
                        log.to(per.device)
                    log.update_multiple(t, per.detach())

        def reduce_loss(per, weights, ready):
            if weights is None:
                return torch.mean(per)
            if ready is None:
                return torch.sum(weights * per)
            return torch.where(ready, torch.sum(weights * per), torch.mean(per).to(weights.dtype))

        if plan is not None:
            # ---- hand-scheduled step (plan.py): no autograd, gradients born in one flat arena
            # (kept in `st`: the graph reads it on every replay, long after this function's locals are gone)
            uniform_scale = st["uniform_scale"] = torch.full((B,), 1.0 / B, dtype=torch.float32, device=dev)

            def body():
                plan.repack()  # one launch refreshes every bf16 GEMM operand from the fp32 masters
                noise, t, weights, ready = draw()
                x_t = self.get_q_t(st["x"], noise, t)
                out, S = plan.forward(x_t, t, save=True)
                if weights is None:
                    gscale = uniform_scale
                elif ready is None:
                    gscale = weights.float()
                else:
                    gscale = torch.where(ready, weights.float(), uniform_scale)
                per, dout = self.loss_and_grad(out, noise, st["x"], x_t, t, gscale)
                log_losses(t, per)
                loss = reduce_loss(per, weights, ready)
                plan.backward(S, dout, overlap=overlap_wgrad)
                plan.assign_grads()
                if grad_hook is not None:
                    grad_hook(plan if getattr(grad_hook, "takes_plan", False) else params)
                optimizer.step()
                _plan.bump_weight_epoch()
                if self.ema is not None and not fused_ema:
                    self.ema.update(self.model)
                return loss.detach(), per.detach(), t

            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(3):
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            k0 = _lib.KERNELS[0]
            with torch.no_grad(), torch.cuda.graph(graph):
                st["loss"], st["per"], st["t"] = body()
        else:
            arena = ops.WeightArena()

            def fwd_bwd():
                noise, t, weights, ready = draw()
                x_t = self.get_q_t(st["x"], noise, t)
                per = self.per_sample_loss(self.model(x_t, t), noise, st["x"], x_t, t)
                log_losses(t, per)
                loss = reduce_loss(per, weights, ready)
                if overlap_wgrad:
                    with ops.overlap_wgrad():  # weight-gradient GEMMs on a side stream, joined before the optimiser
                        loss.backward()
                else:
                    loss.backward()
                return loss, per, t

            def body():
                arena.repack()  # one launch refreshes every bf16 weight pack from the fp32 masters
                with arena.active():
                    loss, per, t = fwd_bwd()
                if grad_hook is not None:
                    grad_hook(params)
                optimizer.step()
                _plan.bump_weight_epoch()
                ops.invalidate_weight_cache()
                if self.ema is not None and not fused_ema:
                    self.ema.update(self.model)
                return loss.detach(), per.detach(), t

            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                with arena.recording():  # learn which packs the model asks for (forward and backward)
                    fwd_bwd()
                arena.finalize(dev)
                st["arena"] = arena
                for _ in range(3):
                    optimizer.zero_grad(set_to_none=True)
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            optimizer.zero_grad(set_to_none=True)
            k0 = _lib.KERNELS[0]
            with torch.cuda.graph(graph):
                st["loss"], st["per"], st["t"] = body()
        if hasattr(optimizer, "flush_tables"):
            optimizer.flush_tables()  # pointer tables recorded during capture (gradient addresses of the graph pool)
        st["kernels_per_step"] = _lib.KERNELS[0] - k0  # kernels of this library captured in one step
        st["graph"] = graph
        self._train_graph = st

        def step(x):
            st["x"].copy_(x, non_blocking=True)
            graph.replay()
            _plan.bump_weight_epoch()  # the replayed optimiser kernel changed the weights through raw pointers
            return st["loss"]

        def sync_lr():
            for g in optimizer.param_groups:
                if "lr_dev" in g:
                    g["lr_dev"].fill_(float(g["lr"]))

        def scheduler_step(*a, **k):
            if scheduler is not None:
                scheduler.step(*a, **k)
            sync_lr()

        step.state = st
        step.sync_lr = sync_lr
        step.scheduler_step = scheduler_step
        return step
