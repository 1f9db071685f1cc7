"""Adam (+ optional EMA of the weights) over a model's parameter list in ONE kernel launch.

Semantics are ``torch.optim.Adam``'s defaults as the reference configures it (src/engine.py:238-248: lr from the
optimizer config, betas (0.9, 0.999), eps 1e-8, optional weight decay, no amsgrad) and ``Ema.update``
(src/modules/ema.py:21-33).  The kernel (``pddm_adam_ema_multi``) walks a device table of
(param, grad, exp_avg, exp_avg_sq, ema) pointers, so autograd's one-gradient-tensor-per-parameter layout is used
as is: no flattening copies, one launch instead of torch's 18 multi-tensor launches for the CIFAR UNet.

The step count lives on the device (bias correction is computed in the kernel), so ``step()`` can be captured in a
CUDA graph (after one eager warm-up step).  During capture only the launch is recorded; the pointer table (the
gradient addresses are the graph pool's, fixed across replays) lives outside the graph pool and is uploaded by
``flush_tables()`` once the capture has ended.
"""
import ctypes as C

import torch

from . import _lib as L


class AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("ema", C.c_void_p), ("n", C.c_int64)]


ADAM_CHUNK = 8192  # PDDM_ADAM_CHUNK


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, ema_params=None,
                 ema_decay=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._ema = {}
        if ema_params is not None:
            ps = [p for g in self.param_groups for p in g["params"]]
            ema_params = list(ema_params)
            assert len(ema_params) == len(ps)
            self._ema = {id(p): e for p, e in zip(ps, ema_params)}
        self.ema_decay = ema_decay
        self._tables = []   # device tables referenced by captured graphs
        self._pending = []  # (device table, host bytes) recorded under capture, uploaded by flush_tables()
        self._spare = {}    # per param group: a device table allocated eagerly, handed to the next captured step
        self._step_dev = None

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    # ---- checkpoint / resume: the step count lives on the device; state_dict() carries it per parameter as ``step``
    # (torch.optim.Adam's own format, so the two optimizers' checkpoints are interchangeable) and load_state_dict()
    # restores it -- bias correction resumes where it stopped instead of restarting at step 1.
    def state_dict(self):
        step = float(int(self._step_dev.item())) if self._step_dev is not None else 0.0
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] = torch.tensor(step)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        step = 0
        for st in self.state.values():
            if "step" in st:
                step = max(step, int(float(st["step"])))
        ps = [p for g in self.param_groups for p in g["params"]]
        if ps and ps[0].is_cuda:
            self._step_dev = torch.full((1,), step, dtype=torch.int32, device=ps[0].device)
        else:
            self._step_dev = None
            self._resume_step = step
        for st in self.state.values():  # the kernel wants contiguous fp32 moments on the parameter's device
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st:
                    st[k] = st[k].float().contiguous()

    def flush_tables(self):
        """Upload the pointer tables recorded by ``step()`` calls made under CUDA-graph capture (call after capture)."""
        for table, raw in self._pending:
            table.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        self._pending = []

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        capturing = torch.cuda.is_current_stream_capturing()
        stepped = False
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            L.require_device(ps[0])
            if self._step_dev is None:
                self._step_dev = torch.full((1,), int(getattr(self, "_resume_step", 0)), dtype=torch.int32, device=dev)
            if not stepped:
                L.call("pddm_counter_add", L.ptr(self._step_dev), 1, L.stream())
                stepped = True
            descs = (AdamTensor * len(ps))()
            blocks = []
            for i, p in enumerate(ps):
                g = p.grad
                if g.dtype != torch.float32 or p.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 parameters and gradients")
                st = self._state(p)
                e = self._ema.get(id(p))
                d = descs[i]
                d.param, d.grad, d.exp_avg, d.exp_avg_sq = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), \
                    st["exp_avg_sq"].data_ptr()
                d.ema = e.data_ptr() if e is not None else None
                d.n = p.numel()
                blocks += [(i, c0) for c0 in range(0, p.numel(), ADAM_CHUNK)]
            raw = bytes(descs) + torch.tensor(blocks, dtype=torch.int32).numpy().tobytes()
            gi = id(group)
            if capturing:
                # The table must NOT come from the graph's private pool: that memory is only meaningful in capture
                # order (blocks freed earlier in the capture, e.g. activations, are reused for later allocations and
                # rewritten on every replay).  A spare allocated by an earlier eager step is handed to the graph.
                table = self._spare.pop(gi, None)
                if table is None or table.numel() != len(raw):
                    raise RuntimeError("FusedAdam.step() under CUDA-graph capture needs one eager warm-up step first")
                self._pending.append((table, raw))
                self._tables.append(table)  # owned by the graph from now on
            else:
                table = torch.empty(len(raw), dtype=torch.uint8, device=dev)
                table.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory(), non_blocking=True)
                sp = self._spare.get(gi)
                if sp is None or sp.numel() != len(raw):
                    self._spare[gi] = torch.empty(len(raw), dtype=torch.uint8, device=dev)
            hp = L.AdamParams()
            hp.lr, (hp.beta1, hp.beta2), hp.eps, hp.weight_decay = group["lr"], group["betas"], group["eps"], \
                group["weight_decay"]
            hp.ema_decay = self.ema_decay if self.ema_decay is not None else 0.0
            hp.grad_scale, hp.step, hp.step_dev = float(getattr(self, "grad_scale", 1.0)), 1, L.ptr(self._step_dev)
            lr_dev = group.get("lr_dev")
            hp.lr_dev = L.ptr(lr_dev) if lr_dev is not None else None
            L.call("pddm_adam_ema_multi", table.data_ptr(), table.data_ptr() + C.sizeof(descs), len(blocks),
                   C.byref(hp), L.stream())
        # the kernel wrote parameters (and EMA shadows) through raw pointers: ``_version`` did not move, so every
        # cache of bf16 weight packs has to be told
        from . import ops, plan
        ops.invalidate_weight_cache()
        plan.bump_weight_epoch()
        return loss
