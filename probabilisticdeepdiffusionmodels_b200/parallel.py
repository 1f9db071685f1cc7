"""Data parallelism over the GPUs of one node: one process per GPU, ``torch.distributed`` for the plumbing.

Training shards the batch; parameters and optimiser state are replicated; the only exchange step is the
gradient all-reduce (sum, then 1/W), issued on ONE flat bucket so that NCCL sees a single large message
(196 MB fp32 for the CIFAR UNet: ~0.5 ms at the measured 725 GB/s all-reduce bus bandwidth of NVLink 5 /
NVSwitch).  GroupNorm is per sample and attention is per image, so the sharded step equals the single-GPU step
on the concatenated batch up to reduction order.  Sampling shards the batch with no collective at all.
The reference has no explicit distributed code (Lightning DDP is implied by ``gpus=N``, scripts/train.py:139-150).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_batch(x, rank, world):
    """Contiguous equal shards of the leading dimension (the remainder goes to the first ranks)."""
    n = x.shape[0]
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return x[lo: lo + base + (1 if rank < rem else 0)]


def shard_timesteps(diffusion_steps, rank, world):
    """The intermediate NLL terms t = 2 ... T (src/engine.py:455) dealt round-robin to ``world`` ranks: rank r takes
    t = 2 + r, 2 + r + world, ...  (every t exactly once; consecutive ranks see similar noise levels, i.e. similar work)."""
    return list(range(2 + rank, diffusion_steps + 1, world))


def all_ranks(flag, device=None, group=None):
    """True iff ``flag`` is true on EVERY rank (one MIN all-reduce of an int32; a host sync).  Loops whose body
    contains collectives -- a captured training step replayed "until something local happens" -- must leave on this,
    not on the local condition: ranks that iterate a different number of times pair mismatched collectives."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def all_reduce_nll(parts, group=None):
    """Sum the ranks' partial NLL results with ONE all-reduce: ``parts`` = {"L_int", "L_0", "L_T": fp32 [B] per-sample
    vectors, "mse_sum": 0-dim float64 tensor, "mse_count": float}.  Returns the totals (same keys) on every rank."""
    B = parts["L_int"].shape[0]
    dev = parts["L_int"].device
    flat = torch.cat([parts["L_int"].double(), parts["L_0"].double(), parts["L_T"].double(),
                      parts["mse_sum"].double().reshape(1),
                      torch.tensor([parts["mse_count"]], dtype=torch.float64, device=dev)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return {"L_int": flat[:B].float(), "L_0": flat[B: 2 * B].float(), "L_T": flat[2 * B: 3 * B].float(),
            "mse_sum": flat[3 * B], "mse_count": float(flat[3 * B + 1])}


class FlatGradAllReduce:
    """``hook(params)``: average ``p.grad`` over the process group through one flat bucket.

    Works on CPU tensors with gloo (tests) and on CUDA tensors with NCCL (also inside CUDA-graph capture)."""

    def __init__(self, group=None, bucket_dtype=torch.float32):
        self.group = group
        self.bucket_dtype = bucket_dtype
        self._flat = None

    def __call__(self, params):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, dtype=self.bucket_dtype, device=grads[0].device)
        views, off = [], 0
        for g in grads:
            views.append(self._flat[off: off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        self._flat.mul_(1.0 / dist.get_world_size(self.group))
        torch._foreach_copy_(grads, views)


class ArenaGradAllReduce:
    """``hook(plan)`` for the hand-scheduled step (plan.py): the gradients are born in ONE flat fp32 arena, so the
    all-reduce runs on it in place -- no gather / scatter copies, no scaling pass (the 1/W is folded into the fused
    Adam kernel through ``optimizer.grad_scale``).  ``buckets`` > 1 splits the message into that many contiguous
    all-reduces (NCCL pipelines them; useful when something else can run in between)."""

    takes_plan = True

    def __init__(self, optimizer=None, group=None, buckets=1):
        self.group, self.buckets = group, buckets
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if optimizer is not None:
            optimizer.grad_scale = 1.0 / world
        self.scale_in_optimizer = optimizer is not None

    def __call__(self, plan):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        flat = plan.grad_arena
        n = flat.numel()
        step = -(-n // self.buckets)
        step = (step + 1023) // 1024 * 1024
        for lo in range(0, n, step):
            dist.all_reduce(flat[lo: lo + step], op=dist.ReduceOp.SUM, group=self.group)
        if not self.scale_in_optimizer:
            flat.mul_(1.0 / dist.get_world_size(self.group))


class OverlappedArenaAllReduce:
    """The arena all-reduce hidden under the backward pass of the hand-scheduled step (plan.py).

    The gradient arena is laid out in the order the backward pass finishes its parts, last first
    (``plan.bucket_bounds``): [small parameters, emb_layers, time_embed, stem, full-resolution encoder | encoder below
    full resolution | middle block, decoder, head].  ``plan.backward`` calls ``range_final`` as soon as every kernel
    that writes a tail range has been issued; the range is all-reduced on a high-priority communication stream (after
    the main and the weight-gradient streams' work issued so far) while the backward pass goes on.  ``hook(plan)``
    after the backward pass reduces the head of the arena -- all that is still exposed, 10 % of the bytes for the
    CIFAR UNet -- and joins the streams.  1/W is folded into the fused Adam kernel (``optimizer.grad_scale``).

    Why the SM reserve: the GEMM kernels are persistent, one CTA per SM, each with a fixed share of the tiles.  NCCL's
    CTAs cannot share an SM with them (shared memory), so either NCCL waits for a whole GEMM kernel or -- once it holds
    its SMs -- the GEMM's CTAs that found no SM run as a second wave and double that kernel's time (what r1 measured
    as "bucketing is neutral").  While a range is in flight the persistent kernels are therefore launched on
    ``sm_count - sm_reserve`` SMs (``pddm_set_sm_reserve``) and the communicator of the overlapped ranges is created
    with ``max_ctas = sm_reserve``.  The head of the arena goes through the default group (all CTAs NCCL wants: it is
    exposed anyway).  Works inside CUDA-graph capture (fork / join become graph edges) and on CPU tensors with gloo
    (no streams, no reserve)."""

    takes_plan = True

    def __init__(self, optimizer=None, group=None, sm_reserve=8, overlap_group=None, comm="own"):
        self.group, self.sm_reserve = group, int(sm_reserve)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if optimizer is not None:
            optimizer.grad_scale = 1.0 / world
        self.scale_in_optimizer = optimizer is not None
        self.overlap_group = overlap_group
        # comm: "own" = a communicator of its own limited to sm_reserve CTAs; "own-default" = of its own, NCCL's
        # defaults; "same" = the group's communicator
        if overlap_group is None and comm != "same" and dist.is_initialized() and world > 1 and \
                dist.get_backend(group) == "nccl":
            opts = dist.ProcessGroupNCCL.Options()
            if comm == "own":
                opts.config.max_ctas = max(1, self.sm_reserve)
                opts.config.min_ctas = 1
            ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
            self.overlap_group = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
        self.stream = None
        self._done_from = None  # arena offset from which everything has been handed to the communication stream

    def _active(self):
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def attach(self, plan):
        """Have ``plan.backward`` report finished arena ranges to this object."""
        plan.comm = self if self._active() else None
        self._done_from = None
        if plan.comm is not None and plan.grad_arena.is_cuda:
            from . import functional as F
            F.set_sm_reserve(0)
        return self

    def range_final(self, plan, lo, hi):
        """Every kernel writing ``plan.grad_arena[lo:hi]`` has been issued (main or weight-gradient stream)."""
        if not self._active() or hi <= lo:
            return
        flat = plan.grad_arena
        if hi == flat.numel():
            self._done_from = None  # first range of a step (also after a step that was abandoned half-way)
        if self._done_from is not None and hi != self._done_from:
            raise RuntimeError("OverlappedArenaAllReduce: ranges must arrive tail first and be contiguous")
        if flat.is_cuda:
            from . import functional as F
            dev = flat.device
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=dev, priority=-1)
            self.stream.wait_stream(torch.cuda.current_stream(dev))
            if getattr(plan, "side", None) is not None and getattr(plan, "overlap", False):
                self.stream.wait_stream(plan.side)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.overlap_group or self.group)
            if self.sm_reserve > 0:
                F.set_sm_reserve(self.sm_reserve)  # kernels launched from here on leave NCCL its SMs
        else:
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        self._done_from = lo

    def __call__(self, plan):
        if not self._active():
            return
        if not hasattr(plan, "grad_arena"):
            raise TypeError("OverlappedArenaAllReduce needs the hand-scheduled step (plan.UNetPlan): this model takes "
                            "the op-by-op path -- use FlatGradAllReduce / BucketedGradAllReduce")
        flat = plan.grad_arena
        head = flat.numel() if self._done_from is None else self._done_from
        self._done_from = None
        if flat.is_cuda:
            from . import functional as F
            F.set_sm_reserve(0)
            if self.stream is not None:  # join first: two communicators are never in flight together
                torch.cuda.current_stream(flat.device).wait_stream(self.stream)
        if head > 0:
            dist.all_reduce(flat[:head], op=dist.ReduceOp.SUM, group=self.group)
        if not self.scale_in_optimizer:
            flat.mul_(1.0 / dist.get_world_size(self.group))


class BucketedGradAllReduce:
    """Gradient averaging that overlaps the backward pass: ``hook = BucketedGradAllReduce(params)`` registers a
    post-accumulate-grad hook on every parameter; as soon as the gradients of one bucket (parameters in the order
    their gradients become ready, ~32 MB of fp32 each) are complete, the bucket is copied into its slice of ONE flat
    buffer and all-reduced on a communication stream while autograd keeps producing the next bucket.  Calling
    ``hook(params)`` after ``backward()`` joins the communication stream and points every ``p.grad`` at its
    (averaged) slice of the flat buffer -- no copy back.  The first backward only learns the ready order and
    reduces everything at the end.  Works eagerly, under CUDA-graph capture (fork / join become graph edges) and on
    CPU tensors with gloo (no streams).  With the weight-gradient GEMMs on a side stream (``ops.overlap_wgrad``) the
    communication stream also waits for that stream."""

    def __init__(self, params, bucket_bytes=32 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_bytes = bucket_bytes
        self._fired = []
        self.buckets = None  # list of lists of param indices
        self.handles = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]
        self.comm = None

    def _active(self):
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _make_hook(self, i):
        def hook(_p):
            if not self._active():
                return
            if self.buckets is None:
                self._fired.append(i)
                return
            b = self.bucket_of[i]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _build(self):
        order = list(dict.fromkeys(self._fired))
        order += [i for i in range(len(self.params)) if i not in set(order) and self.params[i].grad is not None]
        self.buckets, cur, size = [], [], 0
        for i in order:
            cur.append(i)
            size += self.params[i].numel() * 4
            if size >= self.bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {i: b for b, idx in enumerate(self.buckets) for i in idx}
        dev = self.params[order[0]].device
        n = sum(self.params[i].numel() for i in order)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.views, self.slices, off = [], [], 0
        for idx in self.buckets:
            o0, vs = off, []
            for i in idx:
                p = self.params[i]
                vs.append(self.flat[off: off + p.numel()].view_as(p))
                off += p.numel()
            self.views.append(vs)
            self.slices.append(self.flat[o0: off])
        if dev.type == "cuda":
            self.comm = torch.cuda.Stream(device=dev)
        self._reset()

    def _reset(self):
        self.pending = [len(idx) for idx in self.buckets]
        self.launched = [False] * len(self.buckets)

    def _launch(self, b):
        grads = [self.params[i].grad for i in self.buckets[b]]
        inv = 1.0 / dist.get_world_size(self.group)
        if self.comm is not None:
            from . import ops
            cur = torch.cuda.current_stream(self.flat.device)
            self.comm.wait_stream(cur)
            side = ops._OVERLAP["side"]
            if ops._OVERLAP["on"] and side is not None:
                self.comm.wait_stream(side)  # weight gradients are written by GEMMs on the side stream
            with torch.cuda.stream(self.comm):
                torch._foreach_copy_(self.views[b], grads)
                dist.all_reduce(self.slices[b], op=dist.ReduceOp.SUM, group=self.group)
                self.slices[b].mul_(inv)
        else:
            torch._foreach_copy_(self.views[b], grads)
            dist.all_reduce(self.slices[b], op=dist.ReduceOp.SUM, group=self.group)
            self.slices[b].mul_(inv)
        self.launched[b] = True

    def __call__(self, params=None):
        if not self._active():
            return
        if self.buckets is None:
            self._build()
        for b in range(len(self.buckets)):  # learning pass, or parameters whose gradient never arrived
            if not self.launched[b]:
                if any(self.params[i].grad is None for i in self.buckets[b]):
                    raise RuntimeError("BucketedGradAllReduce: a parameter of the recorded set received no gradient")
                self._launch(b)
        if self.comm is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm)
        for b, idx in enumerate(self.buckets):
            for i, v in zip(idx, self.views[b]):
                self.params[i].grad = v
        self._reset()

    def remove(self):
        for h in self.handles:
            h.remove()


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
