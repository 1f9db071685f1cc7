"""Operator seam: drop-in for the reference's ``src/modules/nn.py`` factories (same names, argument meaning and
error behaviour), backed by the sm_100a kernels.

Each factory returns a module that *is* the corresponding ``torch.nn`` class for parameter creation,
initialisation and ``state_dict`` naming (so reference checkpoints load and ``copy.deepcopy`` works), with a
``forward`` that calls the ``pddm`` ops.  Modules accept the reference's logical NCHW tensors: an fp32 NCHW
tensor is converted once to the internal NHWC-bf16 format, and outputs are logical-NCHW views of NHWC storage
(``torch.channels_last``), so a stack of these modules never converts again.  ``unet.py`` skips even that and
drives the ops on NHWC tensors directly.
"""
import math

import torch
import torch.nn as nn

from . import ops  # noqa: F401  (registers torch.ops.pddm.*)

P = torch.ops.pddm
bf16 = torch.bfloat16


def as_nhwc(x):
    """logical [B,C,H,W] (or [B,C,T]) -> contiguous NHWC bf16 [B,H,W,C] (or [B,1,T,C])."""
    if x.dim() == 3:
        x = x.unsqueeze(2)
    if x.dtype == bf16:
        v = x.permute(0, 2, 3, 1)
        if v.is_contiguous():
            return v
        return P.to_nhwc(x.float())
    return P.to_nhwc(x.float().contiguous())


def as_nchw_view(y, like):
    """NHWC bf16 result -> logical NCHW view (3-D if the input was 3-D)."""
    v = y.permute(0, 3, 1, 2)
    return v.squeeze(2) if like.dim() == 3 else v


class SiLU(nn.Module):
    """src/modules/nn.py:13-15.  Inside this package's ``unet.py`` SiLU is always fused into the GroupNorm kernel;
    the stand-alone module serves models assembled from these factories the way the reference's own unet.py does
    (``Sequential(normalization, SiLU, conv)``, src/modules/unet.py:146-150): fp32 [B, N] embedding vectors and
    feature maps (logical NCHW / [B, C, T] views of NHWC bf16 storage, or fp32 NCHW which is converted once)."""

    def forward(self, x):
        if x.dtype == torch.float32 and x.dim() == 2:
            return P.silu_vec(x.contiguous())
        if x.dim() == 2:
            return P.silu_vec(x.float().contiguous())
        if x.dim() not in (3, 4):
            raise RuntimeError(f"SiLU: unsupported input rank {x.dim()}")
        h = as_nhwc(x)
        if not h.is_contiguous():
            h = h.contiguous()
        return as_nchw_view(P.silu_map(h), x)


class GroupNorm32(nn.GroupNorm):
    """src/modules/nn.py:18-20: statistics and affine in fp32, result in the activation dtype."""

    def forward(self, x, silu=False):
        y, _, _ = P.gn_silu(as_nhwc(x), self.weight, self.bias, None, None, self.num_groups, self.eps, silu)
        return as_nchw_view(y, x)


class Conv2d(nn.Conv2d):
    def forward(self, x):
        if self.kernel_size not in ((3, 3), (1, 1)) or self.stride not in ((1, 1), (2, 2)) or \
                self.padding != ((self.kernel_size[0] - 1) // 2,) * 2:
            raise ValueError(f"unsupported conv geometry k={self.kernel_size} s={self.stride} p={self.padding}")
        if self.kernel_size == (3, 3) and self.stride == (1, 1):
            # thin convolutions take the dedicated paths the UNet itself uses for its stem and head
            if self.in_channels <= 4 and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous():
                return as_nchw_view(P.stem_conv(x, self.weight, self.bias)[0], x)
            if self.out_channels <= 8:
                return P.head_conv(as_nhwc(x), self.weight, self.bias)  # fp32 NCHW, like the model output
        y, _ = P.conv2d(as_nhwc(x), self.weight, self.bias, None, None, self.stride[0], False)
        return as_nchw_view(y, x)


class Conv1d(nn.Conv1d):
    def forward(self, x):
        if self.kernel_size != (1,) or self.stride != (1,):
            raise ValueError("only pointwise Conv1d is supported (attention qkv / proj_out)")
        y, _ = P.conv2d(as_nhwc(x), self.weight, self.bias, None, None, 1, False)
        return as_nchw_view(y, x)


class Linear(nn.Linear):
    def forward(self, x):
        if x.dtype != bf16:
            x = P.cast_bf16(x.float().contiguous())
        return P.linear(x.contiguous(), self.weight, self.bias)


def conv_nd(dims, *args, **kwargs):
    """src/modules/nn.py:23-33"""
    if dims == 1:
        return Conv1d(*args, **kwargs)
    elif dims == 2:
        return Conv2d(*args, **kwargs)
    elif dims == 3:
        raise ValueError("3D convolutions are not part of the sm_100a hot path")
    raise ValueError(f"unsupported dimensions: {dims}")


def linear(*args, **kwargs):
    """src/modules/nn.py:36-40"""
    return Linear(*args, **kwargs)


def avg_pool_nd(dims, *args, **kwargs):
    """src/modules/nn.py:43-53 -- unreachable in the reference (``conv_resample`` is dropped by get_unet,
    src/modules/__init__.py:22-23 vs 36-49); kept for API parity as the stock torch module."""
    if dims == 1:
        return nn.AvgPool1d(*args, **kwargs)
    elif dims == 2:
        return nn.AvgPool2d(*args, **kwargs)
    elif dims == 3:
        return nn.AvgPool3d(*args, **kwargs)
    raise ValueError(f"unsupported dimensions: {dims}")


def update_ema(target_params, source_params, rate=0.99):
    """targ <- rate * targ + (1 - rate) * src over two parameter sequences (src/modules/nn.py:56-66): two multi-tensor
    launches instead of two launches per tensor."""
    targs = [t.detach() for t in target_params]
    srcs = [s.detach() for s in source_params][: len(targs)]
    targs = targs[: len(srcs)]
    if targs:
        torch._foreach_mul_(targs, rate)
        torch._foreach_add_(targs, srcs, alpha=1 - rate)


def zero_module(module):
    """src/modules/nn.py:69-75"""
    for p in module.parameters():
        p.detach().zero_()
    return module


def scale_module(module, scale):
    for p in module.parameters():
        p.detach().mul_(scale)
    return module


def mean_flat(tensor):
    """src/modules/nn.py:87-91"""
    return tensor.mean(dim=list(range(1, len(tensor.shape))))


def normalization(channels):
    """src/modules/nn.py:94-101"""
    return GroupNorm32(32, channels)


def timestep_embedding(timesteps, dim, max_period=10000):
    """src/modules/nn.py:104-122 -> [N, dim] (bf16: it feeds the time-embedding GEMM)."""
    return P.timestep_embedding(timesteps.contiguous(), dim, float(max_period))


def checkpoint(func, inputs, params, flag):
    """src/modules/nn.py:125-143"""
    if flag:
        args = tuple(inputs) + tuple(params)
        return CheckpointFunction.apply(func, len(inputs), *args)
    return func(*inputs)


class CheckpointFunction(torch.autograd.Function):
    """src/modules/nn.py:146-171: recompute-in-backward."""

    @staticmethod
    def forward(ctx, run_function, length, *args):
        ctx.run_function = run_function
        ctx.input_tensors = list(args[:length])
        ctx.input_params = list(args[length:])
        with torch.no_grad():
            return ctx.run_function(*ctx.input_tensors)

    @staticmethod
    def backward(ctx, *output_grads):
        ins = [x.detach().requires_grad_(True) for x in ctx.input_tensors]
        with torch.enable_grad():
            outs = ctx.run_function(*[x.view_as(x) for x in ins])
        grads = torch.autograd.grad(outs, ins + ctx.input_params, output_grads, allow_unused=True)
        del ctx.input_tensors, ctx.input_params
        return (None, None) + grads
