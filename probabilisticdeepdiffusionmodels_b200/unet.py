"""The eps / variance network: same constructor, module tree (hence ``state_dict`` keys and shapes) and
``forward(x, timesteps, y=None)`` contract as the reference ``UNetModel`` (src/modules/unet.py:282-495), but the
forward pass is written B200-first:

* activations live in NHWC bf16 from the stem to the head (TMA-box friendly, 128-bit channel vectors); NCHW fp32
  exists only at the model boundary, and the stem / head convolutions do the layout change themselves;
* every conv / linear is a tcgen05 tap-GEMM whose epilogue already adds the bias, the per-sample timestep
  embedding (``h + emb_out``, unet.py:199) and the residual / skip tensor (unet.py:201, 234);
* GroupNorm32 + SiLU is one fused kernel; attention is one fused tcgen05 kernel per block.

Parameters stay fp32 with the reference's names, so reference checkpoints, ``copy.deepcopy`` (EMA), Adam and
``requires_grad_`` behave exactly as before.
"""
from abc import abstractmethod

import torch
import torch.nn as nn

from .nn import (SiLU, as_nhwc, avg_pool_nd, checkpoint, conv_nd, linear, normalization, timestep_embedding,
                 zero_module)

P = torch.ops.pddm
bf16 = torch.bfloat16


def _gn(x, norm, silu, scale=None, shift=None):
    return P.gn_silu(x, norm.weight, norm.bias, scale, shift, norm.num_groups, norm.eps, silu)[0]


def _conv(x, conv, bcast=None, residual=None, stride=1, upsample=False):
    return P.conv2d(x, conv.weight, conv.bias, bcast, residual, stride, upsample)[0]


class _SplitCols(torch.autograd.Function):
    """Column slices of one [B, sum(sizes)] matrix as separate tensors; the backward concatenates the slice
    gradients with ONE kernel (torch slicing would zero-fill and add a full-size gradient per slice)."""

    @staticmethod
    def forward(ctx, x, sizes):
        ctx.sizes = sizes
        ctx.set_materialize_grads(False)
        outs, off = [], 0
        for n in sizes:
            outs.append(x.narrow(1, off, n))
            off += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        ref = next(g for g in grads if g is not None)
        cols = [g if g is not None else ref.new_zeros((ref.shape[0], n)) for g, n in zip(grads, ctx.sizes)]
        return torch.cat(cols, dim=1), None


class EmbCtx:
    """What a ResBlock needs from the timestep embedding: SiLU(emb) (bf16) and, when the model has already run all
    ``emb_layers`` projections as ONE batched GEMM, this block's [B, Cout] slice of the result."""

    def __init__(self, act, out=None):
        self.act = act
        self.out = out or {}


class TimestepBlock(nn.Module):
    """Any module whose forward takes the (activated) timestep embedding as a second argument."""

    @abstractmethod
    def forward(self, x, emb):
        ...


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """src/modules/unet.py:39-51"""

    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if isinstance(layer, TimestepBlock) else layer(x)
        return x


class Upsample(nn.Module):
    """src/modules/unet.py:54-82: nearest x2, then conv3x3 (the upsample is an input transform of the conv op)."""

    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        self.channels, self.use_conv, self.dims = channels, use_conv, dims
        if dims != 2:
            raise ValueError("only 2D signals are supported")
        if use_conv:
            self.conv = conv_nd(dims, channels, channels, 3, padding=1)

    def forward(self, x):
        assert x.shape[-1] == self.channels
        if not self.use_conv:
            from . import functional as F
            return F.upsample2x(x)
        return _conv(x, self.conv, upsample=True)


class Downsample(nn.Module):
    """src/modules/unet.py:85-108: conv3x3 stride 2 (phase-split tap-GEMM)."""

    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        self.channels, self.use_conv, self.dims = channels, use_conv, dims
        if dims != 2:
            raise ValueError("only 2D signals are supported")
        if use_conv:
            self.op = conv_nd(dims, channels, channels, 3, stride=2, padding=1)
        else:
            self.op = avg_pool_nd(2)

    def forward(self, x):
        assert x.shape[-1] == self.channels
        if not self.use_conv:
            raise RuntimeError("avg-pool downsampling is unreachable in the reference (get_unet drops conv_resample)")
        return _conv(x, self.op, stride=2)


class ResBlock(TimestepBlock):
    """src/modules/unet.py:111-201.  ``emb`` is SiLU(time embedding) in bf16 (shared by all blocks)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False):
        super().__init__()
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm

        self.in_layers = nn.Sequential(normalization(channels), SiLU(),
                                       conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(
            SiLU(), linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm else self.out_channels))
        self.out_layers = nn.Sequential(
            normalization(self.out_channels), SiLU(), nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward(self, x, emb):
        return checkpoint(self._forward, (x, emb), self.parameters(), self.use_checkpoint)

    def _forward(self, x, emb):
        lin = self.emb_layers[1]
        emb_out = emb.out.get(id(self)) if isinstance(emb, EmbCtx) else None
        if emb_out is None:
            act = emb.act if isinstance(emb, EmbCtx) else emb
            emb_out = P.linear(act, lin.weight, lin.bias)  # fp32 [B, Cout] (or [B, 2*Cout])
        h = _gn(x, self.in_layers[0], True)
        if self.use_scale_shift_norm:
            h = _conv(h, self.in_layers[2])
            scale, shift = [t.contiguous() for t in torch.chunk(emb_out, 2, dim=1)]
            h = _gn(h, self.out_layers[0], True, scale, shift)
        else:
            h = _conv(h, self.in_layers[2], bcast=emb_out)  # conv + bias + emb_out in one epilogue
            h = _gn(h, self.out_layers[0], True)
        if self.dropout > 0 and self.training:
            h = torch.nn.functional.dropout(h, self.dropout, True)
        skip = x if isinstance(self.skip_connection, nn.Identity) else _conv(x, self.skip_connection)
        return _conv(h, self.out_layers[3], residual=skip)


class QKVAttention(nn.Module):
    """src/modules/unet.py:237-256 on the reference's [N, 3C', T] layout (N = batch*heads)."""

    def forward(self, qkv):
        n, c3, t = qkv.shape
        x = as_nhwc(qkv).reshape(n, t, c3)
        out, _ = P.attention(x, 1)
        return out.permute(0, 2, 1)


class AttentionBlock(nn.Module):
    """src/modules/unet.py:204-234"""

    def __init__(self, channels, num_heads=1, use_checkpoint=False):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads
        self.use_checkpoint = use_checkpoint
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.attention = QKVAttention()
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))

    def forward(self, x):
        return checkpoint(self._forward, (x,), self.parameters(), self.use_checkpoint)

    def _forward(self, x):
        B, H, W, C = x.shape
        qkv = _conv(_gn(x, self.norm, False), self.qkv)  # [B, H, W, 3C], head-major [q|k|v] channel packing
        a, _ = P.attention(qkv.view(B, H * W, 3 * C), self.num_heads)
        return _conv(a.view(B, H, W, C), self.proj_out, residual=x)


class UNetModel(nn.Module):
    """src/modules/unet.py:282-495"""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False):
        super().__init__()
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.num_heads_upsample = num_heads_upsample

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), SiLU(),
                                        linear(time_embed_dim, time_embed_dim))
        if self.num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_embed_dim)

        rb = dict(dims=dims, use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)
        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        input_block_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=mult * model_channels, **rb)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims)))
                input_block_chans.append(ch)
                ds *= 2

        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, time_embed_dim, dropout, **rb),
            AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads),
            ResBlock(ch, time_embed_dim, dropout, **rb))

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + input_block_chans.pop(), time_embed_dim, dropout,
                                   out_channels=model_channels * mult, **rb)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))

        self.out = nn.Sequential(normalization(ch), SiLU(),
                                 zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))
        self.batch_emb_projections = True  # run the 1 + num_blocks emb_layers projections as one GEMM

    def _emb_ctx(self, act):
        """All ResBlock ``emb_layers`` share the input SiLU(emb) (unet.py:151-157): one [B, 4mc] x [sum Cout, 4mc]^T GEMM."""
        if self.use_checkpoint:
            # activation checkpointing re-runs each block from its *tensor* inputs (nn.CheckpointFunction detaches and
            # re-requires grad on them): hand the blocks the plain SiLU(emb) tensor so that the embedding gradient
            # flows through the checkpoint like in the reference (src/modules/unet.py:160-161)
            return act
        if not self.batch_emb_projections:
            return EmbCtx(act)
        blocks = [m for m in self.modules() if isinstance(m, ResBlock)]
        w_all = torch.cat([m.emb_layers[1].weight for m in blocks], dim=0)
        b_all = torch.cat([m.emb_layers[1].bias for m in blocks], dim=0)
        sizes = tuple(m.emb_layers[1].weight.shape[0] for m in blocks)
        outs = _SplitCols.apply(P.linear(act, w_all, b_all), sizes)
        return EmbCtx(act, {id(m): o for m, o in zip(blocks, outs)})

    @property
    def inner_dtype(self):
        """dtype of the torso's *parameters* (fp32 masters); activations are bf16 with fp32 accumulation."""
        return next(self.input_blocks.parameters()).dtype

    def convert_to_fp16(self):
        raise RuntimeError("the sm_100a path always computes in bf16 with fp32 accumulation; fp16 torso is not offered")

    def convert_to_fp32(self):
        return None

    def embed(self, timesteps, y=None):
        """time_embed(timestep_embedding(t)) followed by the SiLU every ResBlock applies first (unet.py:151-157):
        returns (emb fp32 [B, 4*mc], SiLU(emb) bf16)."""
        te = timestep_embedding(timesteps, self.model_channels)
        h = P.linear(te, self.time_embed[0].weight, self.time_embed[0].bias)
        emb = P.linear(P.silu_vec(h), self.time_embed[2].weight, self.time_embed[2].bias)
        if self.num_classes is not None:
            assert y.shape == (timesteps.shape[0],)
            emb = emb + self.label_emb(y)
        return emb, P.silu_vec(emb)

    def forward(self, x, timesteps, y=None):
        """x: [N, C, H, W] (fp32 NCHW), timesteps: [N] int64 or float -> [N, out_channels, H, W] fp32."""
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if not x.is_cuda:
            raise RuntimeError("pddm_b200.UNetModel runs only on a CUDA sm_100a device (no CPU fallback)")
        out_dtype = x.dtype
        x = x.float().contiguous()
        if getattr(self, "high_precision", False) and not torch.is_grad_enabled() and y is None:
            return self.forward_high_precision(x, timesteps).to(out_dtype)
        # Fast path: the hand-scheduled forward/backward plan (plan.py) for the configurations the reference ships.
        if y is None and not (self.training and self.dropout > 0):
            from . import plan as _plan
            pl = _plan.plan_for(self, x)
            if pl is not None:
                if torch.is_grad_enabled() and any(p.requires_grad for p in pl.params):
                    out = _plan.PlanFunction.apply(pl, x, timesteps, *pl.params)
                else:
                    pl.refresh_packs()
                    out = pl.forward(x, timesteps, save=False)
                return out if out_dtype == torch.float32 else out.to(out_dtype)
        return self.forward_ops(x, timesteps, y).to(out_dtype)

    def forward_high_precision(self, x, timesteps):
        """Opt-in parity mode (highprec.py): fp32 activations, GEMM operands split into two bf16 terms on the same
        tensor-core kernels -- eps within 1e-3 of the fp32 reference, ~3x the GEMM work, forward only."""
        from .highprec import HighPrecisionUNet
        hp = self.__dict__.get("_hp_runner")
        if hp is None:
            hp = self.__dict__["_hp_runner"] = HighPrecisionUNet(self)
        return hp(x, timesteps)

    def forward_ops(self, x, timesteps, y=None):
        """The same network op by op through ``torch.ops.pddm.*`` (autograd derives the backward pass): every
        configuration, incl. scale-shift norm, dropout, class conditioning, activation checkpointing."""
        out_dtype = torch.float32
        x = x.float().contiguous()
        _, emb = self.embed(timesteps.contiguous(), y)
        emb = self._emb_ctx(emb)

        stem = self.input_blocks[0][0]
        if self.in_channels <= 4:
            h = P.stem_conv(x, stem.weight, stem.bias)[0]
        else:
            h = _conv(P.to_nhwc(x), stem)
        hs = [h]
        for module in list(self.input_blocks)[1:]:
            h = module(h, emb)
            hs.append(h)
        h = self.middle_block(h, emb)
        for module in self.output_blocks:
            h = module(P.concat_channels(h, hs.pop()), emb)
        h = _gn(h, self.out[0], True)
        head = self.out[2]
        if self.out_channels <= 8:
            out = P.head_conv(h, head.weight, head.bias)
        else:
            out = P.to_nchw(_conv(h, head))
        return out if out_dtype == torch.float32 else out.to(out_dtype)

    def get_feature_vectors(self, x, timesteps, y=None):
        """src/modules/unet.py:497-527 (NCHW fp32 copies of the hidden states)."""
        _, emb = self.embed(timesteps.contiguous(), y)
        emb = self._emb_ctx(emb)
        result = dict(down=[], up=[])
        x = x.float().contiguous()
        stem = self.input_blocks[0][0]
        h = P.stem_conv(x, stem.weight, stem.bias)[0] if self.in_channels <= 4 else _conv(P.to_nhwc(x), stem)
        hs = [h]
        result["down"].append(P.to_nchw(h))
        for module in list(self.input_blocks)[1:]:
            h = module(h, emb)
            hs.append(h)
            result["down"].append(P.to_nchw(h))
        h = self.middle_block(h, emb)
        result["middle"] = P.to_nchw(h)
        for module in self.output_blocks:
            h = module(P.concat_channels(h, hs.pop()), emb)
            result["up"].append(P.to_nchw(h))
        return result
