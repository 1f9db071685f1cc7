"""torch.library registration of the sm_100a kernels (namespace ``pddm``) with fake (shape) and autograd rules.

Each forward op has a matching ``*_bwd`` op; ``register_autograd`` wires them together, so the ops compose
under ``loss.backward()``, ``torch.no_grad()`` and CUDA-graph capture.  Tensors inside the network are
contiguous NHWC bf16; the ops never fall back to eager torch math -- on a non-sm_100 device they raise.

Weight operands: parameters stay fp32 ``nn.Parameter``s with the reference's names/shapes; the bf16 GEMM packs
are produced by ``pddm_pack_conv_weight`` (eager training: re-packed on every call, the weights change each step),
by ONE ``pddm_pack_weights_multi`` launch per step over a flat arena (``WeightArena``, the captured training step:
0.28 ms for 49 M parameters), or cached per parameter object and version inside ``frozen_weights()`` (sampling /
evaluation).
"""
import contextlib
import weakref
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import functional as F

bf16, f32 = torch.bfloat16, torch.float32

_FROZEN = [0]
_EPOCH = [0]
_PACK_CACHE = {}


@contextlib.contextmanager
def frozen_weights():
    """Parameters are not going to change inside this block: reuse cached bf16 weight packs."""
    _FROZEN[0] += 1
    try:
        yield
    finally:
        _FROZEN[0] -= 1


def invalidate_weight_cache():
    _EPOCH[0] += 1
    _PACK_CACHE.clear()


def _packed(w, mode):
    if not _FROZEN[0]:
        return F.pack_weight(w, mode)
    key = (w.data_ptr(), mode)
    ent = _PACK_CACHE.get(key)
    stamp = (w._version, _EPOCH[0], tuple(w.shape))
    # the weak reference ties the entry to this very tensor object: a freed parameter's address can be handed to a
    # new parameter of the same shape and version (another model built later in the process)
    owner = w._base if w._base is not None else w  # views (linear weights seen as [N, K, 1]) belong to their base
    if ent is None or ent[0] != stamp or ent[2]() is not owner:
        if len(_PACK_CACHE) > 4096:  # entries of dead tensors
            for k in [k for k, e in _PACK_CACHE.items() if e[2]() is None]:
                del _PACK_CACHE[k]
        ent = (stamp, F.pack_weight(w, mode), weakref.ref(owner))
        _PACK_CACHE[key] = ent
    return ent[1]


_OVERLAP = {"on": 0, "side": None, "keep": []}


@contextlib.contextmanager
def overlap_wgrad():
    """Run every conv weight-gradient GEMM of the enclosed backward pass on a side stream.

    A weight gradient is a leaf of the backward graph (only the optimiser consumes it), while the data gradient
    feeds the next GroupNorm backward; the wgrad kernel is TMA/tensor bound and light on registers, the GroupNorm /
    reduction kernels are LSU/L2 bound, so they co-reside on the SMs.  The streams are joined on exit (also valid
    inside CUDA-graph capture: fork and join become graph edges).  The gradient tensor must reach ``weight.grad`` without
    any main-stream kernel touching it before the join: autograd's AccumulateGrad adopts it as is when ``.grad is None``
    and nobody else holds a reference (which is why the keep-alive list below excludes it -- an extra reference made
    AccumulateGrad clone the still unwritten buffer: found by tests/test_fullsize_gpu.py); parameters that already
    carry a gradient fall back to the main stream."""
    cur = torch.cuda.current_stream()
    if _OVERLAP["side"] is None or _OVERLAP["side"].device != cur.device:
        _OVERLAP["side"] = torch.cuda.Stream(device=cur.device)
    _OVERLAP["on"] += 1
    try:
        yield
    finally:
        _OVERLAP["on"] -= 1
        torch.cuda.current_stream().wait_stream(_OVERLAP["side"])
        _OVERLAP["keep"].clear()


def _wgrad(xin, dy, taps, B, H, W, Cin, Cout, shape, weight=None):
    # The side-stream result is only safe if autograd ADOPTS it as ``weight.grad`` (no kernel).  If a gradient is
    # already there autograd would accumulate into it on the main stream, racing with the GEMM: stay on the main stream.
    if not _OVERLAP["on"] or (weight is not None and getattr(weight, "grad", None) is not None):
        return F.tap_wgrad(xin, dy, taps, B, H, W, Cin, Cout, shape)
    side = _OVERLAP["side"]
    side.wait_stream(torch.cuda.current_stream())  # dy / xin are ready
    return F.tap_wgrad(xin, dy, taps, B, H, W, Cin, Cout, shape, launch_stream=side, keep=_OVERLAP["keep"])


class WeightArena:
    """All bf16 weight packs of a model in one flat buffer, refreshed by ONE kernel launch per optimiser step.

    Usage (Engine.capture_train_step): run one step inside ``recording()`` to learn which (weight, mode, padding)
    packs the model asks for, ``finalize()``, then call ``repack()`` once per step and run the model inside
    ``active()``: every ``pack_weight`` request is served as a view of the arena."""

    def __init__(self):
        self.requests = {}
        self.views = {}
        self.mode = "off"
        self._descs = self._blocks = self._flat = None
        self.nblocks = 0

    @staticmethod
    def _key(w, mode, cop, cip):
        return (w.data_ptr(), mode, cop, cip)

    def _hook(self, w, mode, cop, cip, ntaps):
        key = self._key(w, mode, cop, cip)
        if self.mode == "record":
            self.requests[key] = (w, mode, cop, cip, ntaps)
            return None
        return self.views.get(key)

    @contextlib.contextmanager
    def recording(self):
        prev, self.mode = (F.PACK_HOOK[0], self.mode), "record"
        F.PACK_HOOK[0] = self._hook
        try:
            yield self
        finally:
            F.PACK_HOOK[0], self.mode = prev

    @contextlib.contextmanager
    def active(self):
        prev, self.mode = (F.PACK_HOOK[0], self.mode), "serve"
        F.PACK_HOOK[0] = self._hook
        try:
            yield self
        finally:
            F.PACK_HOOK[0], self.mode = prev

    def finalize(self, device):
        import ctypes as C

        from . import _lib as L
        total, metas = 0, []
        for key, (w, mode, cop, cip, ntaps) in self.requests.items():
            n = cop * cip * ntaps
            metas.append((key, w, mode, cop, cip, ntaps, total, n))
            total += (n + 7) // 8 * 8  # keep every pack 16-byte aligned
        self._flat = torch.empty(max(total, 8), dtype=bf16, device=device)
        descs = (L.PackDesc * len(metas))()
        blocks = []
        for i, (key, w, mode, cop, cip, ntaps, off, n) in enumerate(metas):
            view = self._flat[off: off + n].view((cop, ntaps, cip) if mode == 0 else (cip, ntaps, cop))
            self.views[key] = view
            d = descs[i]
            d.src, d.dst = w.data_ptr(), view.data_ptr()
            d.Cout, d.Cin, d.ntaps, d.mode, d.Cout_pad, d.Cin_pad = w.shape[0], w.shape[1], ntaps, mode, cop, cip
            tiles = -(-cop // L.PACK_TILE) * -(-cip // L.PACK_TILE)
            blocks += [(i, t) for t in range(tiles)]
            self.max_ntaps = max(getattr(self, "max_ntaps", 1), ntaps)
        self._descs = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(device)
        self._blocks = torch.tensor(blocks, dtype=torch.int32).to(device)
        self.nblocks = len(blocks)
        self._keep = [m[1] for m in metas]
        return self

    def repack(self):
        from . import _lib as L
        L.call("pddm_pack_weights_multi", L.ptr(self._descs), L.ptr(self._blocks), self.nblocks, self.max_ntaps,
               L.stream())


def _empty(like):
    return torch.empty(0, dtype=f32, device=like.device)


def _opt(t):
    return None if (t is None or t.numel() == 0) else t


# ================================================================================================ conv2d
@torch.library.custom_op("pddm::conv2d", mutates_args=())
def conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor], bcast: Optional[Tensor], residual: Optional[Tensor],
           stride: int, upsample: bool) -> Tuple[Tensor, Tensor]:
    """NHWC bf16 conv (3x3 pad 1 | 1x1), optional nearest-x2 upsample in front, stride 1|2, fused epilogue
    ``+ bias[n] + bcast[b, n] + residual``.  Returns (y, x_gemm) where x_gemm is the tensor the GEMM actually read
    (upsampled / phase-split copy) or an empty placeholder when that is ``x`` itself."""
    B, H, W, Cin = x.shape
    k = weight.shape[-1] if weight.dim() == 4 else 1
    wp = _packed(weight, 0)
    xin = x
    aux = _empty(x)
    if upsample:
        xin = aux = F.upsample2x(x)
        H, W = 2 * H, 2 * W
    if stride == 2:
        assert k == 3 and H % 2 == 0 and W % 2 == 0
        xin = aux = F.phase_split(xin)
        H, W = H // 2, W // 2
        taps = F.taps_stride2(B)
    else:
        taps = F.taps_3x3() if k == 3 else F.taps_1x1()
    y = F.tap_gemm(xin, wp, taps, B, H, W, bias=bias, bcast=bcast, residual=residual)
    return y, aux


@conv2d.register_fake
def _(x, weight, bias, bcast, residual, stride, upsample):
    B, H, W, _ = x.shape
    if upsample:
        H, W = 2 * H, 2 * W
    aux = x.new_empty((0,), dtype=f32)
    if stride == 2:
        aux = x.new_empty((4 * B, H // 2, W // 2, x.shape[-1]))
        H, W = H // 2, W // 2
    elif upsample:
        aux = x.new_empty((B, H, W, x.shape[-1]))
    return x.new_empty((B, H, W, weight.shape[0])), aux


@torch.library.custom_op("pddm::conv2d_bwd", mutates_args=())
def conv2d_bwd(dy: Tensor, xin: Tensor, weight: Tensor, stride: int, upsample: bool, in_h: int, in_w: int,
               need_dx: bool, need_dbias: bool, need_dbcast: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (dx, dweight, dbias, dbcast); unused outputs are empty placeholders.  xin = tensor the forward GEMM read."""
    B, H, W, Cout = dy.shape  # tile space of the forward GEMM
    Cin = weight.shape[1]
    k = weight.shape[-1] if weight.dim() == 4 else 1
    if stride == 2:
        taps = F.taps_stride2(B)
    else:
        taps = F.taps_3x3() if k == 3 else F.taps_1x1()
    dw = _wgrad(xin, dy, taps, B, H, W, Cin, Cout, tuple(weight.shape), weight)
    dbias = F.colsum(dy, Cout) if need_dbias else _empty(dy)
    dbcast = F.colsum_per_sample(dy) if need_dbcast else _empty(dy)
    dx = _empty(dy)
    if need_dx:
        wp1 = _packed(weight, 1)  # [Cin, taps(flipped), Cout]
        if stride == 2:
            dx = torch.empty((B, 2 * H, 2 * W, Cin), dtype=bf16, device=dy.device)
            for a in range(2):
                for b in range(2):
                    F.tap_gemm(dy, wp1, F.taps_stride2_dgrad(a, b), B, H, W, out=dx, out_hw=(2 * H, 2 * W),
                               out_map=(2, 2, a, b))
        else:
            dx = F.tap_gemm(dy, wp1, taps, B, H, W)
        if upsample:
            dx = F.upsample2x_bwd(dx)
    return dx, dw, dbias, dbcast


@conv2d_bwd.register_fake
def _(dy, xin, weight, stride, upsample, in_h, in_w, need_dx, need_dbias, need_dbcast):
    B = dy.shape[0]
    e = dy.new_empty((0,), dtype=f32)
    dx = dy.new_empty((B, in_h, in_w, weight.shape[1])) if need_dx else e
    return (dx, weight.new_empty(weight.shape), dy.new_empty((dy.shape[-1],), dtype=f32) if need_dbias else e,
            dy.new_empty((B, dy.shape[-1]), dtype=f32) if need_dbcast else e)


def _conv_setup(ctx, inputs, output):
    x, weight, bias, bcast, residual, stride, upsample = inputs
    _, aux = output
    ctx.set_materialize_grads(False)  # no zero-filled gradient tensors for the auxiliary outputs
    ctx.save_for_backward(x if aux.numel() == 0 else aux, weight)
    ctx.meta = (stride, upsample, x.shape[1], x.shape[2])
    ctx.needs = (x.requires_grad, bias is not None and bias.requires_grad, bcast is not None and bcast.requires_grad,
                 residual is not None and residual.requires_grad)


def _conv_backward(ctx, dy, _daux):
    xin, weight = ctx.saved_tensors
    stride, upsample, in_h, in_w = ctx.meta
    need_dx, need_db, need_dbc, need_dres = ctx.needs
    dy = dy.contiguous()
    # Column sums of dy may already exist: the GroupNorm backward that produced dy emits sum_hw(dx) per sample as
    # a by-product (attribute on the gradient tensor), and sibling convs sharing dy share its bias sum.
    # (the stamps guard against autograd accumulating another gradient into the same tensor in place)
    per_sample = bias_sum = None
    st = getattr(dy, "_pddm_colsum", None)
    if st is not None and st[1] == dy._version:
        per_sample = st[0]
    st = getattr(dy, "_pddm_bias_sum", None)
    if st is not None and st[1] == dy._version:
        bias_sum = st[0]
    if per_sample is not None and bias_sum is None and need_db:
        bias_sum = F.colsum_f32(per_sample)
    want_db = need_db and bias_sum is None
    want_dbc = need_dbc and per_sample is None
    dx, dw, db, dbc = torch.ops.pddm.conv2d_bwd(dy, xin, weight, stride, upsample, in_h, in_w, need_dx, want_db, want_dbc)
    if need_db:
        if want_db:
            dy._pddm_bias_sum = (db, dy._version)
        else:
            db = bias_sum
    if need_dbc and not want_dbc:
        dbc = per_sample
    return (dx if need_dx else None, dw, db if need_db else None, dbc if need_dbc else None,
            dy if need_dres else None, None, None)


torch.library.register_autograd("pddm::conv2d", _conv_backward, setup_context=_conv_setup)


# ================================================================================================ linear
@torch.library.custom_op("pddm::linear", mutates_args=())
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """[M, K] bf16 @ weight[N, K]^T + bias -> fp32 [M, N]  (src/modules/nn.py:36-40), tcgen05 tap-GEMM with one tap."""
    M, K = x.shape
    wp = _packed(weight.view(weight.shape[0], K, 1), 0)
    y = F.tap_gemm(x.view(1, 1, M, K), wp, F.taps_1x1(), 1, 1, M, bias=bias, out_dtype=f32)
    return y.view(M, weight.shape[0])


@linear.register_fake
def _(x, weight, bias):
    return x.new_empty((x.shape[0], weight.shape[0]), dtype=f32)


@torch.library.custom_op("pddm::linear_bwd", mutates_args=())
def linear_bwd(dy: Tensor, x: Tensor, weight: Tensor, need_dx: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """dy fp32 [M, N] -> (dx fp32 [M, K], dweight, dbias)."""
    M, N = dy.shape
    K = x.shape[1]
    dyb = F.convert(dy, bf16)
    dw = F.tap_wgrad(x.view(1, 1, M, K), dyb.view(1, 1, M, N), F.taps_1x1(), 1, 1, M, K, N, (N, K, 1)).view(N, K)
    db = F.colsum(dyb, N)
    dx = _empty(dy)
    if need_dx:
        wp1 = _packed(weight.view(N, K, 1), 1)
        dx = F.tap_gemm(dyb.view(1, 1, M, N), wp1, F.taps_1x1(), 1, 1, M, out_dtype=f32).view(M, K)
    return dx, dw, db


@linear_bwd.register_fake
def _(dy, x, weight, need_dx):
    return (dy.new_empty(x.shape, dtype=f32) if need_dx else dy.new_empty((0,)), weight.new_empty(weight.shape),
            dy.new_empty((weight.shape[0],)))


def _lin_setup(ctx, inputs, output):
    x, weight, bias = inputs
    ctx.save_for_backward(x, weight)
    ctx.need_dx = x.requires_grad
    ctx.has_bias = bias is not None


def _lin_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    dx, dw, db = torch.ops.pddm.linear_bwd(dy.contiguous(), x, weight, ctx.need_dx)
    return (dx if ctx.need_dx else None), dw, (db if ctx.has_bias else None)


torch.library.register_autograd("pddm::linear", _lin_backward, setup_context=_lin_setup)


# ================================================================================================ silu / cast on vectors
@torch.library.custom_op("pddm::silu_vec", mutates_args=())
def silu_vec(x: Tensor) -> Tensor:
    """fp32 [.., N] -> bf16 SiLU (the emb_layers / time_embed activations, src/modules/unet.py:151-157,340-345)."""
    return F.silu_vec(x, bf16)


@silu_vec.register_fake
def _(x):
    return x.new_empty(x.shape, dtype=bf16)


@torch.library.custom_op("pddm::silu_vec_bwd", mutates_args=())
def silu_vec_bwd(x: Tensor, dy: Tensor) -> Tensor:
    return F.silu_vec_bwd(x, F.convert(dy.contiguous(), f32))


@silu_vec_bwd.register_fake
def _(x, dy):
    return x.new_empty(x.shape)


torch.library.register_autograd(
    "pddm::silu_vec", lambda ctx, dy: torch.ops.pddm.silu_vec_bwd(ctx.saved_tensors[0], dy),
    setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


@torch.library.custom_op("pddm::cast_bf16", mutates_args=())
def cast_bf16(x: Tensor) -> Tensor:
    return F.convert(x, bf16) if x.dtype != bf16 else x.clone()


@cast_bf16.register_fake
def _(x):
    return x.new_empty(x.shape, dtype=bf16)


@torch.library.custom_op("pddm::cast_f32", mutates_args=())
def cast_f32(x: Tensor) -> Tensor:
    return F.convert(x, f32) if x.dtype != f32 else x.clone()


@cast_f32.register_fake
def _(x):
    return x.new_empty(x.shape, dtype=f32)


torch.library.register_autograd("pddm::cast_bf16", lambda ctx, g: torch.ops.pddm.cast_f32(g.contiguous()))
torch.library.register_autograd("pddm::cast_f32", lambda ctx, g: torch.ops.pddm.cast_bf16(g.contiguous()))


@torch.library.custom_op("pddm::timestep_embedding", mutates_args=())
def timestep_embedding(t: Tensor, dim: int, max_period: float) -> Tensor:
    """src/modules/nn.py:104-122 -> bf16 [B, dim]."""
    return F.timestep_embedding(t, dim, max_period, bf16)


@timestep_embedding.register_fake
def _(t, dim, max_period):
    return t.new_empty((t.shape[0], dim), dtype=bf16)


# ================================================================================================ group norm (+SiLU)
@torch.library.custom_op("pddm::gn_silu", mutates_args=())
def gn_silu(x: Tensor, gamma: Tensor, beta: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], groups: int,
            eps: float, silu: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """GroupNorm32 (+scale-shift) (+SiLU) on NHWC (src/modules/nn.py:13-20,94-101) -> (y bf16, mean, rstd)."""
    return F.gn_silu_fwd(x, gamma, beta, groups, eps, silu, scale, shift)


@gn_silu.register_fake
def _(x, gamma, beta, scale, shift, groups, eps, silu):
    return x.new_empty(x.shape, dtype=bf16), x.new_empty((x.shape[0], groups), dtype=f32), \
        x.new_empty((x.shape[0], groups), dtype=f32)


@torch.library.custom_op("pddm::gn_silu_bwd", mutates_args=())
def gn_silu_bwd(x: Tensor, dy: Tensor, gamma: Tensor, beta: Tensor, scale: Optional[Tensor], shift: Optional[Tensor],
                mean: Tensor, rstd: Tensor, groups: int, silu: bool
                ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (dx, dgamma, dbeta, dscale, dshift, sum_hw(dx) [B, C])"""
    dx, dg, db, cs, dsc, dsh = F.gn_silu_bwd(x, dy, gamma, beta, mean, rstd, groups, silu, scale, shift,
                                             dx_dtype=x.dtype, want_colsum=True)
    return dx, dg, db, dsc if dsc is not None else _empty(x), dsh if dsh is not None else _empty(x), cs


@gn_silu_bwd.register_fake
def _(x, dy, gamma, beta, scale, shift, mean, rstd, groups, silu):
    e = x.new_empty((0,), dtype=f32)
    ss = x.new_empty(scale.shape, dtype=f32) if scale is not None else e
    return (x.new_empty(x.shape), gamma.new_empty(gamma.shape), gamma.new_empty(gamma.shape), ss, ss,
            x.new_empty((x.shape[0], x.shape[-1]), dtype=f32))


def _gn_setup(ctx, inputs, output):
    x, gamma, beta, scale, shift, groups, eps, silu = inputs
    _, mean, rstd = output
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(x, gamma, beta, mean, rstd, *([scale, shift] if scale is not None else []))
    ctx.meta = (groups, silu, scale is not None)


def _gn_backward(ctx, dy, _dm, _dr):
    groups, silu, has_ss = ctx.meta
    if has_ss:
        x, gamma, beta, mean, rstd, scale, shift = ctx.saved_tensors
    else:
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        scale = shift = None
    dx, dg, db, dsc, dsh, cs = torch.ops.pddm.gn_silu_bwd(x, dy.contiguous(), gamma, beta, scale, shift, mean, rstd,
                                                          groups, silu)
    dx._pddm_colsum = (cs, dx._version)  # consumed by the backward of the conv that produced x (bias / timestep-embedding grads)
    return dx, dg, db, (dsc if has_ss else None), (dsh if has_ss else None), None, None, None


torch.library.register_autograd("pddm::gn_silu", _gn_backward, setup_context=_gn_setup)


# ================================================================================================ silu on feature maps
@torch.library.custom_op("pddm::silu_map", mutates_args=())
def silu_map(x: Tensor) -> Tensor:
    """Stand-alone SiLU on a contiguous bf16 feature map (src/modules/nn.py:13-15 used as in unet.py:146-150)."""
    return F.silu_map(x)


@silu_map.register_fake
def _(x):
    return torch.empty_like(x)


@torch.library.custom_op("pddm::silu_map_bwd", mutates_args=())
def silu_map_bwd(x: Tensor, dy: Tensor) -> Tensor:
    return F.silu_map_bwd(x, dy.contiguous())


@silu_map_bwd.register_fake
def _(x, dy):
    return torch.empty_like(x)


torch.library.register_autograd(
    "pddm::silu_map", lambda ctx, dy: torch.ops.pddm.silu_map_bwd(ctx.saved_tensors[0], dy),
    setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


# ================================================================================================ attention
@torch.library.custom_op("pddm::attention", mutates_args=())
def attention(qkv: Tensor, heads: int) -> Tuple[Tensor, Tensor]:
    """QKVAttention (src/modules/unet.py:237-256) on bf16 [B, T, 3C] -> (out [B, T, C], lse [B, heads, T])."""
    return F.attn_fwd(qkv, heads)


@attention.register_fake
def _(qkv, heads):
    B, T, C3 = qkv.shape
    return qkv.new_empty((B, T, C3 // 3)), qkv.new_empty((B, heads, T), dtype=f32)


@torch.library.custom_op("pddm::attention_bwd", mutates_args=())
def attention_bwd(qkv: Tensor, out: Tensor, dout: Tensor, lse: Tensor, heads: int) -> Tensor:
    return F.attn_bwd(qkv, out, dout, lse, heads)


@attention_bwd.register_fake
def _(qkv, out, dout, lse, heads):
    return qkv.new_empty(qkv.shape)


def _attn_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(inputs[0], output[0], output[1])
    ctx.heads = inputs[1]


def _attn_backward(ctx, dout, _dlse):
    qkv, out, lse = ctx.saved_tensors
    return torch.ops.pddm.attention_bwd(qkv, out, dout.contiguous(), lse, ctx.heads), None


torch.library.register_autograd("pddm::attention", _attn_backward, setup_context=_attn_setup)


# ================================================================================================ concat / add
@torch.library.custom_op("pddm::concat_channels", mutates_args=())
def concat_channels(a: Tensor, b: Tensor) -> Tensor:
    """th.cat([h, skip], dim=1) on NHWC bf16 (src/modules/unet.py:492)."""
    return F.concat_channels(a, b)


@concat_channels.register_fake
def _(a, b):
    return a.new_empty(a.shape[:-1] + (a.shape[-1] + b.shape[-1],))


@torch.library.custom_op("pddm::split_channels", mutates_args=())
def split_channels(x: Tensor, c1: int) -> Tuple[Tensor, Tensor]:
    return F.split_channels(x, c1)


@split_channels.register_fake
def _(x, c1):
    return x.new_empty(x.shape[:-1] + (c1,)), x.new_empty(x.shape[:-1] + (x.shape[-1] - c1,))


torch.library.register_autograd(
    "pddm::concat_channels", lambda ctx, g: tuple(torch.ops.pddm.split_channels(g.contiguous(), ctx.c1)),
    setup_context=lambda ctx, inputs, output: setattr(ctx, "c1", inputs[0].shape[-1]))


@torch.library.custom_op("pddm::add", mutates_args=())
def add(a: Tensor, b: Tensor) -> Tensor:
    """bf16 elementwise add (gradient fan-in of tensors that feed two consumers)."""
    return F.add_bf16(a, b)


@add.register_fake
def _(a, b):
    return a.new_empty(a.shape)


torch.library.register_autograd("pddm::add", lambda ctx, g: (g, g))


# ================================================================================================ stem / head
@torch.library.custom_op("pddm::stem_conv", mutates_args=())
def stem_conv(x: Tensor, weight: Tensor, bias: Tensor) -> Tuple[Tensor, Tensor]:
    """First conv (Cin <= 4): NCHW fp32 model input -> NHWC bf16 (src/modules/unet.py:353); also returns the im2col
    patch matrix the tensor-core GEMM read (needed by the weight gradient)."""
    return F.stem_conv_fwd(x, weight, bias)


@stem_conv.register_fake
def _(x, weight, bias):
    B, Cin, H, W = x.shape
    return x.new_empty((B, H, W, weight.shape[0]), dtype=bf16), x.new_empty((B, H, W, (Cin * 9 + 31) // 32 * 32), dtype=bf16)


@torch.library.custom_op("pddm::stem_conv_bwd", mutates_args=())
def stem_conv_bwd(patches: Tensor, dy: Tensor, weight: Tensor) -> Tuple[Tensor, Tensor]:
    return F.stem_conv_wgrad(patches, dy, tuple(weight.shape))


@stem_conv_bwd.register_fake
def _(patches, dy, weight):
    return weight.new_empty(weight.shape), weight.new_empty((weight.shape[0],))


def _stem_backward(ctx, dy, _dp):
    patches, w = ctx.saved_tensors
    dw, db = torch.ops.pddm.stem_conv_bwd(patches, dy.contiguous(), w)
    return None, dw, db  # the model input needs no gradient (x_t is data)


def _stem_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(output[1], inputs[1])


torch.library.register_autograd("pddm::stem_conv", _stem_backward, setup_context=_stem_setup)


@torch.library.custom_op("pddm::head_conv", mutates_args=())
def head_conv(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """Last conv (Cout <= 8): NHWC bf16 -> NCHW fp32 model output (src/modules/unet.py:440)."""
    return F.head_conv_fwd(x, weight, bias)


@head_conv.register_fake
def _(x, weight, bias):
    return x.new_empty((x.shape[0], weight.shape[0], x.shape[1], x.shape[2]), dtype=f32)


@torch.library.custom_op("pddm::head_conv_bwd", mutates_args=())
def head_conv_bwd(x: Tensor, weight: Tensor, dy: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    return F.head_conv_bwd(x, weight, dy)


@head_conv_bwd.register_fake
def _(x, weight, dy):
    return x.new_empty(x.shape), weight.new_empty(weight.shape), weight.new_empty((weight.shape[0],))


def _head_backward(ctx, dy):
    x, w = ctx.saved_tensors
    return tuple(torch.ops.pddm.head_conv_bwd(x, w, dy.contiguous()))


torch.library.register_autograd("pddm::head_conv", _head_backward,
                                setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0], inputs[1]))


# ================================================================================================ layout at the seam
@torch.library.custom_op("pddm::to_nhwc", mutates_args=())
def to_nhwc(x: Tensor) -> Tensor:
    """NCHW fp32 -> NHWC bf16."""
    return F.nchw_to_nhwc(x.contiguous(), bf16)


@to_nhwc.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[2], x.shape[3], x.shape[1]), dtype=bf16)


@torch.library.custom_op("pddm::to_nchw", mutates_args=())
def to_nchw(x: Tensor) -> Tensor:
    """NHWC bf16/fp32 -> NCHW fp32."""
    return F.nhwc_to_nchw(x.contiguous())


@to_nchw.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[3], x.shape[1], x.shape[2]), dtype=f32)


torch.library.register_autograd("pddm::to_nhwc", lambda ctx, g: torch.ops.pddm.to_nchw(g))
torch.library.register_autograd("pddm::to_nchw", lambda ctx, g: torch.ops.pddm.to_nhwc(g))


# ================================================================================================ losses
@torch.library.custom_op("pddm::simple_loss", mutates_args=())
def simple_loss(pred: Tensor, noise: Tensor) -> Tensor:
    """per-sample L_simple = mean_flat((noise - pred)^2)  (src/engine.py:266) -> fp32 [B]."""
    return F.sq_err(pred.contiguous(), noise.contiguous())[0]


@simple_loss.register_fake
def _(pred, noise):
    return pred.new_empty((pred.shape[0],))


@torch.library.custom_op("pddm::simple_loss_bwd", mutates_args=())
def simple_loss_bwd(pred: Tensor, noise: Tensor, g: Tensor) -> Tensor:
    return F.sq_err(pred, noise, g.float().contiguous(), want_grad=True)[1]


@simple_loss_bwd.register_fake
def _(pred, noise, g):
    return pred.new_empty(pred.shape)


torch.library.register_autograd(
    "pddm::simple_loss", lambda ctx, g: (torch.ops.pddm.simple_loss_bwd(*ctx.saved_tensors, g), None),
    setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0].contiguous(), inputs[1].contiguous()))


class HybridLoss(torch.autograd.Function):
    """per-sample L_simple + vb_weight * L_vlb with learned variance: model_out = [eps | v];
    the variational term sees eps through a stop-gradient, so only v receives its gradient."""

    @staticmethod
    def forward(ctx, model_out, noise, x0, x_t, t, tabs, vb_weight):
        model_out = model_out.contiguous()
        per, _ = F.sq_err(model_out, noise)
        vb, gv = F.vlb_terms(x0, x_t, model_out, t, tabs, mode=1, want_grad_v=True)
        ctx.save_for_backward(model_out, noise, gv)
        ctx.set_materialize_grads(False)
        ctx.vb_weight = vb_weight
        return per + vb_weight * vb, vb

    @staticmethod
    def backward(ctx, g, _gvb):
        model_out, noise, gv = ctx.saved_tensors
        _, grad = F.sq_err(model_out, noise, g.float().contiguous(), want_grad=True, grad_v_unit=gv,
                           v_scale=ctx.vb_weight)
        return grad, None, None, None, None, None, None
