"""Host-side launch helpers: torch tensors in, C-ABI kernel calls out (no autograd here; see ops.py).

Activation convention inside the network: contiguous ``[B, H, W, C]`` (NHWC) tensors, bf16 unless noted.
Every function only enqueues work on the current CUDA stream (capturable in a CUDA graph).
"""
import ctypes as C
import math

import torch

from . import _lib as L

bf16 = torch.bfloat16
f32 = torch.float32


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _chk(t, dtype=None):
    if not t.is_cuda:
        raise RuntimeError("pddm_b200 ops need CUDA tensors (sm_100a); there is no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError("pddm_b200 ops need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t


# ------------------------------------------------------------------------------------------ tap tables
def taps_3x3():
    return [(0, r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)]


def taps_1x1():
    return [(0, 0, 0, 0)]


def taps_stride2(B):
    """3x3 / stride 2 / pad 1 over the phase-split input [4B, H/2, W/2, C] (phase p = 2*(row parity)+(col parity))."""
    out = []
    for r in range(3):
        for s in range(3):
            a, dh = (0, 0) if r == 1 else (1, -1 if r == 0 else 0)
            b, dw = (0, 0) if s == 1 else (1, -1 if s == 0 else 0)
            out.append(((a * 2 + b) * B, dh, dw, r * 3 + s))
    return out


def taps_stride2_dgrad(a, b):
    """Taps of the stride-2 data gradient for output phase (a, b): dX[2i+a, 2j+b] = sum dY[i+di, j+dj] W[r, s];
    weight slots refer to the mode-1 (flipped) pack, where tap (r, s) lives in slot 8 - (3r + s)."""
    rows = [(1, 0)] if a == 0 else [(0, 1), (2, 0)]
    cols = [(1, 0)] if b == 0 else [(0, 1), (2, 0)]
    return [(0, di, dj, 8 - (r * 3 + s)) for r, di in rows for s, dj in cols]


def _fill_taps(p, taps):
    p.ntaps = len(taps)
    for i, (db, dh, dw, slot) in enumerate(taps):
        p.tap_db[i], p.tap_dh[i], p.tap_dw[i] = db, dh, dw
        if hasattr(p, "tap_w"):
            p.tap_w[i] = slot


# ------------------------------------------------------------------------------------------ conv / linear
PACK_HOOK = [None]  # set by ops.WeightArena: (w, mode, cout_pad, cin_pad) -> packed tensor | None


def pack_weight(w, mode, cout_pad=None, cin_pad=None):
    """fp32 [Cout, Cin, *k] -> bf16 GEMM operand (optionally zero-padded).
    mode 0: [Cout_pad, taps, Cin_pad]; mode 1 (dgrad): [Cin_pad, taps(flipped), Cout_pad]."""
    L.require_device(w)
    _chk(w, f32)
    cout, cin = w.shape[0], w.shape[1]
    ntaps = w.numel() // (cout * cin)
    cop, cip = cout_pad or cout, cin_pad or cin
    if PACK_HOOK[0] is not None:
        hit = PACK_HOOK[0](w, mode, cop, cip, ntaps)
        if hit is not None:
            return hit
    dst = torch.empty((cop, ntaps, cip) if mode == 0 else (cip, ntaps, cop), dtype=bf16, device=w.device)
    L.call("pddm_pack_conv_weight", L.ptr(w), L.ptr(dst), cout, cin, ntaps, mode, cop, cip, L.stream())
    return dst


def tap_gemm(x, wp, taps, B, H, W, *, bias=None, bcast=None, residual=None, out=None, out_dtype=bf16,
             out_hw=None, out_map=(1, 1, 0, 0), cin=None, x2=None):
    """y[b, oh, ow, :] = sum_taps x[b+db, h+dh, w+dw, :] @ wp[:, slot, :]^T (+bias +bcast[b] +residual).

    x: bf16 [NB, H, W, ldx] (cin <= ldx leading channels used; channel-slice views allowed); wp: bf16
    [Cout, slots, cin].  ``x2`` (bf16 [NB, H, W, c2]): second source -- the GEMM's input channels are
    [x[..., :cin - c2] | x2] without materialising the concat (src/modules/unet.py:492)."""
    L.require_device(x)
    if x.dtype != bf16 or wp.dtype != bf16:
        raise TypeError("tap_gemm operands must be bf16")
    _chk(wp, bf16)
    NB, ldx = x.shape[0], _ld(x, "x")
    cout, slots, wcin = wp.shape
    cin = wcin if cin is None else cin
    assert wcin == cin and x.shape[1] == H and x.shape[2] == W
    oH, oW = out_hw if out_hw is not None else (H, W)
    if out is None:
        out = torch.empty((B, oH, oW, cout), dtype=out_dtype, device=x.device)
    p = L.ConvParams()
    p.x, p.w, p.y = L.ptr(x), L.ptr(wp), L.ptr(out)
    if x2 is not None:
        assert x2.dtype == bf16 and x2.shape[:3] == x.shape[:3] and x.shape[-1] + x2.shape[-1] == cin
        p.x2, p.Cin_a, p.ldx2 = L.ptr(x2), x.shape[-1], _ld(x2, "x2")
    p.bias = L.ptr(_chk(bias, f32)) if bias is not None else None
    if bcast is not None:  # [B, Cout] fp32, rows may be strided (column slice of a wider matrix)
        assert bcast.dtype == f32 and bcast.is_cuda and bcast.stride(-1) == 1 and bcast.shape[-1] == cout
        p.bcast, p.ld_bcast = L.ptr(bcast), bcast.stride(0)
    if residual is not None:
        _chk(residual)
        assert residual.shape == out.shape
        p.residual, p.res_dtype = L.ptr(residual), L.dt(residual)
    p.y_dtype = L.dt(out)
    p.x_NB, p.B, p.H, p.W, p.Cin, p.ldx, p.Cout = NB, B, H, W, cin, ldx, cout
    _fill_taps(p, taps)
    p.w_ntaps = slots
    p.out_H, p.out_W = oH, oW
    p.out_sh, p.out_sw, p.out_oh, p.out_ow = out_map
    L.call("pddm_conv2d_fwd", C.byref(p), L.stream())
    return out


def set_sm_reserve(n):
    """Leave ``n`` SMs free in subsequent launches of the persistent kernels (room for a concurrent collective)."""
    L.check(L.load().pddm_set_sm_reserve(int(n)), "pddm_set_sm_reserve")


def wgrad_workspace_bytes(B, H, W, cin, cout, ntaps, x_NB=None):
    """Bytes of split-K workspace ``tap_wgrad`` needs for this shape (host-side query, no launch)."""
    p = L.WgradParams()
    p.x_NB, p.B, p.H, p.W, p.Cin, p.ldx, p.Cout, p.lddy = x_NB or B, B, H, W, cin, cin, cout, cout
    p.ntaps = ntaps
    return int(L.load().pddm_conv2d_wgrad_workspace(C.byref(p)))


def tap_wgrad(x, dy, taps, B, H, W, cin, cout, w_shape, accumulate_into=None, launch_stream=None, keep=None,
              out=None, ws=None, dw_ldc=0, dw_c0=0):
    """dw[n, c, tap] = sum_pixels dy[b,h,w,n] * x[b+db, h+dh, w+dw, c]  -> fp32 tensor of shape w_shape.

    ``launch_stream``: enqueue on that stream instead of the current one (buffers are still allocated from the
    current stream's pool; the caller keeps them alive through ``keep`` until it has joined the streams).
    ``out``: write the gradient there (a parameter-shaped fp32 tensor, e.g. a slice of a flat gradient arena);
    with ``dw_ldc`` / ``dw_c0`` it is the input-channel slice [dw_c0, dw_c0 + cin) of a parameter with dw_ldc input
    channels.  ``ws``: caller-owned split-K workspace (uint8)."""
    L.require_device(x)
    if x.dtype != bf16 or dy.dtype != bf16:
        raise TypeError("tap_wgrad operands must be bf16")
    dw = out if out is not None else (accumulate_into if accumulate_into is not None
                                      else torch.empty(w_shape, dtype=f32, device=x.device))
    p = L.WgradParams()
    p.x, p.dy, p.dw = L.ptr(x), L.ptr(dy), L.ptr(dw)
    p.x_NB, p.B, p.H, p.W, p.Cin, p.ldx, p.Cout, p.lddy = x.shape[0], B, H, W, cin, _ld(x, "x"), cout, _ld(dy, "dy")
    _fill_taps(p, taps)
    p.dw_layout = 1
    p.accumulate = 1 if accumulate_into is not None else 0
    p.dw_ldc, p.dw_c0 = dw_ldc, dw_c0
    nbytes = L.load().pddm_conv2d_wgrad_workspace(C.byref(p))
    if ws is None:
        ws = _ws(nbytes, x.device)
    elif ws.numel() < nbytes:
        raise RuntimeError(f"tap_wgrad: workspace of {ws.numel()} bytes, {nbytes} needed")
    st = L.stream() if launch_stream is None else C.c_void_p(launch_stream.cuda_stream)
    L.call("pddm_conv2d_wgrad", C.byref(p), L.ptr(ws), C.c_size_t(ws.numel()), st)
    if keep is not None:
        keep.extend((ws, x, dy))  # NOT dw: an extra reference would make autograd's AccumulateGrad clone it (on the
        # main stream, before the side-stream GEMM has written it) instead of adopting it as .grad
    return dw


def colsum_rows(x, out, accumulate=False):
    """out[b, c] (+)= sum_hw x[b, hw, c]  (x bf16 [B, ..., C] or a channel-slice view; out fp32 [B, >=C] view)."""
    if x.dtype != bf16 or out.dtype != f32:
        raise TypeError("colsum_rows: bf16 in, fp32 out")
    B, C_ = x.shape[0], x.shape[-1]
    L.call("pddm_colsum_rows", L.ptr(x), _ld(x, "x"), B, x.numel() // (B * C_), C_, L.ptr(out), out.stride(0),
           1 if accumulate else 0, L.stream())
    return out


def batch_fold(ps, src_of, n, dst):
    """dst[j] = sum_b ps[b, src_of[j]], j < n (ps fp32 [B, ld]; src_of int32 device tensor or None = identity)."""
    L.call("pddm_batch_fold", L.ptr(ps), ps.stride(0), ps.shape[0], L.ptr(src_of), n, L.ptr(dst), L.stream())
    return dst


def convert_rows(src, dst):
    """dst (bf16 [rows, cols]) = src (fp32 [rows, cols]), either may be a column-slice view of a wider matrix."""
    assert src.dtype == f32 and dst.dtype == bf16 and src.shape == dst.shape and src.dim() == 2
    L.call("pddm_convert_rows", L.ptr(src), src.stride(0), L.ptr(dst), dst.stride(0), src.shape[0], src.shape[1],
           L.stream())
    return dst


def colsum(x2d_bf16, C_, out=None):
    """sum over all leading dims of a bf16 [..., C] tensor -> fp32 [C]."""
    _chk(x2d_bf16, bf16)
    out = torch.empty(C_, dtype=f32, device=x2d_bf16.device) if out is None else out
    M = x2d_bf16.numel() // x2d_bf16.shape[-1]
    ws = _ws(L.load().pddm_colsum_workspace(C.c_int64(M), 1, C_), x2d_bf16.device)
    L.call("pddm_colsum", L.ptr(x2d_bf16), x2d_bf16.shape[-1], C.c_int64(M), C_, L.ptr(out), 0, L.ptr(ws),
           C.c_size_t(ws.numel()), L.stream())
    return out


def colsum_f32(x):
    """fp32 [M, C] -> [C] (deterministic)."""
    _chk(x, f32)
    out = torch.empty(x.shape[-1], dtype=f32, device=x.device)
    L.call("pddm_colsum_f32", L.ptr(x), x.numel() // x.shape[-1], x.shape[-1], L.ptr(out), L.stream())
    return out


def colsum_per_sample(x):
    """bf16 [B, ..., C] -> fp32 [B, C]."""
    _chk(x, bf16)
    B, C_ = x.shape[0], x.shape[-1]
    out = torch.empty((B, C_), dtype=f32, device=x.device)
    HW = x.numel() // (B * C_)
    ws = _ws(L.load().pddm_colsum_workspace(C.c_int64(HW), B, C_), x.device)
    L.call("pddm_colsum_per_sample", L.ptr(x), B, HW, C_, L.ptr(out), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    return out


# ------------------------------------------------------------------------------------------ layout helpers
def nchw_to_nhwc(x, dtype=bf16):
    L.require_device(x)
    _chk(x, f32)
    B, C_, H, W = x.shape
    y = torch.empty((B, H, W, C_), dtype=dtype, device=x.device)
    L.call("pddm_nchw_to_nhwc", L.ptr(x), L.ptr(y), L.dt(y), B, C_, H * W, L.stream())
    return y


def nhwc_to_nchw(x):
    L.require_device(x)
    _chk(x)
    B, H, W, C_ = x.shape
    y = torch.empty((B, C_, H, W), dtype=f32, device=x.device)
    L.call("pddm_nhwc_to_nchw", L.ptr(x), L.dt(x), L.ptr(y), B, C_, H * W, L.stream())
    return y


def copy_channels(src, src_off, dst, dst_off, nch):
    M = src.numel() // src.shape[-1]
    L.call("pddm_copy_channels", L.ptr(src), src.shape[-1], src_off, L.ptr(dst), dst.shape[-1], dst_off,
           C.c_int64(M), nch, L.stream())


def concat_channels(a, b):
    """th.cat([a, b], dim=channels) for NHWC bf16 (src/modules/unet.py:492)."""
    _chk(a, bf16)
    _chk(b, bf16)
    out = torch.empty(a.shape[:-1] + (a.shape[-1] + b.shape[-1],), dtype=bf16, device=a.device)
    copy_channels(a, 0, out, 0, a.shape[-1])
    copy_channels(b, 0, out, a.shape[-1], b.shape[-1])
    return out


def split_channels(x, c1):
    a = torch.empty(x.shape[:-1] + (c1,), dtype=bf16, device=x.device)
    b = torch.empty(x.shape[:-1] + (x.shape[-1] - c1,), dtype=bf16, device=x.device)
    copy_channels(x, 0, a, 0, c1)
    copy_channels(x, c1, b, 0, x.shape[-1] - c1)
    return a, b


def upsample2x(x):
    B, H, W, C_ = x.shape
    y = torch.empty((B, 2 * H, 2 * W, C_), dtype=bf16, device=x.device)
    L.call("pddm_upsample2x", L.ptr(_chk(x, bf16)), L.ptr(y), B, H, W, C_, L.stream())
    return y


def upsample2x_bwd(g):
    B, H2, W2, C_ = g.shape
    y = torch.empty((B, H2 // 2, W2 // 2, C_), dtype=bf16, device=g.device)
    L.call("pddm_upsample2x_bwd", L.ptr(_chk(g, bf16)), L.ptr(y), B, H2 // 2, W2 // 2, C_, L.stream())
    return y


def phase_split(x):
    B, H, W, C_ = x.shape
    y = torch.empty((4 * B, H // 2, W // 2, C_), dtype=bf16, device=x.device)
    L.call("pddm_phase_split", L.ptr(_chk(x, bf16)), L.ptr(y), B, H, W, C_, L.stream())
    return y


def add_bf16(a, b):
    y = torch.empty_like(a)
    L.call("pddm_add_bf16", L.ptr(_chk(a, bf16)), L.ptr(_chk(b, bf16)), L.ptr(y), C.c_int64(a.numel()), L.stream())
    return y


def convert(x, dtype):
    """fp32 <-> bf16 elementwise conversion (our own kernel: no torch math on the hot path)."""
    if x.dtype == dtype:
        return x
    _chk(x)
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    L.call("pddm_convert", L.ptr(x), L.dt(x), L.ptr(y), L.dt(y), C.c_int64(x.numel()), L.stream())
    return y


def silu_vec(x, dtype=bf16):
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    L.call("pddm_silu", L.ptr(_chk(x, f32)), L.ptr(y), L.dt(y), C.c_int64(x.numel()), L.stream())
    return y


def silu_vec_bwd(x, dy):
    dx = torch.empty_like(x)
    L.call("pddm_silu_bwd", L.ptr(_chk(x, f32)), L.ptr(_chk(dy, f32)), L.ptr(dx), C.c_int64(x.numel()), L.stream())
    return dx


def silu_map(x):
    """SiLU on a contiguous bf16 feature map (any shape, numel % 8 == 0)."""
    L.require_device(x)
    y = torch.empty_like(x)
    L.call("pddm_silu_map", L.ptr(_chk(x, bf16)), L.ptr(y), C.c_int64(x.numel()), L.stream())
    return y


def silu_map_bwd(x, dy):
    dx = torch.empty_like(x)
    L.call("pddm_silu_map_bwd", L.ptr(_chk(x, bf16)), L.ptr(_chk(dy, bf16)), L.ptr(dx), C.c_int64(x.numel()), L.stream())
    return dx


def timestep_embedding(t, dim, max_period=10000, dtype=bf16):
    """src/modules/nn.py:104-122; t int64 or float32 [B]."""
    L.require_device(t)
    if t.dtype not in (torch.int64, torch.float32):
        t = t.float() if t.is_floating_point() else t.long()
    t = t.contiguous()
    out = torch.empty((t.shape[0], dim), dtype=dtype, device=t.device)
    L.call("pddm_timestep_embedding", L.ptr(t), 1 if t.dtype == torch.float32 else 0, L.ptr(out), L.dt(out),
           t.shape[0], dim, C.c_float(float(max_period)), L.stream())
    return out


# ------------------------------------------------------------------------------------------ group norm
def _ld(t, what="tensor"):
    """Row stride (elements) of a [B, ..., C] activation that may be a channel-slice view of a wider NHWC tensor."""
    if not t.is_cuda:
        raise RuntimeError("pddm_b200 ops need CUDA tensors (sm_100a); there is no CPU fallback")
    if t.stride(-1) != 1:
        raise RuntimeError(f"{what}: channels must be innermost")
    ld = t.stride(-2) if t.dim() >= 2 else t.shape[-1]
    for i in range(t.dim() - 2):
        if t.shape[i] != 1 and t.stride(i) != t.stride(i + 1) * t.shape[i + 1]:
            raise RuntimeError(f"{what}: only channel-slice views of a contiguous NHWC tensor are supported")
    return int(ld)


def gn_pipe_slots(B, HW, C, groups, C_a=0, ntens=1):
    """> 0 if the persistent bulk-tensor GroupNorm kernel takes this shape (ntens: 1 fwd, 2 bwd, 3 bwd with gres)."""
    return int(L.load().pddm_gn_pipe_slots(B, HW, C, groups, C_a, ntens))


def gn_silu_fwd(x, gamma, beta, groups=32, eps=1e-5, silu=True, scale=None, shift=None, x2=None, out=None):
    """x: [B, ..., C_a] bf16/fp32 (+ optional second source x2 [B, ..., C - C_a]: the concat-free th.cat of
    src/modules/unet.py:492) -> (y bf16 [B, ..., C], mean [B,G], rstd [B,G])."""
    L.require_device(x)
    B, C_a = x.shape[0], x.shape[-1]
    C_ = C_a + (x2.shape[-1] if x2 is not None else 0)
    HW = x.numel() // (B * C_a)
    y = out if out is not None else torch.empty(x.shape[:-1] + (C_,), dtype=bf16, device=x.device)
    mean = torch.empty((B, groups), dtype=f32, device=x.device)
    rstd = torch.empty((B, groups), dtype=f32, device=x.device)
    p = L.GnFwdParams()
    p.x, p.x_dtype, p.gamma, p.beta, p.y, p.mean, p.rstd = L.ptr(x), L.dt(x), L.ptr(gamma), L.ptr(beta), L.ptr(y), \
        L.ptr(mean), L.ptr(rstd)
    p.ldx, p.ldy = _ld(x, "x"), _ld(y, "y")
    if x2 is not None:
        assert x2.dtype == x.dtype and x2.shape[:-1] == x.shape[:-1]
        p.x2, p.C_a, p.ldx2 = L.ptr(x2), C_a, _ld(x2, "x2")
    if scale is not None:
        p.scale, p.shift, p.ld_ss = L.ptr(scale), L.ptr(shift), scale.stride(0)
    p.B, p.HW, p.C, p.G, p.eps, p.silu = B, HW, C_, groups, eps, 1 if silu else 0
    ws = _ws(L.load().pddm_gn_silu_fwd_workspace(B, groups), x.device)
    L.call("pddm_gn_silu_fwd", C.byref(p), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    return y, mean, rstd


def gn_silu_bwd(x, dy, gamma, beta, mean, rstd, groups=32, silu=True, scale=None, shift=None, dx_dtype=bf16,
                want_colsum=False, x2=None, gres=None, dx=None, dx2=None, dx_accumulate=False, dx2_accumulate=False,
                part_dgamma=None, part_dbeta=None, colsum=None, colsum2=None, colsum_accumulate=False,
                colsum2_accumulate=False):
    """-> dx, dgamma, dbeta, dx_colsum [B,C] | None, dscale, dshift.

    Extensions of the persistent kernel (bf16, no scale-shift; see include/pddm.h): ``x2`` second source, ``gres`` a
    gradient added to dx on the way out, ``dx`` / ``dx2`` caller-provided destinations for channels [0, C_a) /
    [C_a, C) (dx alone: one tensor of all C channels), ``*_accumulate`` add into the destination, ``part_dgamma`` /
    ``part_dbeta`` [B, >=C] per-sample partial sums instead of the batch-reduced dgamma / dbeta (returned as None),
    ``colsum`` / ``colsum2`` destinations of sum_hw dx per segment."""
    L.require_device(x)
    if dy.dtype != bf16:
        raise TypeError(f"expected {bf16}, got {dy.dtype}")
    B, C_a = x.shape[0], x.shape[-1]
    C_ = C_a + (x2.shape[-1] if x2 is not None else 0)
    HW = x.numel() // (B * C_a)
    dev = x.device
    if dx is None:
        if x2 is not None and dx2 is None:
            dx2 = torch.empty(x2.shape, dtype=dx_dtype, device=dev)
        dx = torch.empty(x.shape if dx2 is not None else x.shape[:-1] + (C_,), dtype=dx_dtype, device=dev)
    partials = part_dgamma is not None
    dgamma = dbeta = None
    if not partials:
        dgamma = torch.empty(C_, dtype=f32, device=dev)
        dbeta = torch.empty(C_, dtype=f32, device=dev)
    if colsum is None and want_colsum:
        colsum = torch.empty((B, C_a if colsum2 is not None else C_), dtype=f32, device=dev)
    dscale = dshift = None
    p = L.GnBwdParams()
    p.x, p.x_dtype, p.dy, p.gamma, p.beta, p.mean, p.rstd = L.ptr(x), L.dt(x), L.ptr(dy), L.ptr(gamma), L.ptr(beta), \
        L.ptr(mean), L.ptr(rstd)
    p.ldx, p.lddy, p.ld_dx = _ld(x, "x"), _ld(dy, "dy"), _ld(dx, "dx")
    if x2 is not None:
        p.x2, p.C_a, p.ldx2 = L.ptr(x2), C_a, _ld(x2, "x2")
    if gres is not None:
        assert gres.dtype == bf16 and gres.shape[-1] == C_
        p.gres, p.ld_gres = L.ptr(gres), _ld(gres, "gres")
    if dx2 is not None:
        p.dx2, p.ld_dx2 = L.ptr(dx2), _ld(dx2, "dx2")
    p.dx_accumulate, p.dx2_accumulate = int(bool(dx_accumulate)), int(bool(dx2_accumulate))
    if partials:
        assert part_dbeta is not None and part_dgamma.stride(0) == part_dbeta.stride(0)
        p.part_dgamma, p.part_dbeta, p.ld_part = L.ptr(part_dgamma), L.ptr(part_dbeta), part_dgamma.stride(0)
    if scale is not None:
        dscale = torch.empty((B, C_), dtype=f32, device=dev)
        dshift = torch.empty((B, C_), dtype=f32, device=dev)
        p.scale, p.shift, p.ld_ss = L.ptr(scale), L.ptr(shift), scale.stride(0)
        # dscale/dshift are written with the same leading dimension as scale/shift
        if scale.stride(0) != C_:
            dscale = torch.empty((B, scale.stride(0)), dtype=f32, device=dev)
            dshift = torch.empty((B, scale.stride(0)), dtype=f32, device=dev)
        p.dscale, p.dshift = L.ptr(dscale), L.ptr(dshift)
    p.dx, p.dx_dtype, p.dgamma, p.dbeta = L.ptr(dx), L.dt(dx), L.ptr(dgamma), L.ptr(dbeta)
    if colsum is not None:
        p.dx_colsum, p.ld_colsum = L.ptr(colsum), colsum.stride(0)
    if colsum2 is not None:
        p.dx_colsum2, p.ld_colsum2 = L.ptr(colsum2), colsum2.stride(0)
    p.colsum_accumulate, p.colsum2_accumulate = int(bool(colsum_accumulate)), int(bool(colsum2_accumulate))
    p.B, p.HW, p.C, p.G, p.silu = B, HW, C_, groups, 1 if silu else 0
    ws = _ws(L.load().pddm_gn_silu_bwd_workspace(B, C_), dev) if not partials else None
    L.call("pddm_gn_silu_bwd", C.byref(p), L.ptr(ws), C.c_size_t(ws.numel() if ws is not None else 0), L.stream())
    return (dx if dx2 is None else (dx, dx2)), dgamma, dbeta, colsum, dscale, dshift


# ------------------------------------------------------------------------------------------ attention
def attn_fwd(qkv, heads):
    """qkv: bf16 [B, T, 3C] head-major [q|k|v] packing (src/modules/unet.py:230) -> (out [B,T,C], lse [B,heads,T])."""
    L.require_device(qkv)
    _chk(qkv, bf16)
    B, T, C3 = qkv.shape
    Cc = C3 // 3
    out = torch.empty((B, T, Cc), dtype=bf16, device=qkv.device)
    lse = torch.empty((B, heads, T), dtype=f32, device=qkv.device)
    p = L.AttnFwdParams()
    p.qkv, p.out, p.lse, p.B, p.T, p.heads, p.d = L.ptr(qkv), L.ptr(out), L.ptr(lse), B, T, heads, Cc // heads
    L.call("pddm_attn_fwd", C.byref(p), L.stream())
    return out, lse


def attn_bwd(qkv, out, dout, lse, heads):
    _chk(dout, bf16)
    B, T, C3 = qkv.shape
    dqkv = torch.empty_like(qkv)
    p = L.AttnBwdParams()
    p.qkv, p.out, p.dout, p.lse, p.dqkv = L.ptr(qkv), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(dqkv)
    p.B, p.T, p.heads, p.d = B, T, heads, C3 // 3 // heads
    nws = L.load().pddm_attn_bwd_workspace_bytes(B, T, heads, p.d)  # two key tiles exchange fp32 partial dQ products
    ws = _ws(nws, qkv.device) if nws else None
    p.ws, p.ws_bytes = (L.ptr(ws) if nws else None), nws
    L.call("pddm_attn_bwd", C.byref(p), L.stream())
    return dqkv


# ------------------------------------------------------------------------------------------ thin convs
def _pad32(n):
    return (n + 31) // 32 * 32


def im2col3x3(x_nchw):
    """NCHW fp32 model input -> [B, H, W, Kp] bf16 patch matrix (k = ci*9 + tap, Kp = Cin*9 rounded up to 32)."""
    L.require_device(x_nchw)
    _chk(x_nchw, f32)
    B, Cin, H, W = x_nchw.shape
    Kp = _pad32(Cin * 9)
    out = torch.empty((B, H, W, Kp), dtype=bf16, device=x_nchw.device)
    L.call("pddm_im2col3x3", L.ptr(x_nchw), L.ptr(out), B, Cin, H, W, Kp, L.stream())
    return out


def stem_conv_fwd(x_nchw, w, bias):
    """conv3x3(Cin<=4) as a K=32 tensor-core GEMM over the im2col patches -> (y NHWC bf16, patches)."""
    patches = im2col3x3(x_nchw)
    B, H, W, Kp = patches.shape
    cout = w.shape[0]
    wp = pack_weight(w.view(cout, -1, 1), 0, cin_pad=Kp)  # [Cout, 1, Kp]
    return tap_gemm(patches, wp, taps_1x1(), B, H, W, bias=bias), patches


def stem_conv_wgrad(patches, dy, w_shape):
    B, H, W, Kp = patches.shape
    cout, k = w_shape[0], w_shape[1] * 9
    dwp = tap_wgrad(patches, dy, taps_1x1(), B, H, W, Kp, cout, (cout, Kp, 1))
    return dwp.view(cout, Kp)[:, :k].reshape(w_shape), colsum(dy, cout)


def head_conv_fwd(x, w, bias):
    """conv3x3 to Cout<=8 channels with the output channels zero-padded to 8; result NCHW fp32."""
    L.require_device(x)
    _chk(x, bf16)
    B, H, W, Cin = x.shape
    cout = w.shape[0]
    wp = pack_weight(w, 0, cout_pad=8)
    bias_p = None
    if bias is not None:
        bias_p = torch.zeros(8, dtype=f32, device=x.device)
        bias_p[:cout].copy_(bias)
    y8 = tap_gemm(x, wp, taps_3x3(), B, H, W, bias=bias_p, out_dtype=f32)
    y = torch.empty((B, cout, H, W), dtype=f32, device=x.device)
    L.call("pddm_nhwc_slice_to_nchw", L.ptr(y8), L.ptr(y), B, cout, H * W, 8, L.stream())
    return y


def head_conv_bwd(x, w, dy_nchw):
    B, H, W, Cin = x.shape
    cout = w.shape[0]
    dyp = torch.empty((B, H, W, 32), dtype=bf16, device=x.device)
    L.call("pddm_nchw_to_nhwc_padded", L.ptr(_chk(dy_nchw, f32)), L.ptr(dyp), B, cout, H * W, 32, L.stream())
    wp1 = pack_weight(w, 1, cout_pad=32)  # [Cin, 9, 32]
    dx = tap_gemm(dyp, wp1, taps_3x3(), B, H, W)
    dw32 = tap_wgrad(x, dyp, taps_3x3(), B, H, W, Cin, 32, (32, Cin, 3, 3))
    return dx, dw32[:cout].contiguous(), colsum(dyp, 32)[:cout].contiguous()


# ------------------------------------------------------------------------------------------ diffusion math
class DeviceTables:
    """The Engine's fp32 coefficient tables resident on one device (built once; src/engine.py:121-150)."""

    def __init__(self, tables_cpu, device):
        self.T = int(tables_cpu["betas"].shape[0])
        pv = tables_cpu["posterior_variance"]
        extra = {
            "posterior_log_variance_clipped": torch.log(torch.cat([pv[1:2], pv[1:]])) if self.T > 1 else torch.log(pv),
            "log_betas": torch.log(tables_cpu["betas"]),
        }
        self.t = {k: v.to(device=device, dtype=f32).contiguous() for k, v in {**tables_cpu, **extra}.items()}
        self.device = torch.device(device)

    def struct(self):
        s = L.Tables()
        for name, _ in L.Tables._fields_:
            if name != "T":
                setattr(s, name, L.ptr(self.t[name]))
        s.T = self.T
        return s


def q_sample(x0, noise, t, tabs: DeviceTables):
    L.require_device(x0)
    _chk(x0, f32)
    _chk(noise, f32)
    out = torch.empty_like(x0)
    p = L.QSampleParams()
    p.x0, p.noise, p.x_t = L.ptr(x0), L.ptr(noise), L.ptr(out)
    if isinstance(t, int):
        p.t, p.t_const = None, t
    else:
        t = t.to(device=x0.device, dtype=torch.int64).contiguous()
        if t.numel() == 1 and x0.shape[0] != 1:
            t = t.expand(x0.shape[0]).contiguous()
        p.t = L.ptr(t)
    p.alphas_hat_sqrt, p.one_min_alphas_hat_sqrt = L.ptr(tabs.t["alphas_hat_sqrt"]), L.ptr(tabs.t["one_min_alphas_hat_sqrt"])
    p.B, p.chw = x0.shape[0], x0.numel() // x0.shape[0]
    L.call("pddm_q_sample", C.byref(p), L.stream())
    return out


def sq_err(pred, noise, gscale=None, want_grad=False, grad_v_unit=None, v_scale=0.0):
    """per-sample mean (noise - pred[:, :C])^2 ; optionally d/dpred scaled by gscale[b].  The remaining (variance)
    channels of the gradient are zero, or gscale[b]*v_scale*grad_v_unit when given (L_hybrid)."""
    L.require_device(pred)
    _chk(pred, f32)
    _chk(noise, f32)
    B, Cc = noise.shape[0], noise.shape[1]
    hw = noise.numel() // (B * Cc)
    per = torch.empty(B, dtype=f32, device=pred.device)
    grad = None
    p = L.SqErrParams()
    p.pred, p.noise, p.per_sample = L.ptr(pred), L.ptr(noise), L.ptr(per)
    if want_grad:
        full = pred.shape[1] == Cc or grad_v_unit is not None
        grad = torch.empty_like(pred) if full else torch.zeros_like(pred)
        p.grad_pred, p.gscale = L.ptr(grad), L.ptr(_chk(gscale, f32))
        if grad_v_unit is not None:
            p.grad_v_unit, p.v_scale = L.ptr(_chk(grad_v_unit, f32)), float(v_scale)
    p.B, p.C, p.c_total, p.hw = B, Cc, pred.shape[1], hw
    L.call("pddm_sq_err", C.byref(p), L.stream())
    return per, grad


SIGMA_MODES = {"beta": 0, "beta_tilde": 1, "learned": 2}


def p_sample_step(x_t, model_out, z, t_step, tabs: DeviceTables, clip, sigma_mode, out=None, t_dev=None):
    L.require_device(x_t)
    _chk(x_t, f32)
    _chk(model_out, f32)
    out = torch.empty_like(x_t) if out is None else out
    p = L.PSampleParams()
    p.x_t, p.model_out, p.z, p.x_prev = L.ptr(x_t), L.ptr(model_out), L.ptr(z), L.ptr(out)
    p.tab = tabs.struct()
    p.t_step_dev = L.ptr(t_dev)
    p.t_step = int(t_step) if t_dev is None else 1
    B, Cc = x_t.shape[0], x_t.shape[1]
    p.B, p.C, p.c_out, p.hw = B, Cc, model_out.shape[1], x_t.numel() // (B * Cc)
    p.clip, p.sigma_mode = 1 if clip else 0, SIGMA_MODES[sigma_mode]
    L.call("pddm_p_sample_step", C.byref(p), L.stream())
    return out


def vlb_terms(x0, x_t, model_out, t, tabs: DeviceTables, mode, sigma_mode="beta", want_grad_v=False):
    L.require_device(x0)
    _chk(x0, f32)
    B, Cc = x0.shape[0], x0.shape[1]
    out = torch.empty(B, dtype=f32, device=x0.device)
    gv = torch.empty_like(x0) if want_grad_v else None
    p = L.VlbParams()
    p.x0, p.out, p.grad_v = L.ptr(x0), L.ptr(out), L.ptr(gv)
    if mode != 2:
        t = t.to(device=x0.device, dtype=torch.int64).contiguous()
        p.x_t, p.model_out, p.t = L.ptr(_chk(x_t, f32)), L.ptr(_chk(model_out, f32)), L.ptr(t)
        p.c_out = model_out.shape[1]
    p.tab = tabs.struct()
    p.B, p.C, p.hw, p.mode = B, Cc, x0.numel() // (B * Cc), mode
    p.sigma_mode = SIGMA_MODES[sigma_mode] if sigma_mode in SIGMA_MODES else 0
    L.call("pddm_vlb_terms", C.byref(p), L.stream())
    return out, gv


def images_to_uint8(x, mean=None, std=None):
    """fp32 NCHW samples -> uint8 NHWC pixels: uint8(255 * clip(x * std + mean, 0, 1)) in one pass (the reference's
    unnormalize(clip=True) + image writer, src/datasets/data.py:108-128).  mean/std: per-channel sequences or None."""
    L.require_device(x)
    _chk(x, f32)
    B, Cc = x.shape[0], x.shape[1]
    out = torch.empty((B,) + tuple(x.shape[2:]) + (Cc,), dtype=torch.uint8, device=x.device)
    m = s = None
    if mean is not None:
        m = torch.tensor(list(mean), dtype=torch.float64, device=x.device)
        s = torch.tensor(list(std), dtype=torch.float64, device=x.device)
        assert m.numel() == Cc and s.numel() == Cc
    L.call("pddm_images_to_uint8", L.ptr(x), L.ptr(out), B, Cc, x.numel() // (B * Cc), L.ptr(m), L.ptr(s), L.stream())
    return out


def step_advance(t_dev, t_vec):
    L.call("pddm_step_advance", L.ptr(t_dev), L.ptr(t_vec), t_vec.shape[0], L.stream())


def adam_ema_step(param, grad, exp_avg, exp_avg_sq, ema, lr, beta1, beta2, eps, weight_decay, ema_decay, step,
                  grad_scale=1.0, step_dev=None, lr_dev=None):
    p = L.AdamParams()
    p.param, p.grad, p.exp_avg, p.exp_avg_sq, p.ema = L.ptr(param), L.ptr(grad), L.ptr(exp_avg), L.ptr(exp_avg_sq), L.ptr(ema)
    p.n = param.numel()
    p.lr, p.beta1, p.beta2, p.eps, p.weight_decay = lr, beta1, beta2, eps, weight_decay
    p.ema_decay, p.grad_scale, p.step = (ema_decay if ema_decay is not None else 0.0), grad_scale, step
    p.step_dev, p.lr_dev = L.ptr(step_dev), L.ptr(lr_dev)
    L.call("pddm_adam_ema_step", C.byref(p), L.stream())
