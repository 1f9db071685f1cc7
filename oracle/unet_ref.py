"""Oracle (test infrastructure): functional fp32 restatement of the reference UNet.

``unet_forward(params, arch, x, t)`` evaluates the network of
src/modules/unet.py:282-495 from a plain ``{name: tensor}`` dict whose keys and
shapes are exactly the reference ``UNetModel.state_dict()`` (so a reference
checkpoint, the product model and this oracle are interchangeable).  It is
written against ``torch.nn.functional`` only -- no Module classes -- and is
differentiable, so it also serves as the gradient oracle.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .diffusion_ref import timestep_embedding


def arch_from_config(resolution, in_channels, model_channels, num_res_blocks, attention_resolutions,
                     dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                     use_checkpoint=False, num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False,
                     learn_sigma=False, name="unet"):
    """Mirror of get_unet (src/modules/__init__.py:14-49): pixel resolutions -> downsample rates."""
    if name != "unet":
        raise ValueError("Only 'unet' model supported.")
    return dict(
        in_channels=in_channels, model_channels=model_channels,
        out_channels=in_channels * (2 if learn_sigma else 1),
        num_res_blocks=num_res_blocks,
        attention_ds=tuple(resolution // int(r) for r in attention_resolutions),
        channel_mult=tuple(channel_mult), num_heads=num_heads,
        num_heads_upsample=num_heads if num_heads_upsample == -1 else num_heads_upsample,
        use_scale_shift_norm=use_scale_shift_norm, dropout=dropout,
    )


def block_plan(arch):
    """Enumerate the blocks in construction order (src/modules/unet.py:347-441).

    Returns (input_blocks, middle, output_blocks); each block is a list of layer descriptors
    ``(kind, index_in_block, cin, cout, heads)`` with kind in {stem, res, attn, down, up}.
    """
    mc, mult, nrb = arch["model_channels"], arch["channel_mult"], arch["num_res_blocks"]
    att = arch["attention_ds"]
    inp = [[("stem", 0, arch["in_channels"], mc, 0)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", 0, ch, m * mc, 0)]
            ch = m * mc
            if ds in att:
                layers.append(("attn", 1, ch, ch, arch["num_heads"]))
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", 0, ch, ch, 0)])
            chans.append(ch)
            ds *= 2
    mid = [("res", 0, ch, ch, 0), ("attn", 1, ch, ch, arch["num_heads"]), ("res", 2, ch, ch, 0)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            layers = [("res", 0, ch + chans.pop(), mc * m, 0)]
            ch = mc * m
            if ds in att:
                layers.append(("attn", len(layers), ch, ch, arch["num_heads_upsample"]))
            if level and i == nrb:
                layers.append(("up", len(layers), ch, ch, 0))
                ds //= 2
            out.append(layers)
    return inp, mid, out


def param_shapes(arch):
    """``{state_dict key: shape}`` in reference registration order."""
    mc = arch["model_channels"]
    ted = 4 * mc
    shapes = {}

    def lin(p, i, o):
        shapes[p + ".weight"] = (o, i)
        shapes[p + ".bias"] = (o,)

    def conv(p, i, o, k):
        shapes[p + ".weight"] = (o, i, k, k)
        shapes[p + ".bias"] = (o,)

    def gn(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)

    def res(p, cin, cout):
        gn(p + ".in_layers.0", cin)
        conv(p + ".in_layers.2", cin, cout, 3)
        lin(p + ".emb_layers.1", ted, cout * (2 if arch["use_scale_shift_norm"] else 1))
        gn(p + ".out_layers.0", cout)
        conv(p + ".out_layers.3", cout, cout, 3)
        if cin != cout:
            conv(p + ".skip_connection", cin, cout, 1)

    def attn(p, c):
        gn(p + ".norm", c)
        shapes[p + ".qkv.weight"] = (3 * c, c, 1)
        shapes[p + ".qkv.bias"] = (3 * c,)
        shapes[p + ".proj_out.weight"] = (c, c, 1)
        shapes[p + ".proj_out.bias"] = (c,)

    def layer(p, kind, cin, cout):
        if kind == "stem":
            conv(p, cin, cout, 3)
        elif kind == "res":
            res(p, cin, cout)
        elif kind == "attn":
            attn(p, cin)
        elif kind == "down":
            conv(p + ".op", cin, cout, 3)
        elif kind == "up":
            conv(p + ".conv", cin, cout, 3)

    inp, mid, out = block_plan(arch)
    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)
    for bi, layers in enumerate(inp):
        for kind, li, cin, cout, _ in layers:
            layer(f"input_blocks.{bi}.{li}", kind, cin, cout)
    for kind, li, cin, cout, _ in mid:
        layer(f"middle_block.{li}", kind, cin, cout)
    for bi, layers in enumerate(out):
        for kind, li, cin, cout, _ in layers:
            layer(f"output_blocks.{bi}.{li}", kind, cin, cout)
    gn("out.0", mc)
    conv("out.2", mc, arch["out_channels"], 3)
    return shapes


def make_params(arch, seed=0, scale=1.0):
    """Deterministic synthetic parameters (numpy RandomState; stable across machines and versions).

    The reference zero-initialises every ResBlock out-conv, attention proj_out and the head conv
    (src/modules/nn.py:69-75), which makes a fresh model output identically 0 --
    useless for parity.  Instead every weight is N(0, s^2) with a fan-in-scaled s, biases are small and
    GroupNorm affine parameters are perturbed around (1, 0).
    """
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in param_shapes(arch).items():
        if name.endswith(".weight") and len(shape) == 1:  # GroupNorm gamma
            a = 1.0 + 0.1 * rs.standard_normal(shape)
        elif name.endswith(".bias"):
            a = 0.05 * rs.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = scale * rs.standard_normal(shape) / math.sqrt(fan_in)
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


def _gn(x, w, b):
    """GroupNorm32: 32 groups, eps 1e-5, fp32 (src/modules/nn.py:18-20,94-101)."""
    return F.group_norm(x.float(), 32, w, b, eps=1e-5).type(x.dtype)


def _silu(x):
    """src/modules/nn.py:13-15"""
    return x * torch.sigmoid(x)


def _res(P, p, x, emb, scale_shift, dropout_p, training):
    """ResBlock._forward  (src/modules/unet.py:188-201)."""
    h = F.conv2d(_silu(_gn(x, P[p + ".in_layers.0.weight"], P[p + ".in_layers.0.bias"])),
                 P[p + ".in_layers.2.weight"], P[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(_silu(emb), P[p + ".emb_layers.1.weight"], P[p + ".emb_layers.1.bias"])[:, :, None, None]
    if scale_shift:
        sc, sh = torch.chunk(e, 2, dim=1)
        h = _gn(h, P[p + ".out_layers.0.weight"], P[p + ".out_layers.0.bias"]) * (1 + sc) + sh
    else:
        h = _gn(h + e, P[p + ".out_layers.0.weight"], P[p + ".out_layers.0.bias"])
    h = F.dropout(_silu(h), dropout_p, training)
    h = F.conv2d(h, P[p + ".out_layers.3.weight"], P[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in P:
        x = F.conv2d(x, P[p + ".skip_connection.weight"], P[p + ".skip_connection.bias"])
    return x + h


def _attn(P, p, x, heads):
    """AttentionBlock._forward + QKVAttention  (src/modules/unet.py:226-256)."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(xf, P[p + ".norm.weight"], P[p + ".norm.bias"]), P[p + ".qkv.weight"], P[p + ".qkv.bias"])
    qkv = qkv.reshape(b * heads, -1, qkv.shape[2])
    ch = qkv.shape[1] // 3
    q, k, v = torch.split(qkv, ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1).type(q.dtype)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, hh * ww)
    a = F.conv1d(a, P[p + ".proj_out.weight"], P[p + ".proj_out.bias"])
    return (xf + a).reshape(b, c, hh, ww)


def _layer(P, p, kind, heads, h, emb, arch, training):
    if kind == "stem":
        return F.conv2d(h, P[p + ".weight"], P[p + ".bias"], padding=1)
    if kind == "res":
        return _res(P, p, h, emb, arch["use_scale_shift_norm"], arch["dropout"], training)
    if kind == "attn":
        return _attn(P, p, h, heads)
    if kind == "down":  # src/modules/unet.py:85-108, strided conv
        return F.conv2d(h, P[p + ".op.weight"], P[p + ".op.bias"], stride=2, padding=1)
    if kind == "up":  # src/modules/unet.py:54-82, nearest x2 then conv
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        return F.conv2d(h, P[p + ".conv.weight"], P[p + ".conv.bias"], padding=1)
    raise ValueError(kind)


def unet_forward(P, arch, x, timesteps, training=False):
    """UNetModel.forward  (src/modules/unet.py:466-495)."""
    inp, mid, out = block_plan(arch)
    emb = timestep_embedding(timesteps, arch["model_channels"])
    emb = F.linear(emb, P["time_embed.0.weight"], P["time_embed.0.bias"])
    emb = F.linear(_silu(emb), P["time_embed.2.weight"], P["time_embed.2.bias"])
    hs = []
    h = x
    for bi, layers in enumerate(inp):
        for kind, li, _, _, heads in layers:
            h = _layer(P, f"input_blocks.{bi}.{li}", kind, heads, h, emb, arch, training)
        hs.append(h)
    for kind, li, _, _, heads in mid:
        h = _layer(P, f"middle_block.{li}", kind, heads, h, emb, arch, training)
    for bi, layers in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        for kind, li, _, _, heads in layers:
            h = _layer(P, f"output_blocks.{bi}.{li}", kind, heads, h, emb, arch, training)
    h = _silu(_gn(h, P["out.0.weight"], P["out.0.bias"]))
    return F.conv2d(h, P["out.2.weight"], P["out.2.bias"], padding=1)


def fwd_flops_per_image(arch, resolution):
    """Algorithmic forward FLOPs per image (SURVEY.md section 8(d) convention):
    2*Cout*Cin*kh*kw*Hout*Wout per conv, 2*in*out per linear, 4*heads*T^2*d per attention."""
    mc = arch["model_channels"]
    ted = 4 * mc
    total = 2 * mc * ted + 2 * ted * ted
    inp, mid, out = block_plan(arch)
    res = resolution

    def res_fl(cin, cout, r):
        f = 2 * cout * cin * 9 * r * r + 2 * cout * cout * 9 * r * r
        f += 2 * ted * cout * (2 if arch["use_scale_shift_norm"] else 1)
        if cin != cout:
            f += 2 * cin * cout * r * r
        return f

    def attn_fl(c, heads, r):
        T = r * r
        return 2 * c * 3 * c * T + 2 * c * c * T + 4 * heads * T * T * (c // heads)

    def walk(layers, r):
        f = 0
        for kind, _, cin, cout, heads in layers:
            if kind == "stem":
                f += 2 * cout * cin * 9 * r * r
            elif kind == "res":
                f += res_fl(cin, cout, r)
            elif kind == "attn":
                f += attn_fl(cin, heads, r)
            elif kind == "down":
                r = (r + 1) // 2
                f += 2 * cout * cin * 9 * r * r
            elif kind == "up":
                r = r * 2
                f += 2 * cout * cin * 9 * r * r
        return f, r

    for layers in inp:
        f, res = walk(layers, res)
        total += f
    f, res = walk(mid, res)
    total += f
    for layers in out:
        f, res = walk(layers, res)
        total += f
    total += 2 * arch["out_channels"] * mc * 9 * res * res
    return total


MODEL_CONFIGS = {
    # config/model/unet.yaml:1-14 (CIFAR)
    "unet": dict(name="unet", in_channels=3, model_channels=128, num_res_blocks=3, attention_resolutions=[16, 8],
                 dropout=0, channel_mult=[1, 2, 2, 2], conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, num_heads=4, num_heads_upsample=-1, use_scale_shift_norm=False),
    # config/model/unet_small_grey.yaml
    "unet_small_grey": dict(name="unet", in_channels=1, model_channels=32, num_res_blocks=1, attention_resolutions=[],
                            dropout=0, channel_mult=[1, 2, 2], conv_resample=True, dims=2, num_classes=None,
                            use_checkpoint=False, num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False),
    # config/model/unet_small.yaml
    "unet_small": dict(name="unet", in_channels=3, model_channels=32, num_res_blocks=1, attention_resolutions=[],
                       dropout=0, channel_mult=[1, 2, 2], conv_resample=True, dims=2, num_classes=None,
                       use_checkpoint=False, num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False),
    # config/model/unet_grey.yaml
    "unet_grey": dict(name="unet", in_channels=1, model_channels=128, num_res_blocks=2, attention_resolutions=[16, 8],
                      dropout=0, channel_mult=[1, 2, 2, 2], conv_resample=True, dims=2, num_classes=None,
                      use_checkpoint=False, num_heads=4, num_heads_upsample=-1, use_scale_shift_norm=False),
    # config/model/unet_celeba.yaml
    "unet_celeba": dict(name="unet", in_channels=3, model_channels=128, num_res_blocks=3,
                        attention_resolutions=[16, 8], dropout=0, channel_mult=[1, 2, 3, 4], conv_resample=True,
                        dims=2, num_classes=None, use_checkpoint=False, num_heads=4, num_heads_upsample=-1,
                        use_scale_shift_norm=False),
    # config/model/unet_celebahq.yaml
    "unet_celebahq": dict(name="unet", in_channels=3, model_channels=128, num_res_blocks=3,
                          attention_resolutions=[16, 8], dropout=0, channel_mult=[1, 1, 2, 2, 4, 4],
                          conv_resample=True, dims=2, num_classes=None, use_checkpoint=False, num_heads=4,
                          num_heads_upsample=-1, use_scale_shift_norm=False),
}
