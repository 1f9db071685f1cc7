"""The reference's model configurations (config/model/*.yaml) as plain dicts, a deterministic synthetic
initialisation for benchmarks, and an algorithmic FLOP counter that walks the module tree.

(The CPU oracle keeps its own copy of the configurations -- ``tests/test_host_cpu.py`` checks the two agree -- so that
nothing on the product or measurement path imports ``oracle/``.)"""
import torch
import torch.nn as nn

MODEL_CONFIGS = {
    # config/model/unet.yaml (CIFAR-10)
    "unet": dict(name="unet", in_channels=3, model_channels=128, num_res_blocks=3, attention_resolutions=[16, 8],
                 dropout=0, channel_mult=[1, 2, 2, 2], conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, num_heads=4, num_heads_upsample=-1, use_scale_shift_norm=False),
    # config/model/unet_small_grey.yaml (MNIST)
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


def synthetic_init_(model, seed=0, std=0.02):
    """Deterministic benchmark weights: keep torch's default initialisation of every layer, but give the reference's
    ``zero_module`` layers (all-zero at construction: the second conv of each ResBlock, attention ``proj_out``, the
    output conv) small random values, so that no branch of the network or of its backward pass is identically zero."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() > 1 and float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return model


def unet_fwd_flops_per_image(model, resolution):
    """Algorithmic forward FLOPs per image (SURVEY.md section 8(d) convention): 2*Cout*Cin*kh*kw*Hout*Wout per conv,
    2*in*out per linear, 4*heads*T^2*d per attention -- counted by walking the UNet's module tree while tracking the
    spatial resolution through the Downsample / Upsample blocks."""
    from .unet import AttentionBlock, Downsample, ResBlock, Upsample

    def conv_fl(conv, r_out):
        k = conv.kernel_size[0] * (conv.kernel_size[1] if len(conv.kernel_size) > 1 else 1)
        return 2 * conv.out_channels * conv.in_channels * k * r_out * r_out

    def lin_fl(m):
        return 2 * m.in_features * m.out_features

    def block_fl(mod, r):
        """-> (flops, resolution after the module)"""
        if isinstance(mod, ResBlock):
            f = conv_fl(mod.in_layers[2], r) + conv_fl(mod.out_layers[3], r) + lin_fl(mod.emb_layers[1])
            if not isinstance(mod.skip_connection, nn.Identity):
                f += conv_fl(mod.skip_connection, r)
            return f, r
        if isinstance(mod, AttentionBlock):
            c, t = mod.channels, r * r
            return 2 * c * 3 * c * t + 2 * c * c * t + 4 * mod.num_heads * t * t * (c // mod.num_heads), r
        if isinstance(mod, Downsample):
            r2 = (r + 1) // 2
            return conv_fl(mod.op, r2), r2
        if isinstance(mod, Upsample):
            return conv_fl(mod.conv, 2 * r), 2 * r
        if isinstance(mod, (nn.Conv2d, nn.Conv1d)):
            return conv_fl(mod, r), r
        return 0, r

    total = sum(lin_fl(m) for m in model.time_embed if isinstance(m, nn.Linear))
    r = resolution
    for seq in list(model.input_blocks) + [model.middle_block] + list(model.output_blocks):
        for mod in seq:
            f, r = block_fl(mod, r)
            total += f
    total += conv_fl(model.out[2], r)
    return total


def conv3x3_census(model, resolution):
    """The model's forward 3x3 convolutions (stride 1 and 2, stem / head excluded: they run thin-GEMM paths) as
    ``[(H_out, Cin, Cout, count)]`` -- the workload of the dominant tap-GEMM kernel."""
    from collections import OrderedDict

    from .unet import Downsample, ResBlock, Upsample
    census = OrderedDict()

    def add(r, conv):
        key = (r, conv.in_channels, conv.out_channels)
        census[key] = census.get(key, 0) + 1

    r = resolution
    for seq in list(model.input_blocks)[1:] + [model.middle_block] + list(model.output_blocks):
        for mod in seq:
            if isinstance(mod, ResBlock):
                add(r, mod.in_layers[2])
                add(r, mod.out_layers[3])
            elif isinstance(mod, Downsample):
                r = (r + 1) // 2
                add(r, mod.op)
            elif isinstance(mod, Upsample):
                r = 2 * r
                add(r, mod.conv)
    return [(k[0], k[1], k[2], n) for k, n in census.items()]
