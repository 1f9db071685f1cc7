"""Model factory: drop-in for ``src.modules.get_model`` / ``get_unet`` (src/modules/__init__.py:7-49)."""
from .unet import UNetModel


def get_model(resolution, cfg):
    """``cfg`` keys = config/model/unet*.yaml; raises ValueError for anything but ``unet`` like the reference."""
    name = cfg.pop("name")
    if name != "unet":
        raise ValueError(f"Only 'unet' model supported.")
    return get_unet(resolution, **cfg)


def get_unet(resolution, in_channels, model_channels, num_res_blocks, attention_resolutions, dropout=0,
             channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
             num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False, learn_sigma=False):
    """Pixel ``attention_resolutions`` become downsample rates by integer division (src/modules/__init__.py:30-32).
    ``conv_resample`` / ``dims`` are accepted and ignored exactly as the reference does (:22-23 vs :36-49).
    ``learn_sigma`` (default False = reference behaviour, where it is hard-coded False at :34) doubles the output
    channels to [eps | v]."""
    attention_ds = [resolution // int(res) for res in attention_resolutions]
    return UNetModel(
        in_channels=in_channels, model_channels=model_channels,
        out_channels=(in_channels if not learn_sigma else in_channels * 2),
        num_res_blocks=num_res_blocks, attention_resolutions=tuple(attention_ds), dropout=dropout,
        channel_mult=channel_mult, num_classes=num_classes, use_checkpoint=use_checkpoint, num_heads=num_heads,
        num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm)
