from probabilisticdeepdiffusionmodels_b200.unet import (AttentionBlock, Downsample, QKVAttention, ResBlock,  # noqa: F401
                                                        TimestepBlock, TimestepEmbedSequential, UNetModel, Upsample)
