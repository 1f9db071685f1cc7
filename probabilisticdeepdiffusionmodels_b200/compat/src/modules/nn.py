from probabilisticdeepdiffusionmodels_b200.nn import (CheckpointFunction, GroupNorm32, SiLU, avg_pool_nd,  # noqa: F401
                                                      checkpoint, conv_nd, linear, mean_flat, normalization,
                                                      scale_module, timestep_embedding, update_ema, zero_module)
