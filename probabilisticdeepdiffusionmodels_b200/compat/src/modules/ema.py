from probabilisticdeepdiffusionmodels_b200.weight_average import Ema  # noqa: F401
