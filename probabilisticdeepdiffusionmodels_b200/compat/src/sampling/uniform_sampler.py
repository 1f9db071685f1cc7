from probabilisticdeepdiffusionmodels_b200.timesteps import UniformSampler  # noqa: F401
