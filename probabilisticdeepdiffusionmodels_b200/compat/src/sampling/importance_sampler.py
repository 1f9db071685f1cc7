from probabilisticdeepdiffusionmodels_b200.timesteps import ImportanceSampler  # noqa: F401
