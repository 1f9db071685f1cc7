from probabilisticdeepdiffusionmodels_b200.timesteps import StepwiseLog  # noqa: F401
