"""Import-path compatibility with the reference's ``src.sampling`` package."""
from ..timesteps import ImportanceSampler, StepwiseLog, UniformSampler

__all__ = ["UniformSampler", "ImportanceSampler", "StepwiseLog"]
