from probabilisticdeepdiffusionmodels_b200.mathutils import (approx_standard_normal_cdf,  # noqa: F401
                                                             discretized_gaussian_log_likelihood,
                                                             get_generator_if_specified, mean_flat, normal_kl)
