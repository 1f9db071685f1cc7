from probabilisticdeepdiffusionmodels_b200.engine import Engine  # noqa: F401
from probabilisticdeepdiffusionmodels_b200.schedules import (betas_for_alpha_bar, cosine_alpha_bar, get_betas,  # noqa: F401
                                                             get_linear_alphas_bar, mixed_alpha_bar)
