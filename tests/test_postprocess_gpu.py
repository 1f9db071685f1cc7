"""Sample post-processing on the device and the sharded NLL evaluation."""
import numpy as np
import pytest
import torch

from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params

pytestmark = pytest.mark.gpu
CFG = MODEL_CONFIGS["unet_small_grey"]


def _unnormalize(x, normalize=None, clip=False, channel_dim=0):
    """the reference's host function restated (src/datasets/data.py:108-128)"""
    if normalize is not None:
        mean, std = normalize
        shape = [1] * len(x.shape)
        shape[channel_dim] = x.shape[channel_dim]
        x = x * np.array(std).reshape(shape) + np.array(mean).reshape(shape)
    return np.clip(x, 0, 1) if clip else x


@pytest.mark.parametrize("C,normalize", [(3, None), (3, ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))), (1, ((0.5,), (0.5,)))])
def test_images_to_uint8_is_the_reference_postprocessing(C, normalize):
    from probabilisticdeepdiffusionmodels_b200 import functional as F
    g = torch.Generator().manual_seed(3)
    x = torch.randn((5, C, 12, 10), generator=g) * 0.8 + 0.4
    x[0, 0, 0, :4] = torch.tensor([0.0, 1.0, -0.0, 1.0000001])
    want = np.stack([(255 * _unnormalize(img, normalize, clip=True, channel_dim=0)).astype(np.uint8).transpose(1, 2, 0)
                     for img in x.numpy()])
    mean, std = normalize if normalize is not None else (None, None)
    got = F.images_to_uint8(x.cuda(), mean, std).cpu().numpy()
    assert got.dtype == np.uint8 and got.shape == (5, 12, 10, C)
    np.testing.assert_array_equal(got, want)  # integer output: bit exact


def _engine(steps=20, mode="linear"):
    from probabilisticdeepdiffusionmodels_b200 import Engine
    arch = arch_from_config(28, **{k: v for k, v in CFG.items() if k != "name"})
    eng = Engine(dict(CFG), {"lr": 1e-3}, diffusion_steps=steps, mode=mode, resolution=28, clip_while_generating=True)
    eng.model.load_state_dict(make_params(arch, seed=21))
    return eng.to("cuda")


def test_generate_images_uint8_matches_generate_images():
    eng = _engine(steps=12)
    a = eng.generate_images(n=3, minibatch=2, seed=5)
    b = eng.generate_images_uint8(n=3, minibatch=2, seed=5, normalize="mnist")
    want = np.stack([(255 * _unnormalize(img, ((0.5,), (0.5,)), clip=True)).astype(np.uint8).transpose(1, 2, 0) for img in a])
    assert b.shape == (4, 28, 28, 1) and b.dtype == np.uint8
    np.testing.assert_array_equal(b, want)


def test_sharded_nll_equals_unsharded_terms():
    """shard=(r, W): the ranks' timestep subsets partition [2, T]; with the per-timestep noise pinned, the sum of the
    ranks' partial results (what the all-reduce computes) is the unsharded result."""
    from probabilisticdeepdiffusionmodels_b200 import parallel
    eng = _engine(steps=16, mode="cosine")
    eng.eval()
    x = (torch.rand(3, 1, 28, 28, generator=torch.Generator().manual_seed(2)) * 2 - 1).cuda()
    real = torch.randn_like

    def run(steps):
        # noise keyed by nothing but call order is not comparable across shardings: key it by the timestep instead
        terms = {}
        for t in steps:
            torch.manual_seed(1000 + t)
            L, M = eng._calculate_L_intermediate(x, 1, steps=[t])
            terms[t] = (L[0], M[0])
        return terms

    with torch.no_grad():
        full = run(range(2, 17))
        W = 3
        parts = [run(parallel.shard_timesteps(16, r, W)) for r in range(W)]
    assert sorted(t for p in parts for t in p) == list(range(2, 17))
    tot = sum(L for p in parts for (L, _) in p.values())
    want = sum(L for (L, _) in full.values())
    assert torch.allclose(tot, want, rtol=1e-5, atol=1e-6)
    # single-process shard=(0, 1) goes through the same reduction code path and reproduces the plain call
    with torch.no_grad():
        torch.manual_seed(7)
        a = eng.calculate_likelihood(x)
        torch.manual_seed(7)
        b = eng.calculate_likelihood(x, shard=(0, 1))
    # (the plain call draws L_0's noise first, the sharded one last: compare the noise-independent term exactly and
    #  the noise-dependent ones statistically)
    assert torch.allclose(a["L_T"], b["L_T"])
    assert abs(a["nll"].item() - b["nll"].item()) < 0.05 * abs(a["nll"].item())
