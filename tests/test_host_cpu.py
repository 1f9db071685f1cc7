"""Host-side logic on CPU: schedules against the reference fixtures, samplers / per-t log, EMA, Engine construction
and API surface, state_dict compatibility with the reference module tree, the compat import paths."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params, param_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["linear", "cosine", "mixed"])
@pytest.mark.parametrize("steps", [1000, 50])
def test_schedule_tables_bit_exact_vs_reference(golden, mode, steps):
    from probabilisticdeepdiffusionmodels_b200.schedules import TABLE_NAMES, get_betas, make_tables
    g = golden["schedules"]
    tabs = make_tables(get_betas(None, None, steps, mode))
    for name in TABLE_NAMES:
        assert tabs[name].dtype == torch.float32
        np.testing.assert_array_equal(tabs[name].numpy(), g[f"{mode}_{steps}_{name}"], err_msg=name)
    with pytest.raises(ValueError):
        get_betas(mode="bogus")
    np.testing.assert_array_equal(get_betas(1e-3, 5e-2, 100, "linear").numpy(), g["linear_custom_100"])


def test_mathutils_match_reference_kats(golden):
    from probabilisticdeepdiffusionmodels_b200 import mathutils as M
    g = golden["kats"]
    T = lambda k: torch.from_numpy(g[k])  # noqa: E731
    np.testing.assert_array_equal(M.normal_kl(T("kl_m1"), T("kl_lv1"), T("kl_m2"), T("kl_lv2")).numpy(), g["kl_out"])
    np.testing.assert_array_equal(M.normal_kl(T("kl_m1"), T("kl_lv1"), 0.0, 0.0).numpy(), g["kl_scalar_out"])
    np.testing.assert_array_equal(
        M.discretized_gaussian_log_likelihood(T("dll_x"), T("dll_means"), T("dll_ls")).numpy(), g["dll_out"])
    np.testing.assert_array_equal(M.approx_standard_normal_cdf(T("cdf_in")).numpy(), g["cdf_out"])
    np.testing.assert_array_equal(M.mean_flat(T("kl_m1")).numpy(), g["mean_flat_out"])
    assert M.get_generator_if_specified(None) is None
    assert M.get_generator_if_specified(3).initial_seed() == 3


@pytest.mark.parametrize("name,res", [("unet", 32), ("unet_small_grey", 28), ("unet_celeba", 64), ("unet_celebahq", 64),
                                      ("unet_grey", 32), ("unet_small", 32)])
def test_state_dict_matches_reference_module_tree(name, res):
    from probabilisticdeepdiffusionmodels_b200 import get_model
    cfg = MODEL_CONFIGS[name]
    m = get_model(res, dict(cfg))
    arch = arch_from_config(res, **{k: v for k, v in cfg.items() if k != "name"})
    shapes = param_shapes(arch)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) and sd[k].dtype == torch.float32 for k in shapes)
    assert m.in_channels == cfg["in_channels"]
    m.load_state_dict(make_params(arch, seed=0))
    m2 = copy.deepcopy(m)  # EMA relies on deepcopy
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    # the reference zero-initialises out convs / proj_out / head (src/modules/nn.py:69-75)
    fresh = get_model(res, dict(cfg))
    assert float(fresh.out[2].weight.abs().max()) == 0.0
    assert float(fresh.middle_block[1].proj_out.weight.abs().max()) == 0.0
    assert float(fresh.middle_block[0].out_layers[3].weight.abs().max()) == 0.0


def test_get_model_contract():
    from probabilisticdeepdiffusionmodels_b200 import get_model, get_unet
    from probabilisticdeepdiffusionmodels_b200.unet import AttentionBlock
    with pytest.raises(ValueError):
        get_model(32, {"name": "dense"})
    m = get_unet(32, 3, 32, 1, [16], channel_mult=(1, 2), learn_sigma=True)
    assert m.out_channels == 6
    m = get_unet(168, 3, 32, 1, [16, 8], channel_mult=(1, 2))  # 168//16 = 10, 168//8 = 21: no power-of-two ds
    assert sum(isinstance(x, AttentionBlock) for x in m.modules()) == 1  # only the always-present middle attention


def test_samplers_and_stepwise_log():
    from probabilisticdeepdiffusionmodels_b200 import ImportanceSampler, StepwiseLog, UniformSampler
    torch.manual_seed(0)
    t, w = UniformSampler(50)(1000, "cpu")
    assert w is None and t.dtype == torch.int64 and int(t.min()) >= 1 and int(t.max()) <= 50
    log = StepwiseLog(5, 10)
    imp = ImportanceSampler(5, log, min_counts=2)
    t, w = imp(8, "cpu")
    assert w is None and not imp.is_ready()
    for step in range(1, 6):
        log.update_multiple([step, step, step], [0.1 * step, 0.2 * step, float("nan")])
    assert log.n_per_step.tolist() == [2, 2, 2, 2, 2]  # non-finite entries are dropped
    np.testing.assert_allclose(log.avg_per_step, [0.15 * s for s in range(1, 6)])
    np.testing.assert_allclose(log.avg_sq_per_step[0], np.sqrt((0.01 + 0.04) / 2))
    assert imp.is_ready()
    np.random.seed(0)
    t, w = imp(64, "cpu")
    assert w.dtype == torch.float64 and int(t.min()) >= 1 and int(t.max()) <= 5
    p = log.avg_sq_per_step + 1e-6
    p /= p.sum()
    np.testing.assert_allclose(w.numpy(), 1 / (p[t.numpy() - 1] * 64))
    assert log.get_avg_in_range(1, 3) == pytest.approx(np.mean([0.1, 0.2, 0.2, 0.4]))
    # the reference truncates histories whenever T > max_keep (it tests len(metric_per_t), stepwise_log.py:19-20)
    big = StepwiseLog(12, 3)
    for v in range(6):
        big.update(4, float(v))
    assert big[4] == [3.0, 4.0, 5.0]
    small = StepwiseLog(2, 3)
    for v in range(6):
        small.update(1, float(v))
    assert len(small[1]) == 6


def test_ema_matches_reference_formula():
    from probabilisticdeepdiffusionmodels_b200 import Ema
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
    ema = Ema(net, decay=0.9)
    assert not any(p.requires_grad for p in ema.module.parameters()) and not ema.module.training
    before = {k: v.clone() for k, v in ema.module.state_dict().items()}
    with torch.no_grad():
        for p in net.parameters():
            p.add_(1.0)
    ema.update(net)
    for k, v in ema.module.state_dict().items():
        if v.is_floating_point():
            want = 0.9 * before[k] + (1.0 - 0.9) * net.state_dict()[k]
            assert torch.allclose(v, want, atol=1e-6), k
    ema.set(net)
    assert all(torch.equal(a, b) for a, b in zip(ema.module.state_dict().values(), net.state_dict().values()))


def test_engine_surface_on_cpu():
    from probabilisticdeepdiffusionmodels_b200 import Engine
    cfg = MODEL_CONFIGS["unet_small_grey"]
    eng = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=1000, mode="cosine", resolution=28, ema=0.999,
                 sampling="importance", scheduler_name="CosineAnnealingWarmRestarts", scheduler_kwargs={"T_0": 10})
    for name in ("betas", "alphas_hat_sqrt", "posterior_variance", "denoising_coef", "sqrt_recip_alphas_cumprod"):
        assert getattr(eng, name).dtype == torch.float32 and getattr(eng, name).shape == (1000,)
    for meth in ("training_step", "validation_step", "test_step", "configure_optimizers", "optimizer_step", "get_q_t",
                 "get_loss", "denoising_step", "sample_from_step", "calculate_likelihood", "sample_and_return_steps",
                 "generate_images", "generate_images_grid", "diffuse_and_reconstruct", "diffuse_and_reconstruct_grid",
                 "get_noised_representation", "q_posterior", "xstart_from_epsilon", "model_mean_from_epsilon",
                 "get_sigma", "compute_grad_norm", "ema_on", "on_epoch_end"):
        assert callable(getattr(eng, meth)), meth
    opt = eng.configure_optimizers()
    assert isinstance(opt["optimizer"], torch.optim.Adam) and opt["lr_scheduler"] is not None
    keys = list(eng.state_dict())
    assert any(k.startswith("ema.module.") for k in keys) and any(k.startswith("model.") for k in keys)
    with eng.ema_on():
        assert eng.model is eng.ema.module
    assert eng.model is not eng.ema.module
    assert eng.get_sigma(0).item() == pytest.approx(float(torch.sqrt(eng.betas[0])))
    x = torch.randn(2, 1, 28, 28)
    t = torch.tensor([1, 1000])
    mean, var = eng.q_posterior(t, x, x)  # API-parity helpers are device agnostic ...
    assert mean.shape == x.shape and var.shape == (2, 1, 1, 1)
    with pytest.raises(RuntimeError):  # ... the kernels are not: no CPU fallback
        eng.get_q_t(x, x, t)
    for bad in (dict(sigma_mode="nope"), dict(sampling="nope"), dict(mode="nope")):
        with pytest.raises(ValueError):
            Engine(dict(cfg), {"lr": 1e-3}, **bad)


def test_compat_import_paths():
    compat = os.path.join(ROOT, "probabilisticdeepdiffusionmodels_b200", "compat")
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, compat)
    try:
        from src.engine import Engine, get_betas  # noqa: F401
        from src.modules import get_model  # noqa: F401
        from src.modules.nn import conv_nd, normalization, timestep_embedding, zero_module  # noqa: F401
        from src.modules.unet import UNetModel  # noqa: F401
        from src.sampling.importance_sampler import ImportanceSampler  # noqa: F401
        from src.sampling.uniform_sampler import UniformSampler  # noqa: F401
        from src.utils import discretized_gaussian_log_likelihood, normal_kl  # noqa: F401
        import probabilisticdeepdiffusionmodels_b200 as P
        assert Engine is P.Engine and UNetModel is P.UNetModel
    finally:
        sys.path.remove(compat)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_nn_factories_keep_reference_contract():
    from probabilisticdeepdiffusionmodels_b200 import nn as N
    c = N.conv_nd(2, 8, 16, 3, padding=1)
    assert isinstance(c, torch.nn.Conv2d) and tuple(c.weight.shape) == (16, 8, 3, 3)
    c1 = N.conv_nd(1, 8, 24, 1)
    assert isinstance(c1, torch.nn.Conv1d) and tuple(c1.weight.shape) == (24, 8, 1)
    with pytest.raises(ValueError):
        N.conv_nd(4, 1, 1, 1)
    g = N.normalization(64)
    assert isinstance(g, torch.nn.GroupNorm) and g.num_groups == 32 and g.eps == 1e-5
    z = N.zero_module(N.linear(4, 4))
    assert float(z.weight.abs().sum()) == 0 and float(z.bias.abs().sum()) == 0
    assert isinstance(N.avg_pool_nd(2, 2), torch.nn.AvgPool2d)
    assert N.checkpoint(lambda a: a * 2, (torch.ones(2),), (), False).tolist() == [2.0, 2.0]
    x = torch.ones(2, requires_grad=True)
    y = N.checkpoint(lambda a: a * 3, (x,), (), True)
    y.sum().backward()
    assert x.grad.tolist() == [3.0, 3.0]


def test_package_configs_and_flop_counter_agree_with_the_oracle():
    """configs.MODEL_CONFIGS mirrors the oracle's copy of config/model/*.yaml, and the module-walking FLOP counter
    bench.py uses (nothing on the measured path imports oracle/) equals the oracle's block-plan formula."""
    from oracle.unet_ref import MODEL_CONFIGS as REF, arch_from_config, fwd_flops_per_image
    from probabilisticdeepdiffusionmodels_b200.configs import MODEL_CONFIGS, synthetic_init_, unet_fwd_flops_per_image
    from probabilisticdeepdiffusionmodels_b200.modules import get_unet
    assert MODEL_CONFIGS == REF
    for name, res, ls in [("unet", 32, True), ("unet", 32, False), ("unet_small_grey", 28, False),
                          ("unet_celeba", 64, False), ("unet_grey", 32, False)]:
        kw = {k: v for k, v in MODEL_CONFIGS[name].items() if k != "name"}
        m = get_unet(res, **kw, learn_sigma=ls)
        arch = arch_from_config(res, **kw, learn_sigma=ls)
        assert unet_fwd_flops_per_image(m, res) == fwd_flops_per_image(arch, res), name
    m = synthetic_init_(get_unet(28, **{k: v for k, v in MODEL_CONFIGS["unet_small_grey"].items() if k != "name"}), 3)
    assert all(float(p.abs().sum()) > 0 for p in m.parameters() if p.dim() > 1)


@pytest.mark.parametrize("max_keep", [None, 3, 50])
def test_device_stepwise_log_matches_host_log(max_keep):
    """DeviceStepwiseLog (batched tensor updates, no host loop) reproduces StepwiseLog's running mean / RMS / count,
    including duplicates of a timestep inside one batch, the max_keep truncation rule and skipped non-finite values."""
    import numpy as np
    from probabilisticdeepdiffusionmodels_b200.timesteps import DeviceStepwiseLog, StepwiseLog
    T = 12
    host, dev = StepwiseLog(T, max_keep), DeviceStepwiseLog(T, max_keep)
    rs = np.random.RandomState(3)
    for it in range(25):
        B = int(rs.randint(1, 20))
        ts = rs.randint(1, T + 1, size=B)
        if it % 5 == 0:
            ts[:] = ts[0]  # many duplicates of one timestep (more than max_keep=3)
        ms = rs.rand(B) * 3
        if it % 7 == 0:
            ms[0] = np.inf
        host.update_multiple(ts.tolist(), ms.tolist())
        dev.update_multiple(torch.from_numpy(ts), torch.from_numpy(ms))
        np.testing.assert_allclose(dev.n_per_step.numpy(), host.n_per_step)
        np.testing.assert_allclose(dev.avg_per_step.numpy(), host.avg_per_step, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dev.avg_sq_per_step.numpy(), host.avg_sq_per_step, rtol=1e-12, atol=1e-12)


def test_device_importance_sampler_warmup_and_weights():
    import numpy as np
    from probabilisticdeepdiffusionmodels_b200.timesteps import (DeviceImportanceSampler, DeviceStepwiseLog,
                                                                 ImportanceSampler, StepwiseLog)
    T = 6
    hlog, dlog = StepwiseLog(T), DeviceStepwiseLog(T)
    hs, ds = ImportanceSampler(T, hlog, min_counts=2), DeviceImportanceSampler(T, dlog, min_counts=2)
    g = torch.Generator().manual_seed(0)
    t, w, ready = ds(5, generator=g)
    assert not bool(ready) and not hs.is_ready() and t.min() >= 1 and t.max() <= T
    for rep in range(2):
        ts = np.arange(1, T + 1)
        ms = (np.arange(T) + 1.0) * (rep + 1)
        hlog.update_multiple(ts.tolist(), ms.tolist())
        dlog.update_multiple(torch.from_numpy(ts), torch.from_numpy(ms))
    assert hs.is_ready()
    t, w, ready = ds(4000, generator=g)
    assert bool(ready)
    p = hlog.avg_sq_per_step + 1e-6
    p = p / p.sum()
    np.testing.assert_allclose(ds.probabilities().numpy(), p, rtol=1e-12)
    np.testing.assert_allclose(w.numpy(), 1 / (p[t.numpy() - 1] * 4000), rtol=1e-6)  # the reference's weights
    freq = np.bincount(t.numpy() - 1, minlength=T) / 4000
    assert np.abs(freq - p).max() < 0.03  # draws follow p
