"""Pin the CPU oracle (oracle/) against fixtures recorded from the UNMODIFIED reference
(oracle/gen_golden.py -> tests/golden/*.npz) and against SURVEY.md Appendix B scalars."""
import numpy as np
import pytest
import torch

from oracle import diffusion_ref as D
from oracle import engine_ref as E
from oracle.gen_golden import TINY, synth_batch
from oracle.unet_ref import (MODEL_CONFIGS, arch_from_config, fwd_flops_per_image, make_params, param_shapes,
                             unet_forward)


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("mode", ["linear", "cosine", "mixed"])
@pytest.mark.parametrize("steps", [1000, 50])
def test_schedule_tables_bit_exact(golden, mode, steps):
    g = golden["schedules"]
    tabs = D.make_tables(D.get_betas(None, None, steps, mode))
    for name in D.TABLE_NAMES:
        ref = g[f"{mode}_{steps}_{name}"]
        got = tabs[name].numpy()
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, ref, err_msg=name)


def test_schedule_appendix_b_scalars(golden):
    lin = D.make_tables(D.get_betas(None, None, 1000, "linear"))
    cos = D.make_tables(D.get_betas(None, None, 1000, "cosine"))
    assert lin["betas"][0].item() == pytest.approx(9.99999975e-05, rel=1e-7)
    assert lin["alphas_hat"][999].item() == pytest.approx(4.03583035e-05, rel=1e-6)
    assert cos["sqrt_recip_alphas_cumprod"][999].item() == pytest.approx(20291.3027, rel=1e-6)
    assert cos["betas"][999].item() == pytest.approx(0.999000013, rel=1e-7)
    assert lin["posterior_variance"][0].item() == 0.0
    assert float(cos["alphas_hat_sqrt"].double().sum()) == pytest.approx(633.262086474, rel=1e-7)
    assert int((cos["betas"] >= 0.999).sum()) == 1
    np.testing.assert_array_equal(D.get_betas(1e-3, 5e-2, 100, "linear").numpy(), golden["schedules"]["linear_custom_100"])
    with pytest.raises(ValueError):
        D.get_betas(mode="nope")


def test_kats(golden):
    g = golden["kats"]
    tt = T(g["temb_t"])
    np.testing.assert_allclose(D.timestep_embedding(tt, 128).numpy(), g["temb_128"], rtol=0, atol=0)
    np.testing.assert_allclose(D.timestep_embedding(tt, 32).numpy(), g["temb_32"], rtol=0, atol=0)
    np.testing.assert_allclose(D.timestep_embedding(tt.float(), 33).numpy(), g["temb_33"], rtol=0, atol=0)
    np.testing.assert_array_equal(
        D.normal_kl(T(g["kl_m1"]), T(g["kl_lv1"]), T(g["kl_m2"]), T(g["kl_lv2"])).numpy(), g["kl_out"])
    np.testing.assert_array_equal(D.normal_kl(T(g["kl_m1"]), T(g["kl_lv1"]), 0.0, 0.0).numpy(), g["kl_scalar_out"])
    np.testing.assert_array_equal(
        D.discretized_gaussian_log_likelihood(T(g["dll_x"]), T(g["dll_means"]), T(g["dll_ls"])).numpy(), g["dll_out"])
    np.testing.assert_array_equal(D.approx_standard_normal_cdf(T(g["cdf_in"])).numpy(), g["cdf_out"])
    np.testing.assert_array_equal(D.mean_flat(T(g["kl_m1"])).numpy(), g["mean_flat_out"])
    # Appendix B numbers typed into SURVEY.md
    np.testing.assert_allclose(g["appB_kl"], [0.19753113389, 0.92600548267, 0.05629798770], rtol=2e-6)
    np.testing.assert_allclose(g["appB_dll"], [-3.6194620132, -3.8357324600, -27.6310214996, -1.5579527617], rtol=2e-6)
    e = D.timestep_embedding(torch.tensor([1, 500, 1000]), 128)
    assert float(e.double().sum()) == pytest.approx(116.21374635442771, rel=1e-6)
    assert e[0, 0].item() == pytest.approx(0.5403023362, rel=1e-6) and e[0, 64].item() == pytest.approx(0.8414709568, rel=1e-6)


UNET_CASES = [("tiny", TINY, 16, 1), ("tiny_ss", dict(TINY, use_scale_shift_norm=True), 16, 1),
              ("tiny_sigma", TINY, 16, 2), ("small_grey28", MODEL_CONFIGS["unet_small_grey"], 28, 1),
              ("small_grey32", MODEL_CONFIGS["unet_small_grey"], 32, 1)]


@pytest.mark.parametrize("tag,cfg,res,out_mult", UNET_CASES)
def test_unet_forward_and_grads(golden, tag, cfg, res, out_mult):
    g = golden["unet"]
    arch = arch_from_config(res, **{k: v for k, v in cfg.items() if k != "name"}, learn_sigma=(out_mult == 2))
    P = make_params(arch, seed=11)
    for p in P.values():
        p.requires_grad_(True)
    _, t, noise = synth_batch(3, 2, cfg["in_channels"], res, 1000)
    y = unet_forward(P, arch, noise, t)
    np.testing.assert_allclose(y.detach().numpy(), g[f"{tag}_y"], rtol=1e-4, atol=2e-5)
    y2 = unet_forward(P, arch, noise, t.float())
    np.testing.assert_allclose(y2.detach().numpy(), g[f"{tag}_y_float_t"], rtol=1e-4, atol=2e-5)
    gy = T(np.random.RandomState(5).standard_normal(tuple(y.shape)).astype(np.float32))
    (y * gy).sum().backward()
    names = list(g[f"{tag}_grad_names"])
    assert names == list(P.keys())
    norms = np.array([float(P[n].grad.double().norm()) for n in names])
    np.testing.assert_allclose(norms, g[f"{tag}_grad_norms"], rtol=2e-4, atol=1e-4)  # atol: biases feeding a 1-channel-per-group GN have mathematically zero grad
    for key in g.files:
        if key.startswith(f"{tag}_grad::"):
            n = key.split("::")[1]
            np.testing.assert_allclose(P[n].grad.numpy(), g[key], rtol=2e-3, atol=2e-5 * max(1.0, float(np.abs(g[key]).max())))


BIG_CASES = [("cifar", MODEL_CONFIGS["unet"], 32, 1), ("cifar_sigma", MODEL_CONFIGS["unet"], 32, 2),
             ("celeba64", MODEL_CONFIGS["unet_celeba"], 64, 1)]


@pytest.mark.parametrize("tag,cfg,res,out_mult", BIG_CASES)
def test_unet_headline_architectures_forward_and_grads(golden, tag, cfg, res, out_mult):
    """The oracle on BASELINE configs[1] (CIFAR UNet) and the CelebA-64 UNet themselves, batch 1, against outputs
    and gradients of the unmodified reference (tests/golden/unet_big.npz, oracle/gen_golden.py:gen_unet_big)."""
    g = golden["unet_big"]
    arch = arch_from_config(res, **{k: v for k, v in cfg.items() if k != "name"}, learn_sigma=(out_mult == 2))
    P = make_params(arch, seed=11)
    for p in P.values():
        p.requires_grad_(True)
    _, t, noise = synth_batch(3, 1, cfg["in_channels"], res, 1000)
    y = unet_forward(P, arch, noise, t)
    np.testing.assert_allclose(y.detach().numpy(), g[f"{tag}_y"], rtol=2e-4, atol=5e-5)
    gy = T(np.random.RandomState(5).standard_normal(tuple(y.shape)).astype(np.float32))
    (y * gy).sum().backward()
    names = list(g[f"{tag}_grad_names"])
    assert names == list(P.keys())
    norms = np.array([float(P[n].grad.double().norm()) for n in names])
    np.testing.assert_allclose(norms, g[f"{tag}_grad_norms"], rtol=5e-4, atol=2e-4)
    for key in g.files:
        if key.startswith(f"{tag}_grad::"):
            n = key.split("::")[1]
            np.testing.assert_allclose(P[n].grad.numpy(), g[key], rtol=5e-3,
                                       atol=5e-5 * max(1.0, float(np.abs(g[key]).max())))


def test_param_census_matches_survey():
    # SURVEY.md Appendix A
    for name, res, nparams, flops in [("unet_small_grey", 28, 1062497, 373418240), ("unet_small_grey", 32, 1062497, 487915520),
                                      ("unet", 32, 49062787, 16759390208), ("unet_celeba", 64, 115938691, 79434219520)]:
        cfg = MODEL_CONFIGS[name]
        arch = arch_from_config(res, **{k: v for k, v in cfg.items() if k != "name"})
        assert sum(int(np.prod(s)) for s in param_shapes(arch).values()) == nparams
        assert fwd_flops_per_image(arch, res) == flops
    cfg = MODEL_CONFIGS["unet"]
    arch = arch_from_config(32, **{k: v for k, v in cfg.items() if k != "name"}, learn_sigma=True)
    assert fwd_flops_per_image(arch, 32) == 16766468096


@pytest.mark.parametrize("mode", ["linear", "cosine"])
def test_engine_math(golden, mode):
    g = golden["engine"]
    cfg = MODEL_CONFIGS["unet_small_grey"]
    arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
    P = make_params(arch, seed=21)
    diff = D.DiffusionRef(1000, mode=mode)
    x0, t, noise = T(g["x0"]), T(g["t"]), T(g["noise"])
    x_t = diff.q_sample(x0, noise, t)
    np.testing.assert_array_equal(x_t.numpy(), g[f"{mode}_x_t"])
    eps = T(g[f"{mode}_eps"])
    with torch.no_grad():
        np.testing.assert_allclose(unet_forward(P, arch, x_t, t).numpy(), g[f"{mode}_eps"], rtol=1e-4, atol=2e-5)
    loss, per = diff.loss_simple(eps, noise)
    assert loss.item() == pytest.approx(float(g[f"{mode}_loss"]), rel=1e-6)
    wl, _ = diff.loss_simple(eps, noise, T(g[f"{mode}_w"]))
    assert wl.dtype == torch.float64 and wl.item() == pytest.approx(float(g[f"{mode}_wloss"]), rel=1e-6)
    for clip in (False, True):
        np.testing.assert_array_equal(diff.model_mean(x_t, t, eps, clip=clip).numpy(), g[f"{mode}_mean_clip{int(clip)}"])
    np.testing.assert_array_equal(diff.xstart_from_eps(x_t, t, eps).numpy(), g[f"{mode}_xstart"])
    pm, pv = diff.q_posterior(t, x0, x_t)
    np.testing.assert_array_equal(pm.numpy(), g[f"{mode}_qpost_mean"])
    np.testing.assert_array_equal(pv.numpy(), g[f"{mode}_qpost_var"])


@pytest.mark.parametrize("mode", ["linear", "cosine"])
def test_engine_one_adam_step(golden, mode):
    g = golden["engine"]
    cfg = MODEL_CONFIGS["unet_small_grey"]
    arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
    P = {k: v.requires_grad_(True) for k, v in make_params(arch, seed=21).items()}
    diff = D.DiffusionRef(1000, mode=mode)
    loss, _ = E.train_loss(P, arch, diff, T(g["x0"]), T(g["t"]), T(g["noise"]))
    assert loss.item() == pytest.approx(float(g[f"{mode}_loss"]), rel=2e-5)
    opt = torch.optim.Adam(list(P.values()), lr=1e-3)
    loss.backward()
    gn = torch.norm(torch.stack([p.grad.norm(2) for p in P.values()]), 2).item()
    assert gn == pytest.approx(float(g[f"{mode}_gradnorm"]), rel=2e-4)
    opt.step()
    for key in g.files:
        if key.startswith(f"{mode}_after_step::"):
            np.testing.assert_allclose(P[key.split("::")[1]].detach().numpy(), g[key], rtol=0, atol=2e-4)


@pytest.mark.parametrize("mode", ["linear", "cosine"])
@pytest.mark.parametrize("sigma_mode", ["beta", "beta_tilde"])
@pytest.mark.parametrize("clip", [True, False])
def test_sample_chain_50(golden, mode, sigma_mode, clip):
    g = golden["engine"]
    cfg = MODEL_CONFIGS["unet_small_grey"]
    arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
    P = make_params(arch, seed=21)
    diff = D.DiffusionRef(1000, mode=mode, sigma_mode=sigma_mode)
    out = E.sample_chain(P, arch, diff, T(g["chain_xT"]).clone(), 50, T(g["chain_zs"]), steps_to_return=(25, 10, 1), clip=clip)
    ref = g[f"{mode}_{sigma_mode}_clip{int(clip)}_chain"]
    np.testing.assert_allclose(out.numpy(), ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("mode", ["linear", "cosine"])
def test_chain_mean_only_and_nll(golden, mode):
    g = golden["engine"]
    cfg = MODEL_CONFIGS["unet_small_grey"]
    arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
    P = make_params(arch, seed=21)
    diff = D.DiffusionRef(1000, mode=mode)
    out = E.sample_chain(P, arch, diff, T(g["chain_xT"]).clone(), 20, None, steps_to_return=(1,), clip=True, mean_only=True)
    np.testing.assert_allclose(out.numpy(), g[f"{mode}_chain_mean_only"], rtol=1e-3, atol=1e-3)
    diff20 = D.DiffusionRef(20, mode=mode)
    torch.manual_seed(77)
    nll = E.calculate_likelihood(P, arch, diff20, T(g["x0"]))
    assert nll["L_0"].item() == pytest.approx(float(g[f"{mode}_nll20_L0"]), rel=1e-4)
    assert nll["L_T"].item() == pytest.approx(float(g[f"{mode}_nll20_LT"]), rel=1e-5)
    np.testing.assert_allclose(nll["L_intermediate"].numpy(), g[f"{mode}_nll20_Lint"], rtol=1e-4)
    np.testing.assert_allclose(torch.stack(nll["L_intermediate_list"]).numpy(), g[f"{mode}_nll20_Lint_list"], rtol=1e-4, atol=1e-6)
    assert nll["nll"].item() == pytest.approx(float(g[f"{mode}_nll20_nll"]), rel=1e-4)


@pytest.mark.parametrize("mode", ["linear", "cosine"])
def test_hybrid_composition(golden, mode):
    """Learned-variance extension (parity UNPINNED by the reference; see oracle/__init__.py)."""
    g = golden["hybrid"]
    diff = D.DiffusionRef(1000, mode=mode)
    x0, t, noise = T(g["x0"]), T(g["t"]), T(g["noise"])
    x_t = diff.q_sample(x0, noise, t)
    np.testing.assert_array_equal(x_t.numpy(), g[f"{mode}_x_t"])
    mo = T(g[f"{mode}_model_out"]).clone().requires_grad_(True)
    loss, per = diff.loss_hybrid(x0, x_t, t, noise, mo)
    eps, v = mo.chunk(2, dim=1)
    np.testing.assert_allclose(diff.vb_term(x0, x_t, t, eps, v).detach().numpy(), g[f"{mode}_vb"], rtol=1e-5)
    np.testing.assert_allclose(per.detach().numpy(), g[f"{mode}_per"], rtol=1e-5)
    loss.backward()
    np.testing.assert_allclose(mo.grad.numpy(), g[f"{mode}_grad_model_out"], rtol=1e-4, atol=1e-7)
