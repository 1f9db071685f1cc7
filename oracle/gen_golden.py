"""Oracle (test infrastructure): generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (``python -m oracle.gen_golden``): imports ``src.engine`` /
``src.modules`` / ``src.utils`` from /root/reference behind ``oracle.ref_shims`` and records their
outputs on seeded inputs.  The committed fixtures are what pins the oracle
(tests/test_oracle_golden.py) and, through it, the CUDA path on the GPU box where
/root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# a deliberately odd little architecture that exercises every layer kind: attention inside the
# down/up paths (ds=2 at 16x16), a 1x1 skip, Down/Upsample, the always-present middle attention
TINY = dict(name="unet", in_channels=3, model_channels=32, num_res_blocks=1, attention_resolutions=[8],
            dropout=0, channel_mult=[1, 2], conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
            num_heads=2, num_heads_upsample=-1, use_scale_shift_norm=False)


def synth_batch(seed, b, c, r, T):
    rs = np.random.RandomState(seed)
    x0 = torch.from_numpy((rs.rand(b, c, r, r) * 2 - 1).astype(np.float32))
    t = torch.from_numpy(rs.randint(1, T + 1, size=(b,)).astype(np.int64))
    noise = torch.from_numpy(rs.standard_normal((b, c, r, r)).astype(np.float32))
    return x0, t, noise


def ref_model(cfg, resolution, seed, out_mult=1):
    from src.modules import get_model
    from src.modules.unet import UNetModel

    arch = arch_from_config(resolution, **{k: v for k, v in cfg.items() if k != "name"},
                            learn_sigma=(out_mult == 2))
    if out_mult == 1:
        m = get_model(resolution, dict(cfg))
    else:  # reference hard-codes learn_sigma=False (src/modules/__init__.py:34); build UNetModel directly
        m = UNetModel(in_channels=cfg["in_channels"], model_channels=cfg["model_channels"],
                      out_channels=cfg["in_channels"] * 2, num_res_blocks=cfg["num_res_blocks"],
                      attention_resolutions=arch["attention_ds"], dropout=cfg["dropout"],
                      channel_mult=cfg["channel_mult"], num_heads=cfg["num_heads"],
                      use_scale_shift_norm=cfg["use_scale_shift_norm"])
    P = make_params(arch, seed=seed)
    sd = m.state_dict()
    assert list(sd.keys()) == list(P.keys()), "state_dict key order mismatch"
    for k in sd:
        assert tuple(sd[k].shape) == tuple(P[k].shape), k
    m.load_state_dict(P)
    return m, arch, P


def gen_schedules():
    from src.engine import get_betas as ref_get_betas
    from src.engine import Engine

    out = {}
    for mode in ("linear", "cosine", "mixed"):
        for T in (1000, 50):
            eng = Engine(dict(MODEL_CONFIGS["unet_small_grey"]), {"lr": 1e-4}, diffusion_steps=T, mode=mode,
                         resolution=28)
            for name in ("betas", "alphas", "alphas_sqrt", "alphas_hat", "alphas_hat_sqrt",
                         "one_min_alphas_hat_sqrt", "alphas_hat_prev", "alphas_hat_next", "posterior_variance",
                         "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                         "posterior_mean_coef2", "denoising_coef"):
                v = getattr(eng, name)
                assert v.dtype == torch.float32
                out[f"{mode}_{T}_{name}"] = v.numpy()
    out["linear_custom_100"] = ref_get_betas(1e-3, 5e-2, 100, "linear").numpy()
    np.savez_compressed(os.path.join(GOLD, "schedules.npz"), **out)


def gen_kats():
    from src.modules.nn import timestep_embedding
    from src.utils import approx_standard_normal_cdf, discretized_gaussian_log_likelihood, normal_kl, mean_flat

    rs = np.random.RandomState(7)
    out = {}
    tt = torch.tensor([1, 2, 17, 500, 999, 1000])
    out["temb_t"] = tt.numpy()
    out["temb_128"] = timestep_embedding(tt, 128).numpy()
    out["temb_32"] = timestep_embedding(tt, 32).numpy()
    out["temb_33"] = timestep_embedding(tt.float(), 33).numpy()  # float timesteps + odd dim (zero pad)
    m1, m2 = [torch.from_numpy(rs.standard_normal((4, 3, 8, 8)).astype(np.float32)) for _ in range(2)]
    lv1, lv2 = [torch.from_numpy((rs.standard_normal((4, 3, 8, 8)) * 2 - 2).astype(np.float32)) for _ in range(2)]
    out.update(kl_m1=m1.numpy(), kl_m2=m2.numpy(), kl_lv1=lv1.numpy(), kl_lv2=lv2.numpy(),
               kl_out=normal_kl(m1, lv1, m2, lv2).numpy(),
               kl_scalar_out=normal_kl(m1, lv1, 0.0, 0.0).numpy())
    x = torch.from_numpy(np.clip(np.round((rs.rand(4, 3, 8, 8)) * 255) / 127.5 - 1, -1, 1).astype(np.float32))
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, -0.9995, 0.9995])
    means = x + torch.from_numpy((rs.standard_normal((4, 3, 8, 8)) * 0.1).astype(np.float32))
    ls = torch.from_numpy((rs.rand(4, 3, 8, 8) * 5 - 5).astype(np.float32))
    out.update(dll_x=x.numpy(), dll_means=means.numpy(), dll_ls=ls.numpy(),
               dll_out=discretized_gaussian_log_likelihood(x, means, ls).numpy(),
               cdf_in=m1.numpy() * 3, cdf_out=approx_standard_normal_cdf(m1 * 3).numpy(),
               mean_flat_out=mean_flat(m1).numpy())
    # SURVEY.md Appendix B scalars
    out["appB_kl"] = normal_kl(torch.tensor([0.1, -0.5, 0.9]), torch.tensor([-2.0, -6.0, 0.0]),
                               torch.tensor([0.0, -0.4, 1.0]), torch.tensor([-1.0, -5.0, 0.5])).numpy()
    out["appB_dll"] = discretized_gaussian_log_likelihood(
        torch.tensor([-1.0, -0.2, 0.5, 1.0]), torch.tensor([-0.9, -0.25, 0.4, 0.7]),
        torch.tensor([-3.0, -2.0, -4.0, -1.0])).numpy()
    np.savez_compressed(os.path.join(GOLD, "kats.npz"), **out)


def gen_unet():
    """eps outputs + gradients of sum(out * g) for three architectures."""
    out = {}
    cases = [
        ("tiny", TINY, 16, 2, 1),
        ("tiny_ss", dict(TINY, use_scale_shift_norm=True), 16, 2, 1),
        ("tiny_sigma", TINY, 16, 2, 2),
        ("small_grey28", MODEL_CONFIGS["unet_small_grey"], 28, 2, 1),
        ("small_grey32", MODEL_CONFIGS["unet_small_grey"], 32, 2, 1),
    ]
    for tag, cfg, res, b, out_mult in cases:
        m, arch, P = ref_model(cfg, res, seed=11, out_mult=out_mult)
        x0, t, noise = synth_batch(3, b, cfg["in_channels"], res, 1000)
        for p in m.parameters():
            p.requires_grad_(True)
        y = m(noise, t)
        rs = np.random.RandomState(5)
        g = torch.from_numpy(rs.standard_normal(tuple(y.shape)).astype(np.float32))
        (y * g).sum().backward()
        out[f"{tag}_y"] = y.detach().numpy()
        out[f"{tag}_y_float_t"] = m(noise, t.float()).detach().numpy()  # sampling passes float32 t
        names = [n for n, _ in m.named_parameters()]
        out[f"{tag}_grad_norms"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
        out[f"{tag}_grad_names"] = np.array(names)
        for n, p in m.named_parameters():
            if n in ("time_embed.0.weight", "input_blocks.0.0.weight", "input_blocks.1.0.in_layers.2.bias",
                     "middle_block.1.qkv.weight", "middle_block.1.norm.weight", "out.2.weight",
                     "output_blocks.0.0.skip_connection.weight", "output_blocks.1.0.emb_layers.1.weight",
                     "input_blocks.2.0.op.weight"):
                out[f"{tag}_grad::{n}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "unet.npz"), **out)


def gen_unet_big():
    """The headline architectures themselves, batch 1: BASELINE configs[1] (config/model/unet.yaml @32, with and
    without the learned-variance head) and config/model/unet_celeba.yaml @64.  Output, every parameter's gradient
    norm, and the small gradients in full (large ones are covered by their norms to keep the fixture small)."""
    out = {}
    cases = [("cifar", MODEL_CONFIGS["unet"], 32, 1), ("cifar_sigma", MODEL_CONFIGS["unet"], 32, 2),
             ("celeba64", MODEL_CONFIGS["unet_celeba"], 64, 1)]
    for tag, cfg, res, out_mult in cases:
        m, arch, P = ref_model(cfg, res, seed=11, out_mult=out_mult)
        x0, t, noise = synth_batch(3, 1, cfg["in_channels"], res, 1000)
        for p in m.parameters():
            p.requires_grad_(True)
        y = m(noise, t)
        g = torch.from_numpy(np.random.RandomState(5).standard_normal(tuple(y.shape)).astype(np.float32))
        (y * g).sum().backward()
        out[f"{tag}_y"] = y.detach().numpy()
        out[f"{tag}_grad_names"] = np.array([n for n, _ in m.named_parameters()])
        out[f"{tag}_grad_norms"] = np.array([float(p.grad.double().norm()) for _, p in m.named_parameters()])
        for n, p in m.named_parameters():
            if p.numel() <= 4096 and (n.endswith("norm.weight") or n.endswith("in_layers.0.weight") or
                                      n.startswith("out.") or n.startswith("input_blocks.0.")):
                out[f"{tag}_grad::{n}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "unet_big.npz"), **out)


def gen_engine():
    """Engine-level: q_sample, loss, one optimiser step, p_sample chains, NLL terms."""
    from src.engine import Engine

    out = {}
    for mode in ("linear", "cosine"):
        cfg = MODEL_CONFIGS["unet_small_grey"]
        eng = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=1000, mode=mode, resolution=28,
                     clip_while_generating=True, sigma_mode="beta")
        arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
        eng.model.load_state_dict(make_params(arch, seed=21))
        x0, t, noise = synth_batch(9, 4, 1, 28, 1000)
        t[0], t[1] = 1, 1000
        x_t = eng.get_q_t(x0, noise, t)
        eps = eng.model(x_t, t)
        loss = eng.get_loss(eps, noise, x0, x_t, t=t, update_loss_log=False)
        w = torch.from_numpy(np.random.RandomState(2).rand(4))  # float64 importance weights
        wloss = eng.get_loss(eps, noise, x0, x_t, t=t, weights=w, update_loss_log=False)
        out[f"{mode}_x_t"] = x_t.detach().numpy()
        out[f"{mode}_eps"] = eps.detach().numpy()
        out[f"{mode}_loss"] = np.array(loss.item())
        out[f"{mode}_wloss"] = np.array(wloss.item())
        out[f"{mode}_w"] = w.numpy()
        # one full optimiser step (training_step math with injected t / noise) -> parameters after Adam
        opt = torch.optim.Adam(eng.parameters(), lr=1e-3)
        opt.zero_grad()
        loss.backward()
        out[f"{mode}_gradnorm"] = np.array(float(eng.compute_grad_norm(eng.model.parameters())))
        opt.step()
        sd = eng.model.state_dict()
        for n in ("out.2.weight", "input_blocks.0.0.weight", "middle_block.1.qkv.bias"):
            out[f"{mode}_after_step::{n}"] = sd[n].detach().clone().numpy()
        eng.model.load_state_dict(make_params(arch, seed=21))

        # per-step pieces for every (clip, sigma_mode) at a few t, per-sample t vector
        for clip in (False, True):
            out[f"{mode}_mean_clip{int(clip)}"] = eng.model_mean_from_epsilon(x_t, t, eps, clip=clip).detach().numpy()
        out[f"{mode}_xstart"] = eng.xstart_from_epsilon(x_t, t, eps, clip=False).detach().numpy()
        pm, pv = eng.q_posterior(t, x0, x_t)
        out[f"{mode}_qpost_mean"], out[f"{mode}_qpost_var"] = pm.detach().numpy(), pv.numpy()

        # 50-step chain from t_start = 50 (BASELINE config 1), fixed z via a CPU generator
        for sigma_mode in ("beta", "beta_tilde"):
            eng.sigma_mode = sigma_mode
            for clip in (True, False):
                eng.clip_while_generating = clip
                rs = np.random.RandomState(33)
                xT = torch.from_numpy(rs.standard_normal((2, 1, 28, 28)).astype(np.float32))
                gen = torch.Generator().manual_seed(1234)
                zs = torch.stack([torch.randn((2, 1, 28, 28), generator=gen) for _ in range(49)])
                gen = torch.Generator().manual_seed(1234)
                steps = eng.sample_and_return_steps(xT.clone(), t_start=50, steps_to_return=(25, 10, 1),
                                                    generator=gen)
                key = f"{mode}_{sigma_mode}_clip{int(clip)}"
                out[f"{key}_chain"] = steps.numpy()
                if key == "linear_beta_clip1":
                    out["chain_xT"], out["chain_zs"] = xT.numpy(), zs.numpy()
        eng.sigma_mode, eng.clip_while_generating = "beta", True
        # mean_only chain
        steps = eng.sample_and_return_steps(torch.from_numpy(out["chain_xT"]).clone(), t_start=20,
                                            steps_to_return=(1,), mean_only=True)
        out[f"{mode}_chain_mean_only"] = steps.numpy()

        # NLL terms on a short T=20 engine (fixed variance)
        eng20 = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=20, mode=mode, resolution=28)
        eng20.model.load_state_dict(make_params(arch, seed=21))
        eng20.eval()
        torch.manual_seed(77)
        with torch.no_grad():
            nll = eng20.calculate_likelihood(x0)
        out[f"{mode}_nll20_L0"] = np.array(nll["L_0"].item())
        out[f"{mode}_nll20_LT"] = np.array(nll["L_T"].item())
        out[f"{mode}_nll20_Lint"] = nll["L_intermediate"].numpy()
        out[f"{mode}_nll20_nll"] = np.array(nll["nll"].item())
        out[f"{mode}_nll20_Lint_list"] = torch.stack(nll["L_intermediate_list"]).numpy()
    out["x0"], out["t"], out["noise"] = x0.numpy(), t.numpy(), noise.numpy()
    np.savez_compressed(os.path.join(GOLD, "engine.npz"), **out)


def gen_hybrid():
    """Learned-variance extension composed from reference functions.

    NOT a reference output (the reference has no learn_sigma path): recorded so that the oracle's and
    the CUDA path's composition can at least be checked against the same composition of the
    reference's own normal_kl / discretized_gaussian_log_likelihood / q_posterior.  Parity unpinned.
    """
    from src.engine import Engine
    from src.utils import discretized_gaussian_log_likelihood, mean_flat, normal_kl

    out = {}
    for mode in ("linear", "cosine"):
        eng = Engine(dict(TINY), {"lr": 1e-3}, diffusion_steps=1000, mode=mode, resolution=16)
        m, arch, P = ref_model(TINY, 16, seed=11, out_mult=2)
        x0, t, noise = synth_batch(13, 6, 3, 16, 1000)
        t[0], t[1], t[2] = 1, 2, 1000
        x_t = eng.get_q_t(x0, noise, t)
        mo = m(x_t, t)
        eps, v = mo.chunk(2, dim=1)
        pv = eng.posterior_variance
        plv = torch.log(torch.cat([pv[1:2], pv[1:]]))
        min_log = plv[t - 1].view(-1, 1, 1, 1)
        max_log = torch.log(eng.betas)[t - 1].view(-1, 1, 1, 1)
        frac = (v + 1) / 2
        logvar = frac * max_log + (1 - frac) * min_log
        true_mean, _ = eng.q_posterior(t, x0, x_t)
        pmean = eng.model_mean_through_start(x_t, t, eps.detach(), clip=False)
        kl = mean_flat(normal_kl(true_mean, min_log, pmean, logvar)) / np.log(2.0)
        nll0 = -mean_flat(discretized_gaussian_log_likelihood(x0, pmean, 0.5 * logvar)) / np.log(2.0)
        vb = torch.where(t == 1, nll0, kl)
        per = mean_flat(torch.square(noise - eps)) + vb
        loss = per.mean()
        gmo, = torch.autograd.grad(loss, mo)
        out[f"{mode}_model_out"] = mo.detach().numpy()
        out[f"{mode}_vb"] = vb.detach().numpy()
        out[f"{mode}_per"] = per.detach().numpy()
        out[f"{mode}_loss"] = np.array(loss.item())
        out[f"{mode}_grad_model_out"] = gmo.numpy()
        out[f"{mode}_x_t"] = x_t.numpy()
    out["x0"], out["t"], out["noise"] = x0.numpy(), t.numpy(), noise.numpy()
    np.savez_compressed(os.path.join(GOLD, "hybrid.npz"), **out)


TRAJ_STEPS = (900, 750, 500, 250, 100, 50, 20, 10, 5, 1)


def gen_traj1000(which=("small", "cifar")):
    """Full 1000-step reverse chains of the UNMODIFIED reference (src/engine.py:399-403, 510-554) with fixed x_T and a
    seeded CPU generator for the per-step z (the test re-draws the same z sequence from the same seed), plus the
    reference's OWN drift when its network runs under bf16 autocast -- the noise floor a bf16 implementation of the
    same chain can be held to.  ``python -m oracle.gen_golden traj`` (the CIFAR chains take minutes on CPU)."""
    from src.engine import Engine

    out = {"steps": np.array(TRAJ_STEPS)}
    path = os.path.join(GOLD, "traj1000.npz")
    if os.path.exists(path):
        out.update({k: v for k, v in np.load(path).items()})
    cases = []
    if "small" in which:
        cases += [("small_linear", "unet_small_grey", 28, 2, "linear", 21), ("small_cosine", "unet_small_grey", 28, 2, "cosine", 21)]
    if "cifar" in which:
        cases += [("cifar_cosine", "unet", 32, 1, "cosine", 11)]
    for tag, name, res, B, mode, pseed in cases:
        cfg = MODEL_CONFIGS[name]
        eng = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=1000, mode=mode, resolution=res,
                     clip_while_generating=True, sigma_mode="beta")
        arch = arch_from_config(res, **{k: v for k, v in cfg.items() if k != "name"})
        eng.model.load_state_dict(make_params(arch, seed=pseed))
        eng.eval()
        C = cfg["in_channels"]
        xT = torch.from_numpy(np.random.RandomState(41).standard_normal((B, C, res, res)).astype(np.float32))
        out[f"{tag}_xT"] = xT.numpy()
        out[f"{tag}_zseed"] = np.array(4321)
        gen = torch.Generator().manual_seed(4321)
        chain = eng.sample_and_return_steps(xT.clone(), t_start=1000, steps_to_return=TRAJ_STEPS, generator=gen)
        out[f"{tag}_chain"] = chain.numpy()
        print(tag, "fp32 chain done", float(chain[:, -1].abs().max()), flush=True)
        # the reference network under bf16 autocast, everything else (schedule tables, posterior math, z) unchanged
        net = eng.model

        class Autocast(torch.nn.Module):
            in_channels = C

            def forward(self, x, t):
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    return net(x, t).float()

        eng.model = Autocast()
        gen = torch.Generator().manual_seed(4321)
        chain16 = eng.sample_and_return_steps(xT.clone(), t_start=1000, steps_to_return=TRAJ_STEPS, generator=gen)
        eng.model = net
        d = (chain16 - chain).abs()
        out[f"{tag}_floor_max"] = d.amax(dim=(0, 2, 3, 4)).numpy()
        out[f"{tag}_floor_mean"] = d.mean(dim=(0, 2, 3, 4)).numpy()
        print(tag, "bf16-autocast drift of the reference: max", out[f"{tag}_floor_max"], "mean", out[f"{tag}_floor_mean"],
              flush=True)
        np.savez_compressed(path, **out)


def gen_floors():
    """The reference's OWN error when its network runs under bf16 autocast instead of fp32 (same inputs as the unet /
    unet_big / engine fixtures): relative L2 of eps and relative deviation of the loss.  This is the noise floor of a
    bf16-operand implementation of the same network; the GPU tests print it next to the measured error of the CUDA
    path.  ``python -m oracle.gen_golden floors`` -> tests/golden/floors.npz."""
    from src.engine import Engine

    out = {}

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())

    cases = [("tiny", TINY, 16, 2, 1), ("tiny_ss", dict(TINY, use_scale_shift_norm=True), 16, 2, 1),
             ("tiny_sigma", TINY, 16, 2, 2), ("small_grey28", MODEL_CONFIGS["unet_small_grey"], 28, 2, 1),
             ("small_grey32", MODEL_CONFIGS["unet_small_grey"], 32, 2, 1), ("cifar", MODEL_CONFIGS["unet"], 32, 1, 1),
             ("cifar_sigma", MODEL_CONFIGS["unet"], 32, 1, 2), ("celeba64", MODEL_CONFIGS["unet_celeba"], 64, 1, 1)]
    for tag, cfg, res, b, out_mult in cases:
        m, arch, P = ref_model(cfg, res, seed=11, out_mult=out_mult)
        x0, t, noise = synth_batch(3, b, cfg["in_channels"], res, 1000)
        with torch.no_grad():
            y = m(noise, t)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                y16 = m(noise, t).float()
        out[f"{tag}_eps_rel"] = np.array(rel(y16, y))
        print(tag, "reference bf16-autocast eps rel-L2", out[f"{tag}_eps_rel"], flush=True)
    for mode in ("linear", "cosine"):
        cfg = MODEL_CONFIGS["unet_small_grey"]
        eng = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=1000, mode=mode, resolution=28,
                     clip_while_generating=True, sigma_mode="beta")
        arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
        eng.model.load_state_dict(make_params(arch, seed=21))
        x0, t, noise = synth_batch(9, 4, 1, 28, 1000)
        t[0], t[1] = 1, 1000
        with torch.no_grad():
            x_t = eng.get_q_t(x0, noise, t)
            eps = eng.model(x_t, t)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                eps16 = eng.model(x_t, t).float()
            loss = eng.get_loss(eps, noise, x0, x_t, t=t, update_loss_log=False)
            loss16 = eng.get_loss(eps16, noise, x0, x_t, t=t, update_loss_log=False)
        out[f"engine_{mode}_eps_rel"] = np.array(rel(eps16, eps))
        out[f"engine_{mode}_loss_rel"] = np.array(abs(loss16.item() - loss.item()) / abs(loss.item()))
        print(mode, "engine eps", out[f"engine_{mode}_eps_rel"], "loss", out[f"engine_{mode}_loss_rel"], flush=True)
    np.savez_compressed(os.path.join(GOLD, "floors.npz"), **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "floors":
        ref_shims.install()
        gen_floors()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "traj":
        ref_shims.install()
        torch.set_num_threads(8)
        gen_traj1000(tuple(sys.argv[2:]) or ("small", "cifar"))
        return
    ref_shims.install()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(4)
    gen_schedules()
    gen_kats()
    gen_unet()
    gen_unet_big()
    gen_engine()
    gen_hybrid()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
