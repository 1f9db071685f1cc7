"""north_star's per-call bar -- eps / variance outputs and the loss within 1e-3 relative of the reference -- met by the
opt-in high-precision forward (highprec.py: fp32 activations, GEMM operands split into two bf16 terms on the same
tcgen05 kernels).  The bf16 product path sits at ~1e-2 on eps, like the reference's own network under bf16 autocast
(tests/golden/floors.npz); this mode shows that the remaining error is operand precision and nothing else."""
import numpy as np
import pytest
import torch

from oracle.gen_golden import TINY, synth_batch
from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params
from _parity import within

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(cfg, res, seed, learn_sigma=False):
    from probabilisticdeepdiffusionmodels_b200.modules import get_unet
    kw = {k: v for k, v in cfg.items() if k != "name"}
    arch = arch_from_config(res, **kw, learn_sigma=learn_sigma)
    m = get_unet(res, **kw, learn_sigma=learn_sigma)
    m.load_state_dict(make_params(arch, seed=seed))
    return m.cuda().eval()


@pytest.mark.parametrize("tag,cfg,res,ls,fixture,b", [
    ("tiny", TINY, 16, False, "unet", 2), ("tiny_sigma", TINY, 16, True, "unet", 2),
    ("small_grey28", MODEL_CONFIGS["unet_small_grey"], 28, False, "unet", 2),
    ("small_grey32", MODEL_CONFIGS["unet_small_grey"], 32, False, "unet", 2),
    ("cifar", MODEL_CONFIGS["unet"], 32, False, "unet_big", 1), ("cifar_sigma", MODEL_CONFIGS["unet"], 32, True, "unet_big", 1),
    ("celeba64", MODEL_CONFIGS["unet_celeba"], 64, False, "unet_big", 1)])
def test_high_precision_eps_within_1e_3(golden, tag, cfg, res, ls, fixture, b):
    g = golden[fixture]
    m = build(cfg, res, 11, ls)
    _, t, noise = synth_batch(3, b, cfg["in_channels"], res, 1000)
    with torch.no_grad():
        y16 = m(noise.cuda(), t.cuda())
        m.high_precision = True
        y = m(noise.cuda(), t.cuda())
    assert y.dtype == torch.float32 and tuple(y.shape) == tuple(g[f"{tag}_y"].shape)
    print(f"[parity] unet[{tag}] bf16 product path eps rel-L2 {rel(y16, g[f'{tag}_y']):.3e}")
    # north_star's bar is 1e-3; measured 1.2e-5 ... 1.6e-5 on every architecture, asserted at 3x that
    within(f"unet[{tag}] HIGH-PRECISION eps rel-L2 (north_star 1e-3)", rel(y, g[f"{tag}_y"]), 5e-5, f"{tag}_eps_rel")


@pytest.mark.parametrize("mode", ["linear", "cosine"])
def test_high_precision_engine_eps_and_loss_within_1e_3(golden, mode):
    from probabilisticdeepdiffusionmodels_b200 import Engine
    g = golden["engine"]
    cfg = MODEL_CONFIGS["unet_small_grey"]
    arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
    eng = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=1000, mode=mode, resolution=28, clip_while_generating=True)
    eng.model.load_state_dict(make_params(arch, seed=21))
    eng = eng.to("cuda")
    eng.eval()
    eng.model.high_precision = True
    T = lambda a: torch.from_numpy(np.asarray(a))
    x0, t, noise = T(g["x0"]).cuda(), T(g["t"]).cuda(), T(g["noise"]).cuda()
    with torch.no_grad():
        x_t = eng.get_q_t(x0, noise, t)
        eps = eng.model(x_t, t)
        loss = eng.get_loss(eps, noise, x0, x_t, t=t, update_loss_log=False)
    within(f"engine[{mode}] HIGH-PRECISION eps rel-L2 (north_star 1e-3)", rel(eps, g[f"{mode}_eps"]), 5e-5,
           f"engine_{mode}_eps_rel")
    within(f"engine[{mode}] HIGH-PRECISION loss relative deviation (north_star 1e-3)",
           abs(loss.item() - float(g[f"{mode}_loss"])) / float(g[f"{mode}_loss"]), 5e-6, f"engine_{mode}_loss_rel")


def test_high_precision_cost():
    """What the mode costs on the CIFAR UNet at B=16 (printed; asserted only to be a sane multiple)."""
    m = build(MODEL_CONFIGS["unet"], 32, 5)
    _, t, noise = synth_batch(7, 16, 3, 32, 1000)
    x, tt = noise.cuda(), t.cuda()

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3

    with torch.no_grad():
        a = timed(lambda: m(x, tt))
        m.high_precision = True
        b = timed(lambda: m(x, tt))
    print(f"[parity] CIFAR UNet B=16 eager forward: bf16 product path {a:.2f} ms, high-precision mode {b:.2f} ms ({b / a:.1f}x)")
    assert b < 40 * a


@pytest.mark.parametrize("tag,mode", [("small_linear", "linear"), ("small_cosine", "cosine")])
def test_high_precision_1000_step_trajectory(golden, tag, mode):
    """The 1000-step chain of tests/test_trajectory_gpu.py in high-precision mode: the per-pixel error against the
    reference's fp32 chain drops from the bf16 level (max 3e-2 ... 5e-2, equal to the reference's own bf16 drift) by two
    orders of magnitude -- the trajectory error of the product path is operand rounding, not the algorithm."""
    from probabilisticdeepdiffusionmodels_b200 import Engine
    g = golden["traj1000"]
    cfg = MODEL_CONFIGS["unet_small_grey"]
    eng = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=1000, mode=mode, resolution=28, clip_while_generating=True,
                 sigma_mode="beta")
    arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
    eng.model.load_state_dict(make_params(arch, seed=21))
    eng = eng.to("cuda")
    eng.model.high_precision = True
    xT = torch.from_numpy(g[f"{tag}_xT"])
    steps = tuple(int(s) for s in g["steps"])
    gen = torch.Generator().manual_seed(int(g[f"{tag}_zseed"]))
    zs = torch.stack([torch.randn(tuple(xT.shape), generator=gen) for _ in range(999)]).cuda()
    out = eng.sample_and_return_steps(xT.cuda(), t_start=1000, steps_to_return=steps, fixed_noise=zs)
    err = np.abs(out.numpy() - g[f"{tag}_chain"])
    print(f"[parity] traj1000[{tag}] HIGH-PRECISION: max |d| {err.max():.3e}  mean |d| {err.mean():.3e} "
          f"(reference bf16 drift: max {g[f'{tag}_floor_max'].max():.3e})")
    within(f"traj1000[{tag}] HIGH-PRECISION max |delta| per pixel", float(err.max()), 2e-4)
    within(f"traj1000[{tag}] HIGH-PRECISION mean |delta| per pixel", float(err.mean()), 1e-5)
