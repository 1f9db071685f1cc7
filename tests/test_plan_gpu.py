"""The hand-scheduled forward/backward plan (plan.py) against the op-by-op path (torch.ops.pddm.* under autograd) on
the same weights: same network, two schedules.  Both compute in bf16 with fp32 accumulation, so they agree to
rounding; the reference itself pins both through tests/test_model_gpu.py (``model(x, t)`` takes the plan)."""
import numpy as np
import pytest
import torch

from oracle.gen_golden import TINY, synth_batch
from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params

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
    return m.cuda()


@pytest.mark.parametrize("tag,cfg,res,B,ls", [("tiny", TINY, 16, 3, True), ("grey28", MODEL_CONFIGS["unet_small_grey"], 28, 5, False),
                                              ("cifar", MODEL_CONFIGS["unet"], 32, 2, True),
                                              ("celebahq", MODEL_CONFIGS["unet_celebahq"], 64, 1, False)])
def test_plan_equals_op_path(tag, cfg, res, B, ls):
    from probabilisticdeepdiffusionmodels_b200 import plan as PL
    m = build(cfg, res, 13, ls)
    _, t, noise = synth_batch(4, B, cfg["in_channels"], res, 1000)
    x, t = noise.cuda(), t.cuda()
    assert PL.plan_for(m, x) is not None
    y = m(x, t)
    gy = torch.from_numpy(np.random.RandomState(2).standard_normal(tuple(y.shape)).astype(np.float32)).cuda()
    (y * gy).sum().backward()
    g_plan = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    y_ops = m.forward_ops(x, t)
    (y_ops * gy).sum().backward()
    e = rel(y, y_ops)
    print(f"{tag}: plan vs op path, output rel-L2 {e:.2e}")
    assert e < 6e-3  # two bf16 schedules of the same network (GroupNorm kernels differ in summation order / SiLU form)
    worst = ("", 0.0)
    for n, p in m.named_parameters():
        assert p.grad is not None and g_plan[n] is not None, n
        if float(p.grad.norm()) > 1e-3:
            en = rel(g_plan[n], p.grad)
            if en > worst[1]:
                worst = (n, en)
    print(f"{tag}: worst parameter-gradient rel-L2 {worst[1]:.2e} ({worst[0]})")
    assert worst[1] < 4e-2, worst


def test_plan_backward_accumulates_like_any_autograd_node():
    m = build(TINY, 16, 3)
    _, t, noise = synth_batch(5, 2, 3, 16, 1000)
    x, t = noise.cuda(), t.cuda()
    m(x, t).square().sum().backward()
    g1 = [p.grad.clone() for p in m.parameters()]
    m(x, t).square().sum().backward()  # second backward without zero_grad: p.grad doubles
    for a, p in zip(g1, m.parameters()):
        torch.testing.assert_close(p.grad, 2 * a, rtol=1e-5, atol=1e-7)
    # two forwards in flight before their backwards (e.g. a validation pass in between)
    m.zero_grad(set_to_none=True)
    ya, yb = m(x, t), m(x * 0.5, t)
    with torch.no_grad():
        m(x, t)
    (ya.square().sum() + yb.square().sum()).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())


def test_sampling_graph_follows_weight_updates():
    """ADVICE (round 1, high): sample, change the weights (optimizer step through raw pointers, in-place update,
    load_state_dict), sample again -- the captured reverse-step graph must use the new weights."""
    from probabilisticdeepdiffusionmodels_b200 import Engine
    from probabilisticdeepdiffusionmodels_b200.optim import FusedAdam
    cfg = MODEL_CONFIGS["unet_small_grey"]
    arch = arch_from_config(28, **{k: v for k, v in cfg.items() if k != "name"})
    eng = Engine(dict(cfg), {"lr": 1e-2}, diffusion_steps=1000, mode="linear", resolution=28,
                 clip_while_generating=True, log_loss_per_t=False)
    eng.model.load_state_dict(make_params(arch, seed=21))
    eng = eng.cuda()
    xT = torch.randn(3, 1, 28, 28, generator=torch.Generator().manual_seed(1)).cuda()

    def sample(graph):
        g = torch.Generator(device="cuda").manual_seed(5)
        with torch.no_grad():
            return eng.sample_from_step(xT.clone(), 8, generator=g, use_graph=graph)

    a0 = sample(True)
    assert torch.equal(a0, sample(False))
    # 1. an optimizer step through the fused kernel (raw pointers: no version bump)
    opt = FusedAdam(eng.model.parameters(), lr=1e-2)
    x0 = torch.rand(3, 1, 28, 28, device="cuda") * 2 - 1
    loss, _ = eng.loss_on(x0, torch.tensor([5, 500, 900], device="cuda"), torch.randn_like(x0))
    loss.backward()
    opt.step()
    a1 = sample(True)
    assert not torch.equal(a1, a0)
    assert torch.equal(a1, sample(False))
    # 2. in-place update + load_state_dict
    with torch.no_grad():
        for p in eng.model.parameters():
            p.mul_(1.01)
    a2 = sample(True)
    assert torch.equal(a2, sample(False)) and not torch.equal(a2, a1)
    eng.model.load_state_dict(make_params(arch, seed=22))
    a3 = sample(True)
    assert torch.equal(a3, sample(False)) and not torch.equal(a3, a2)


def test_gradient_arena_is_flat_and_owned_by_the_step():
    """capture_train_step on the plan: p.grad are views of ONE flat fp32 buffer (what the all-reduce runs on)."""
    from probabilisticdeepdiffusionmodels_b200 import Engine
    cfg = MODEL_CONFIGS["unet_small_grey"]
    eng = Engine(dict(cfg), {"lr": 1e-3}, diffusion_steps=1000, mode="cosine", resolution=28, learn_sigma=True,
                 log_loss_per_t=False).cuda()
    x = torch.rand(8, 1, 28, 28, device="cuda") * 2 - 1
    step = eng.capture_train_step(tuple(x.shape))
    plan = step.state["plan"]
    assert plan is not None
    l0 = float(step(x))
    lo, hi = plan.grad_arena.data_ptr(), plan.grad_arena.data_ptr() + plan.grad_arena.numel() * 4
    for p in eng.model.parameters():
        assert p.grad is not None and lo <= p.grad.data_ptr() < hi
    for _ in range(30):
        l1 = float(step(x))
    assert np.isfinite(l1) and l1 < l0
