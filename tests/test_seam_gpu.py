"""The operator seam (nn.py) used the way the reference uses it: stand-alone SiLU on feature maps inside
``Sequential(normalization, SiLU, conv)`` (src/modules/unet.py:146-150), the reference's OWN UNetModel class
assembled from this package's factories, activation checkpointing, and optimizer checkpoints."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from oracle.gen_golden import TINY, synth_batch
from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params
from _parity import within

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_standalone_silu_between_norm_and_conv():
    from probabilisticdeepdiffusionmodels_b200 import nn as N
    torch.manual_seed(0)
    seq = torch.nn.Sequential(N.normalization(64), N.SiLU(), N.conv_nd(2, 64, 96, 3, padding=1)).cuda()
    with torch.no_grad():
        seq[0].weight.uniform_(0.5, 1.5)
        seq[0].bias.uniform_(-0.5, 0.5)
    x = torch.randn(2, 64, 16, 16)
    gy = torch.randn(2, 96, 16, 16)
    xr = x.clone().requires_grad_(True)
    ref = TF.conv2d(TF.silu(TF.group_norm(xr, 32, seq[0].weight.detach().cpu(), seq[0].bias.detach().cpu(), 1e-5)),
                    seq[2].weight.detach().cpu(), seq[2].bias.detach().cpu(), padding=1)
    ref.backward(gy)
    xd = x.cuda().requires_grad_(True)
    y = seq(xd)
    assert y.shape == (2, 96, 16, 16)
    y.float().backward(gy.cuda())
    assert rel(y.float(), ref.detach()) < 8e-3
    assert rel(xd.grad, xr.grad) < 2e-2
    # and on its own: values and gradient of the elementwise kernel
    h = torch.randn(3, 32, 8, 8, device="cuda").bfloat16().requires_grad_(True)
    s = N.SiLU()(h)
    s.float().sum().backward()
    hf = h.detach().float().requires_grad_(True)
    TF.silu(hf).sum().backward()
    assert rel(s.float(), TF.silu(h.detach().float())) < 4e-3 and rel(h.grad.float(), hf.grad) < 6e-3


def _reference_unet_module():
    """src/modules/unet.py of the reference, executed with ITS relative imports bound to this package's nn.py."""
    for root in (os.environ.get("PDDM_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        path = os.path.join(root, "src", "modules", "unet.py") if root else None
        if path and os.path.exists(path):
            break
    else:
        pytest.skip("reference src/modules/unet.py not available (baseline/_ref is filled by __graft_entry__.build())")
    from probabilisticdeepdiffusionmodels_b200 import nn as our_nn
    pkg = types.ModuleType("refseam")
    pkg.__path__ = []
    fp16 = types.ModuleType("refseam.fp16_util")
    fp16.convert_module_to_f16 = fp16.convert_module_to_f32 = lambda m: None
    sys.modules.update({"refseam": pkg, "refseam.nn": our_nn, "refseam.fp16_util": fp16})
    spec = importlib.util.spec_from_file_location("refseam.unet", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refseam.unet"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("tag,cfg,res", [("tiny", TINY, 16), ("small_grey28", MODEL_CONFIGS["unet_small_grey"], 28)])
def test_reference_unet_class_on_this_seam(golden, tag, cfg, res):
    """the reference's UNetModel, untouched, built from this package's conv_nd / linear /
    normalization / SiLU / timestep_embedding / checkpoint -- output and a parameter gradient against the fixture of
    the all-reference model.  (Attention here is the reference's own einsum code on bf16 tensors.)"""
    mod = _reference_unet_module()
    g = golden["unet"]
    kw = {k: v for k, v in cfg.items() if k != "name"}
    arch = arch_from_config(res, **kw)
    m = mod.UNetModel(in_channels=kw["in_channels"], model_channels=kw["model_channels"],
                      out_channels=kw["in_channels"], num_res_blocks=kw["num_res_blocks"],
                      attention_resolutions=arch["attention_ds"], dropout=0, channel_mult=kw["channel_mult"],
                      num_heads=kw["num_heads"], use_scale_shift_norm=kw["use_scale_shift_norm"])
    m.load_state_dict(make_params(arch, seed=11))
    m = m.cuda()
    _, t, noise = synth_batch(3, 2, cfg["in_channels"], res, 1000)
    y = m(noise.cuda(), t.cuda())
    assert tuple(y.shape) == tuple(g[f"{tag}_y"].shape)
    within(f"reference UNetModel on the nn seam [{tag}] eps rel-L2", rel(y.float(), g[f"{tag}_y"]), 3e-2, f"{tag}_eps_rel")
    gy = torch.from_numpy(np.random.RandomState(5).standard_normal(tuple(y.shape)).astype(np.float32)).cuda()
    (y.float() * gy).sum().backward()
    got = dict(m.named_parameters())
    key = f"{tag}_grad::out.2.weight"
    within(f"reference UNetModel on the nn seam [{tag}] d(out.2.weight) rel-L2", rel(got["out.2.weight"].grad, g[key]), 6e-2)


def test_use_checkpoint_matches_plain_backward():
    """use_checkpoint=True (src/modules/nn.py:125-171): same output, same gradients -- including the timestep-embedding
    MLP, whose gradient has to flow through the checkpointed blocks."""
    from probabilisticdeepdiffusionmodels_b200.modules import get_unet
    cfg = {k: v for k, v in TINY.items() if k != "name"}
    arch = arch_from_config(16, **cfg)
    P_ = make_params(arch, seed=4)
    _, t, noise = synth_batch(5, 2, 3, 16, 1000)
    gy = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(1)).cuda()
    outs, grads = [], []
    for ck in (False, True):
        m = get_unet(16, **{**cfg, "use_checkpoint": ck})
        m.load_state_dict(P_)
        m = m.cuda()
        y = m.forward_ops(noise.cuda(), t.cuda())
        (y * gy).sum().backward()
        outs.append(y.detach())
        grads.append({n: p.grad.clone() for n, p in m.named_parameters()})
    assert torch.equal(outs[0], outs[1])
    for n in grads[0]:
        assert grads[1][n] is not None, n
        # the checkpointed model runs each block's emb projection on its own (bf16 d(emb) summed block by block by
        # autograd) while the plain model batches them into one GEMM: same math, different bf16 rounding order
        assert rel(grads[1][n], grads[0][n]) < 2e-2 or float(grads[0][n].norm()) < 1e-4, n  # (mathematically zero: rounding noise)
    assert float(grads[1]["time_embed.0.weight"].norm()) > 0


def test_fused_adam_checkpoint_resume():
    """state_dict() carries the device-side step count (torch.optim.Adam's ``step`` entry): a resumed optimizer
    continues the bias correction instead of restarting it, and torch Adam's own checkpoints load."""
    from probabilisticdeepdiffusionmodels_b200.optim import FusedAdam
    torch.manual_seed(0)
    w0 = torch.randn(1000, device="cuda")
    gs = [torch.randn(1000, device="cuda") for _ in range(4)]

    def run(opt_cls, n, w, state=None):
        p = torch.nn.Parameter(w.clone())
        opt = opt_cls([p], lr=1e-2)
        if state is not None:
            opt.load_state_dict(state)
        for g_ in gs[4 - n:]:
            p.grad = g_.clone()
            opt.step()
        return p.detach().clone(), opt

    full, _ = run(FusedAdam, 4, w0)
    p = torch.nn.Parameter(w0.clone())
    opt3 = FusedAdam([p], lr=1e-2)
    for g_ in gs[:3]:
        p.grad = g_.clone()
        opt3.step()
    sd = opt3.state_dict()
    assert float(sd["state"][0]["step"]) == 3.0
    p2 = torch.nn.Parameter(p.detach().clone())
    opt4 = FusedAdam([p2], lr=1e-2)
    opt4.load_state_dict(__import__("copy").deepcopy(sd))
    p2.grad = gs[3].clone()
    opt4.step()
    assert torch.equal(p2.detach(), full)
    # torch.optim.Adam's checkpoint -> FusedAdam
    pt = torch.nn.Parameter(w0.clone())
    ot = torch.optim.Adam([pt], lr=1e-2)
    for g_ in gs[:3]:
        pt.grad = g_.clone()
        ot.step()
    p3 = torch.nn.Parameter(pt.detach().clone())
    o3 = FusedAdam([p3], lr=1e-2)
    o3.load_state_dict(__import__("copy").deepcopy(ot.state_dict()))  # (load_state_dict may alias same-device tensors)
    p3.grad = gs[3].clone()
    o3.step()
    pt.grad = gs[3].clone()
    ot.step()
    assert rel(p3.detach(), pt.detach()) < 1e-6
