"""GPU parity of every pddm op against a plain fp32 torch evaluation of the same op (on the bf16-rounded inputs the
kernel sees) and against the CPU oracle.  Tolerances are written next to each check."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

bf16, f32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def P():
    import probabilisticdeepdiffusionmodels_b200.ops  # noqa: F401
    return torch.ops.pddm


def rnd(*shape, seed=0, scale=1.0, dtype=f32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(x):  # NCHW fp32 cpu -> NHWC bf16 cuda
    return x.permute(0, 2, 3, 1).contiguous().to("cuda", bf16)


def nchw(y):  # NHWC cuda -> NCHW fp32 cpu
    return y.float().permute(0, 3, 1, 2).cpu()


CONV_CASES = [
    # B, H, W, Cin, Cout, k, stride, upsample
    (2, 32, 32, 128, 128, 3, 1, False), (3, 16, 16, 256, 128, 3, 1, False), (4, 8, 8, 64, 96, 3, 1, False),
    (9, 4, 4, 256, 256, 3, 1, False), (2, 28, 28, 32, 64, 3, 1, False), (2, 14, 14, 96, 64, 3, 1, False),
    (3, 7, 7, 64, 64, 3, 1, False), (2, 16, 16, 384, 256, 1, 1, False), (2, 32, 32, 128, 128, 3, 2, False),
    (3, 28, 28, 32, 32, 3, 2, False), (2, 8, 8, 256, 256, 3, 1, True), (2, 14, 14, 64, 64, 3, 1, True),
    (1, 16, 16, 64, 768, 1, 1, False),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,up", CONV_CASES)
def test_conv2d_fwd_bwd(P, B, H, W, Cin, Cout, k, stride, up):
    x = rnd(B, Cin, H, W, seed=1).to(bf16).float()
    w = (rnd(Cout, Cin, k, k, seed=2) / math.sqrt(Cin * k * k)).to(bf16).float()
    b = rnd(Cout, seed=3)
    oh = (2 * H if up else H) // stride
    emb = rnd(B, Cout, seed=4)
    res = rnd(B, Cout, oh, oh * W // H, seed=5).to(bf16).float()
    gy = rnd(B, Cout, oh, oh * W // H, seed=6).to(bf16).float()
    # reference
    xr, wr, br, er, rr = [t.clone().requires_grad_(True) for t in (x, w, b, emb, res)]
    xi = TF.interpolate(xr, scale_factor=2, mode="nearest") if up else xr
    yr = TF.conv2d(xi, wr, br, stride=stride, padding=k // 2) + er[:, :, None, None] + rr
    yr.backward(gy)
    # kernel
    xd = nhwc(x).requires_grad_(True)
    wd = w.cuda().requires_grad_(True)
    if k == 1:
        wd = w.reshape(Cout, Cin, 1).cuda().requires_grad_(True)  # Conv1d-shaped weight
    bd, ed = b.cuda().requires_grad_(True), emb.cuda().requires_grad_(True)
    rd = nhwc(res).requires_grad_(True)
    y, _ = P.conv2d(xd, wd, bd, ed, rd, stride, up)
    y.backward(nhwc(gy))
    torch.cuda.synchronize()
    # output is rounded to bf16: 2^-9 relative per element -> ~3e-3 in relative L2 norm
    assert rel(nchw(y), yr.detach()) < 4e-3
    assert rel(nchw(xd.grad), xr.grad) < 6e-3  # bf16 output of the dgrad GEMM
    assert rel(wd.grad.reshape(w.shape), wr.grad) < 1e-3  # fp32 output; inputs identical
    assert rel(bd.grad, br.grad) < 1e-3
    assert rel(ed.grad, er.grad) < 1e-3
    assert rel(nchw(rd.grad), rr.grad) < 1e-6


@pytest.mark.parametrize("M,K,N", [(128, 128, 512), (8, 512, 64), (200, 512, 1024), (3, 32, 128)])
def test_linear(P, M, K, N):
    x = rnd(M, K, seed=1).to(bf16).float()
    w = (rnd(N, K, seed=2) / math.sqrt(K)).to(bf16).float()
    b = rnd(N, seed=3)
    gy = rnd(M, N, seed=4).to(bf16).float()
    xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
    TF.linear(xr, wr, br).backward(gy)
    xd = x.to("cuda", bf16).requires_grad_(True)
    wd, bd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = P.linear(xd, wd, bd)
    y.backward(gy.cuda())
    assert rel(y, TF.linear(x, w, b)) < 1e-5
    assert rel(xd.grad.float(), xr.grad) < 5e-3
    assert rel(wd.grad, wr.grad) < 1e-4
    assert rel(bd.grad, br.grad) < 1e-4


GN_CASES = [(2, 32, 32, 128, True, False), (3, 16, 16, 256, True, False), (2, 28, 28, 32, True, False),
            (2, 7, 7, 96, False, False), (5, 4, 4, 512, True, True), (2, 14, 14, 64, True, True),
            (2, 8, 8, 384, False, False)]


@pytest.mark.parametrize("B,H,W,C,silu,ss", GN_CASES)
@pytest.mark.parametrize("xdtype", [bf16, f32])
def test_gn_silu_fwd_bwd(P, B, H, W, C, silu, ss, xdtype):
    x = (rnd(B, C, H, W, seed=1) * 2 + 0.5).to(xdtype).float()
    g, be = 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    sc, sh = (0.3 * rnd(B, C, seed=4), 0.3 * rnd(B, C, seed=5)) if ss else (None, None)
    gy = rnd(B, C, H, W, seed=6).to(bf16).float()
    leaves = [t.clone().requires_grad_(True) for t in (x, g, be)] + ([t.clone().requires_grad_(True) for t in (sc, sh)] if ss else [])
    z = TF.group_norm(leaves[0], 32, leaves[1], leaves[2], eps=1e-5)
    if ss:
        z = z * (1 + leaves[3][:, :, None, None]) + leaves[4][:, :, None, None]
    yr = z * torch.sigmoid(z) if silu else z
    yr.backward(gy)
    xd = x.permute(0, 2, 3, 1).contiguous().to("cuda", xdtype).requires_grad_(True)
    gd, bd = g.cuda().requires_grad_(True), be.cuda().requires_grad_(True)
    scd = sc.cuda().requires_grad_(True) if ss else None
    shd = sh.cuda().requires_grad_(True) if ss else None
    y, mean, rstd = P.gn_silu(xd, gd, bd, scd, shd, 32, 1e-5, silu)
    y.backward(nhwc(gy))
    assert rel(nchw(y), yr.detach()) < 4e-3  # bf16 output rounding
    xg = leaves[0].grad.reshape(B, 32, -1)
    assert rel(mean.cpu(), x.reshape(B, 32, -1).mean(-1)) < 1e-4 + 1e-3
    assert rel(nchw(xd.grad), leaves[0].grad) < (6e-3 if xdtype == bf16 else 1e-4)
    assert rel(gd.grad, leaves[1].grad) < 1e-3
    assert rel(bd.grad, leaves[2].grad) < 1e-3
    if ss:
        assert rel(scd.grad, leaves[3].grad) < 1e-3
        assert rel(shd.grad, leaves[4].grad) < 1e-3
    del xg


ATTN_CASES = [(2, 256, 4, 64), (3, 64, 4, 64), (2, 16, 4, 64), (2, 49, 1, 64), (2, 64, 2, 32), (1, 64, 4, 128),
              (2, 256, 1, 32), (2, 200, 2, 64), (2, 256, 4, 96), (2, 64, 4, 128)]


@pytest.mark.parametrize("B,T,heads,d", ATTN_CASES)
def test_attention_fwd_bwd(P, B, T, heads, d):
    C = heads * d
    qkv = rnd(B, 3 * C, T, seed=1).to(bf16).float()  # reference layout [B, 3C, T]
    gout = rnd(B, C, T, seed=2).to(bf16).float()
    qr = qkv.clone().requires_grad_(True)
    q, k, v = torch.split(qr.reshape(B * heads, 3 * d, T), d, dim=1)  # src/modules/unet.py:249-256
    s = 1 / math.sqrt(math.sqrt(d))
    wgt = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), dim=-1)
    ar = torch.einsum("bts,bcs->bct", wgt, v).reshape(B, C, T)
    ar.backward(gout)
    qd = qkv.permute(0, 2, 1).contiguous().to("cuda", bf16).requires_grad_(True)  # [B, T, 3C]
    out, lse = P.attention(qd, heads)
    out.backward(gout.permute(0, 2, 1).contiguous().to("cuda", bf16))
    torch.cuda.synchronize()
    # P is rounded to bf16 before the PV product and the output is bf16
    assert rel(out.float().permute(0, 2, 1), ar.detach()) < 6e-3
    assert rel(qd.grad.float().permute(0, 2, 1), qr.grad) < 1.5e-2


def test_stem_head_convs(P):
    for Cin, Cout, H in [(3, 128, 32), (1, 32, 28)]:
        x = rnd(2, Cin, H, H, seed=1)
        w, b = rnd(Cout, Cin, 3, 3, seed=2) * 0.2, rnd(Cout, seed=3)
        gy = rnd(2, Cout, H, H, seed=4).to(bf16).float()
        wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yr = TF.conv2d(x, wr, br, padding=1)
        yr.backward(gy)
        wd, bd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        y, _ = P.stem_conv(x.cuda(), wd, bd)
        y.backward(nhwc(gy))
        # stem: the fp32 model input and weights are rounded to bf16 for the tensor-core GEMM
        assert rel(nchw(y), yr.detach()) < 6e-3
        assert rel(wd.grad, wr.grad) < 5e-3 and rel(bd.grad, br.grad) < 1e-4
    for Cin, Cout, H in [(128, 3, 32), (32, 1, 28), (128, 6, 16), (32, 2, 28)]:
        x = rnd(2, Cin, H, H, seed=1).to(bf16).float()
        w, b = rnd(Cout, Cin, 3, 3, seed=2) * 0.05, rnd(Cout, seed=3)
        gy = rnd(2, Cout, H, H, seed=4)
        xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
        yr = TF.conv2d(xr, wr, br, padding=1)
        yr.backward(gy)
        xd = nhwc(x).requires_grad_(True)
        wd, bd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        y = P.head_conv(xd, wd, bd)
        y.backward(gy.cuda())
        # head: bf16 weights/inputs on the tensor cores, fp32 output; dy is rounded to bf16 for the gradient GEMMs
        assert rel(y, yr.detach()) < 4e-3
        assert rel(nchw(xd.grad), xr.grad) < 8e-3
        assert rel(wd.grad, wr.grad) < 5e-3 and rel(bd.grad, br.grad) < 5e-3


def test_layout_and_small_ops(P):
    from probabilisticdeepdiffusionmodels_b200 import functional as F
    x = rnd(3, 5, 6, 7, seed=1)
    assert torch.equal(F.nchw_to_nhwc(x.cuda(), f32).cpu(), x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(F.nhwc_to_nchw(F.nchw_to_nhwc(x.cuda(), f32)).cpu(), x)
    a, b = rnd(2, 4, 4, 64, seed=2).to(bf16).cuda(), rnd(2, 4, 4, 32, seed=3).to(bf16).cuda()
    c = P.concat_channels(a, b)
    assert torch.equal(c, torch.cat([a, b], -1))
    a2, b2 = P.split_channels(c, 64)
    assert torch.equal(a2, a) and torch.equal(b2, b)
    u = F.upsample2x(a)
    assert torch.equal(u, a.repeat_interleave(2, 1).repeat_interleave(2, 2))
    ub = F.upsample2x_bwd(u)
    assert rel(ub.float(), 4 * a.float()) < 4e-3
    ps = F.phase_split(a)
    assert torch.equal(ps[0:2], a[:, 0::2, 0::2]) and torch.equal(ps[6:8], a[:, 1::2, 1::2])
    assert torch.equal(ps[2:4], a[:, 0::2, 1::2]) and torch.equal(ps[4:6], a[:, 1::2, 0::2])
    assert torch.equal(P.add(a, a), (a.float() * 2).to(bf16))
    e = rnd(4, 512, seed=5).cuda().requires_grad_(True)
    s = P.silu_vec(e)
    s.backward(torch.ones_like(s))
    er = e.detach().cpu().clone().requires_grad_(True)
    sr = er * torch.sigmoid(er)
    sr.sum().backward()
    assert rel(s.float(), sr.detach()) < 3e-3 and rel(e.grad, er.grad) < 1e-5
    m = rnd(1000, 96, seed=6).to(bf16).cuda()
    assert rel(F.colsum(m, 96), m.float().sum(0)) < 1e-5
    m3 = rnd(5, 49, 64, seed=7).to(bf16).cuda()
    assert rel(F.colsum_per_sample(m3), m3.float().sum(1)) < 1e-5


def test_timestep_embedding_matches_oracle(P, golden):
    g = golden["kats"]
    t = torch.from_numpy(g["temb_t"]).cuda()
    from probabilisticdeepdiffusionmodels_b200 import functional as F
    for dim, key in [(128, "temb_128"), (32, "temb_32")]:
        out = F.timestep_embedding(t, dim, 10000, f32).cpu().numpy()
        # the frequencies go through expf (<= 2 ulp): a 2e-7 relative change of f moves the phase t*f (t up to
        # 1000 rad) by 2e-4, which bounds the sin/cos difference; small t agree to ~1e-7
        np.testing.assert_allclose(out, g[key], rtol=0, atol=2.5e-4)
        np.testing.assert_allclose(out[:3], g[key][:3], rtol=0, atol=5e-6)
    out = F.timestep_embedding(t.float(), 33, 10000, f32).cpu().numpy()
    np.testing.assert_allclose(out, g["temb_33"], rtol=0, atol=2.5e-4)
    assert P.timestep_embedding(t, 128, 10000.0).dtype == bf16


def test_fused_adam_multi_matches_torch_adam_and_ema():
    """pddm_adam_ema_multi over a ragged parameter list (odd sizes -> scalar tail, misaligned views) against
    torch.optim.Adam + the reference's EMA recurrence, eagerly and under CUDA-graph capture."""
    from probabilisticdeepdiffusionmodels_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(128, 128, 3, 3), (257,), (3, 5, 7), (1,), (20000,), (64, 33)]
    base = [torch.randn(s, device="cuda") for s in shapes]
    ours = [b.clone().requires_grad_(True) for b in base]
    ref = [b.clone().requires_grad_(True) for b in base]
    ema = [b.clone() for b in base]
    ema_ref = [b.clone() for b in base]
    opt = FusedAdam(ours, lr=3e-3, weight_decay=0.01, ema_params=ema, ema_decay=0.9)
    opt_ref = torch.optim.Adam(ref, lr=3e-3, weight_decay=0.01)
    for it in range(4):
        gs = [torch.randn(s, device="cuda") * (it + 1) for s in shapes]
        for p, q, g in zip(ours, ref, gs):
            p.grad, q.grad = g.clone(), g.clone()
        opt.step()
        opt_ref.step()
        for e, q in zip(ema_ref, ref):
            e.mul_(0.9).add_(q.detach(), alpha=0.1)
    for p, q, e, er in zip(ours, ref, ema, ema_ref):
        torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(e, er, rtol=1e-5, atol=1e-6)
    # captured: gradients are static buffers rewritten between replays
    static_g = [torch.zeros(s, device="cuda") for s in shapes]
    for p, g in zip(ours, static_g):
        p.grad = g
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step()  # warm-up (step 5, zero gradient)
    torch.cuda.current_stream().wait_stream(side)
    for q in ref:
        q.grad = torch.zeros_like(q)
    opt_ref.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    opt.flush_tables()
    for it in range(3):
        for g, q in zip(static_g, ref):
            g.normal_()
            q.grad = g.clone()
        graph.replay()
        opt_ref.step()
    for p, q in zip(ours, ref):
        torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-6)


def test_colsum_is_deterministic_and_handles_big_and_wide_shapes():
    """Two-stage column sums (no atomics): bit-identical across runs; multi-block, per-sample and >4096-column panels."""
    from probabilisticdeepdiffusionmodels_b200 import functional as F
    big = rnd(128 * 1024, 128, seed=11).to(bf16).cuda()  # one 32x32 C=128 gradient tensor at B=128
    a, b = F.colsum(big, 128), F.colsum(big, 128)
    assert torch.equal(a, b)
    assert rel(a, big.double().sum(0)) < 1e-5
    ps = big.view(128, 1024, 128)
    pa, pb = F.colsum_per_sample(ps), F.colsum_per_sample(ps)
    assert torch.equal(pa, pb)
    assert rel(pa, ps.double().sum(1)) < 1e-5
    wide = rnd(64, 5120, seed=12).to(bf16).cuda()  # batched timestep-embedding projection width
    assert rel(F.colsum(wide, 5120), wide.double().sum(0)) < 1e-5
    odd = rnd(3, 7, 40, seed=13).to(bf16).cuda()  # ragged rows per block
    assert rel(F.colsum_per_sample(odd), odd.double().sum(1)) < 1e-5


def test_weight_arena_multi_pack_equals_single_packs():
    """ops.WeightArena (one tiled multi-tensor launch) reproduces pddm_pack_conv_weight bit for bit: 3x3 / 1x1 /
    linear, both pack orders, ragged channel counts (tile edges) and zero padding."""
    from probabilisticdeepdiffusionmodels_b200 import functional as F, ops
    cases = [((128, 128, 3, 3), None, None), ((96, 40, 3, 3), None, None), ((256, 384, 1, 1), None, None),
             ((8, 128, 3, 3), 32, None), ((6, 64, 3, 3), 8, 64), ((160, 96), None, None), ((33 * 8, 24, 3, 3), None, None)]
    ws = [rnd(*shape, seed=40 + i).cuda() for i, (shape, _, _) in enumerate(cases)]
    want = {}
    for i, (w, (_, cop, cip)) in enumerate(zip(ws, cases)):
        w3 = w if w.dim() > 2 else w.view(w.shape[0], w.shape[1], 1)
        for mode in (0, 1):
            want[(i, mode)] = F.pack_weight(w3, mode, cop, cip).clone()
    arena = ops.WeightArena()
    with arena.recording():
        for i, (w, (_, cop, cip)) in enumerate(zip(ws, cases)):
            w3 = w if w.dim() > 2 else w.view(w.shape[0], w.shape[1], 1)
            for mode in (0, 1):
                F.pack_weight(w3, mode, cop, cip)
    arena.finalize(ws[0].device)
    arena.repack()
    with arena.active():
        for i, (w, (_, cop, cip)) in enumerate(zip(ws, cases)):
            w3 = w if w.dim() > 2 else w.view(w.shape[0], w.shape[1], 1)
            for mode in (0, 1):
                got = F.pack_weight(w3, mode, cop, cip)
                assert got.shape == want[(i, mode)].shape
                assert torch.equal(got, want[(i, mode)]), (i, mode)
