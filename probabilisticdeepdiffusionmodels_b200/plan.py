"""``UNetPlan``: the forward AND backward pass of a ``UNetModel`` as one hand-scheduled sequence of kernel launches.

``UNetModel.forward`` normally runs op by op through ``torch.ops.pddm.*`` and lets autograd derive the backward pass.
That is general, but autograd decides where gradients meet: every tensor with two consumers (a ResBlock input feeds
the norm and the residual add, src/modules/unet.py:188-201; a skip tensor feeds the next block and a later
``th.cat``, :487-493) costs an extra elementwise add, every ``th.cat`` a copy forward and a split backward, every
bias a column-sum launch, every GroupNorm a small batch-fold launch -- 1.8 ms of a 15.7 ms step on the CIFAR UNet
(profiles/r1_graph_nodes_train_step_v4.summary.txt).  The plan knows the network's structure instead:

* the gradient of a residual / skip branch rides into the GroupNorm backward kernel as ``gres`` and leaves as part
  of ``dx``; the fan-in of a skip tensor is a bulk-tensor reduce-add into the buffer the decoder already wrote;
* ``th.cat([h, skip])`` is never materialised: GroupNorm and the 1x1 skip conv read two sources, the GroupNorm
  backward writes the two halves of the gradient to two tensors;
* per-sample column sums of every gradient (bias, timestep-embedding, GroupNorm affine) land in ONE
  ``[B, columns]`` matrix that a single launch folds over the batch into the small parameters' gradients;
* gradients are born inside one flat fp32 arena (``p.grad`` are views): the data-parallel all-reduce runs on it
  directly, no gather / scatter copies;
* every bf16 GEMM operand lives in one arena refreshed by one launch (``repack``), also for sampling -- a captured
  reverse-step graph keeps reading valid, current packs after the weights change.

Scope: the configurations the reference ships (2-D, ``use_scale_shift_norm=False``, no class conditioning, no
activation checkpointing, dropout 0 or eval mode, channels % 32 == 0).  Anything else keeps the op-by-op path.
"""
import ctypes as C

import torch

from . import _lib as L
from . import functional as F

bf16, f32 = torch.bfloat16, torch.float32
T3, T1 = F.taps_3x3(), F.taps_1x1()


def supported(model):
    from .unet import AttentionBlock, Downsample, ResBlock, Upsample
    if model.num_classes is not None or model.use_checkpoint or model.in_channels > 4 or model.out_channels > 8:
        return False
    for m in model.modules():
        if isinstance(m, ResBlock):
            if m.use_scale_shift_norm or m.use_conv or m.channels % 32 or m.out_channels % 32:
                return False
        elif isinstance(m, (Upsample, Downsample)):
            if not m.use_conv or m.channels % 32:
                return False
        elif isinstance(m, AttentionBlock):
            if m.channels % 32 or (m.channels // m.num_heads) not in (32, 64, 96, 128):
                return False
    return True


class _Gemm:
    """One conv / linear weight: packs (views of the pack arena), gradient slots (views of the gradient arena)."""

    def __init__(self, weight, bias, ntaps, cout_pad=None, cin_pad=None, need_dgrad=True, pack_view=None,
                 need_fwd=True):
        self.w, self.b = weight, bias
        self.need_fwd = need_fwd
        self.cout, self.cin, self.ntaps = weight.shape[0], weight.shape[1], ntaps
        self.cop, self.cip = cout_pad or self.cout, cin_pad or self.cin
        self.need_dgrad = need_dgrad
        self.pack0 = self.pack1 = None
        self.gw = None  # fp32 gradient view, parameter shape
        self.pack_view = pack_view  # (Cout, Cin, ntaps) source view for thin convs (stem)


class Sink:
    """Where a block's input gradient goes: a fresh tensor, or an existing one to add into (skip fan-in)."""

    def __init__(self, cs, buf=None, accumulate=False):
        self.cs, self.buf, self.accumulate = cs, buf, accumulate


class UNetPlan:
    def __init__(self, model, batch, height, width, device):
        from .unet import AttentionBlock, Downsample, ResBlock, Upsample
        if not supported(model):
            raise ValueError("UNetPlan: this model configuration takes the op-by-op path")
        self.model, self.B, self.H0, self.W0 = model, batch, height, width
        self.dev = torch.device(device)
        L.require_device()
        self.gemms = []
        self.small = []          # (param, ps column start) -> gradients produced by the batch fold
        # small parameters whose gradient a kernel writes directly
        self.direct_small = [model.time_embed[0].bias, model.time_embed[2].bias]
        self.ncols = 0
        self.mc = model.model_channels
        self.side = None
        self.overlap = True
        self._keep = []
        self.wg_bytes = 16       # split-K workspace of the conv weight gradients (side stream)
        self.wg_bytes_main = 16  # ... of the few weight gradients issued on the main stream (stem, head, embedding)
        self.comm = None         # parallel.OverlappedArenaAllReduce: told when a range of the gradient arena is final
        self._gi_down = self._gi_mid = None  # index in self.gemms of the first weight of the first Downsample / middle block

        # ---- timestep embedding MLP and the batched emb_layers projection
        self.te0 = self._gemm(model.time_embed[0].weight, model.time_embed[0].bias, 1, need_dgrad=False)
        self.te2 = self._gemm(model.time_embed[2].weight, model.time_embed[2].bias, 1)
        self.res_blocks = [m for m in model.modules() if isinstance(m, ResBlock)]
        self.emb_sizes = [m.emb_layers[1].weight.shape[0] for m in self.res_blocks]
        self.sumC = sum(self.emb_sizes)
        self.emb_off, off = {}, 0
        for m, n in zip(self.res_blocks, self.emb_sizes):
            self.emb_off[id(m)] = off
            off += n
        self.emb_col0 = self._cols(self.sumC)  # d(emb_out) per sample = sum_hw of conv1's output gradient
        self.E = model.time_embed[2].weight.shape[0]
        # all emb_layers projections as ONE [sumC, E] GEMM operand (its packs are assembled from the 30 parameters)
        self.emb_all = _Gemm(torch.empty(self.sumC, self.E, 1, device="meta"), None, 1)
        for cout in (self.sumC, self.E):  # the embedding linears run as [1, 1, B, K] one-tap GEMMs
            self.wg_bytes_main = max(self.wg_bytes_main, F.wgrad_workspace_bytes(1, 1, batch, self.E, cout, 1, 1))

        # ---- walk the network once: shapes, nodes, skip wiring
        H, W = height, width
        stem = model.input_blocks[0][0]
        kp = (model.in_channels * 9 + 31) // 32 * 32
        self.stem = _StemNode(self, stem, H, W, kp)
        enc_nodes, chans = [[self.stem]], [stem.weight.shape[0]]
        ch = chans[0]
        for seq in list(model.input_blocks)[1:]:
            nodes = []
            for layer in seq:
                if isinstance(layer, ResBlock):
                    nodes.append(_ResNode(self, layer, ch, 0, H, W))
                    ch = layer.out_channels
                elif isinstance(layer, AttentionBlock):
                    nodes.append(_AttnNode(self, layer, H, W))
                elif isinstance(layer, Downsample):
                    if self._gi_down is None:
                        self._gi_down, self._enc_down = len(self.gemms), len(enc_nodes)
                    nodes.append(_DownNode(self, layer, H, W))
                    H, W = H // 2, W // 2
                else:
                    raise ValueError(f"UNetPlan: unexpected layer {type(layer)}")
            enc_nodes.append(nodes)
            chans.append(ch)
        self.enc = enc_nodes
        self.mid = []
        self._gi_mid = len(self.gemms)
        for layer in model.middle_block:
            self.mid.append(_ResNode(self, layer, ch, 0, H, W) if isinstance(layer, ResBlock)
                            else _AttnNode(self, layer, H, W))
        self.dec = []
        skips = list(chans)
        for seq in model.output_blocks:
            cs_ = skips.pop()
            nodes = []
            for layer in seq:
                if isinstance(layer, ResBlock):
                    nodes.append(_ResNode(self, layer, ch, cs_, H, W))
                    ch = layer.out_channels
                elif isinstance(layer, AttentionBlock):
                    nodes.append(_AttnNode(self, layer, H, W))
                elif isinstance(layer, Upsample):
                    nodes.append(_UpNode(self, layer, H, W))
                    H, W = 2 * H, 2 * W
                else:
                    raise ValueError(f"UNetPlan: unexpected layer {type(layer)}")
            self.dec.append(nodes)
        self.head = _HeadNode(self, model.out[0], model.out[2], H, W)
        # which columns of `ps` receive sum_hw of the gradient of each node's INPUT = its producer's output columns
        prev = None
        for nodes in self.enc[1:]:
            for j, nd in enumerate(nodes):
                nd.in_cols = (prev.out_cs, prev.cout) if j else None  # first node: skip fan-in, wired in backward()
                prev = nd
        prev = self.enc[-1][-1]
        for nd in self.mid:
            nd.in_cols = (prev.out_cs, prev.cout)
            prev = nd
        for nodes in self.dec:
            for nd in nodes:
                nd.in_cols = (prev.out_cs, prev.cout)
                prev = nd
        self.head.in_cols = (prev.out_cs, prev.cout)

        self._build_arenas()

    # ------------------------------------------------------------------ construction helpers
    def _cols(self, n):
        c0 = self.ncols
        self.ncols += (n + 7) // 8 * 8
        return c0

    def _gemm(self, weight, bias, ntaps, **kw):
        g = _Gemm(weight, bias, ntaps, **kw)
        self.gemms.append(g)
        return g

    def _small(self, param, col0):
        self.small.append((param, col0))

    def _need_wgrad_ws(self, B, H, W, cin, cout, ntaps, x_NB=None, main=False):
        n = F.wgrad_workspace_bytes(B, H, W, cin, cout, ntaps, x_NB)
        if main:
            self.wg_bytes_main = max(self.wg_bytes_main, n)
        else:
            self.wg_bytes = max(self.wg_bytes, n)

    def _build_arenas(self):
        dev = self.dev
        # ---------------- gradient arena, in the order the backward pass FINISHES its parts, last first:
        #   [folded small | direct small | emb_layers weights | time_embed, stem, full-resolution encoder |
        #    encoder from the first Downsample on | middle block, decoder, head]
        # so that a data-parallel run can all-reduce the tail ranges while the backward pass is still producing the
        # head of the arena (``bucket_bounds``; parallel.OverlappedArenaAllReduce)
        emb_w = [m.emb_layers[1].weight for m in self.res_blocks]
        emb_ids = {id(w) for w in emb_w}
        big = emb_w + [g.w for g in self.gemms if id(g.w) not in emb_ids]
        seen, order = set(), []
        for w in big:
            if id(w) not in seen:
                seen.add(id(w))
                order.append(w)
        off, self.goff = 0, {}
        self.small_off = off
        src_of = []
        for prm, col0 in self.small:
            self.goff[id(prm)] = off
            src_of += list(range(col0, col0 + prm.numel()))
            off += prm.numel()
        self.n_fold = len(src_of)
        off = (off + 3) // 4 * 4
        for prm in self.direct_small:
            self.goff[id(prm)] = off
            off += (prm.numel() + 3) // 4 * 4
        off = (off + 1023) // 1024 * 1024
        for w in order:
            self.goff[id(w)] = off
            off += (w.numel() + 3) // 4 * 4
        covered = set(self.goff)
        missing = [n for n, p_ in self.model.named_parameters() if id(p_) not in covered]
        if missing:
            raise ValueError(f"UNetPlan: parameters without a gradient slot: {missing[:4]}")
        self.grad_arena = torch.zeros(off, dtype=f32, device=dev)

        def first_weight_off(gi):
            ws = [g.w for g in self.gemms[gi:] if id(g.w) not in emb_ids]
            return self.goff[id(ws[0])] if ws else off

        b_mid = first_weight_off(self._gi_mid)
        b_down = first_weight_off(self._gi_down) if self._gi_down is not None else b_mid
        # [0, b_down): final at the very end; [b_down, b_mid): once the first Downsample's backward has been issued;
        # [b_mid, end): once the middle block's backward has been issued
        self.bucket_bounds = (b_down, b_mid, off)
        self.src_of = torch.tensor(src_of, dtype=torch.int32, device=dev)
        self.params = list(self.model.parameters())
        for g in self.gemms:
            g.gw = self.gview(g.w)
        self.emb_gw = self.grad_arena[self.goff[id(emb_w[0])]: self.goff[id(emb_w[0])] + self.sumC * self.E] \
            .view(self.sumC, self.E, 1)
        self.ps = torch.zeros((self.B, self.ncols), dtype=f32, device=dev)
        self.wg_ws = torch.empty(self.wg_bytes, dtype=torch.uint8, device=dev)
        self.wg_ws_main = torch.empty(self.wg_bytes_main, dtype=torch.uint8, device=dev)

        # ---------------- pack arena (bf16 GEMM operands) + descriptor table of the one-launch re-pack
        descs, total = [], 0

        def add(g, mode, both=False):
            nonlocal total
            n = g.cop * g.cip * g.ntaps
            n8 = (n + 7) // 8 * 8
            descs.append((g, mode, total, total + n8 if both else None))
            total += n8 * (2 if both else 1)

        for g in self.gemms:
            if g.need_fwd and g.need_dgrad:
                add(g, 0, both=True)  # one source tile feeds the forward and the data-gradient operand
            elif g.need_fwd:
                add(g, 0)
            elif g.need_dgrad:
                add(g, 1)
        emb0, emb1 = total, total + self.sumC * self.E
        total += 2 * self.sumC * self.E
        self.pack_arena = torch.zeros(total, dtype=bf16, device=dev)
        tab = (L.PackDesc * (len(descs) + 2 * len(emb_w)))()
        blocks = []

        def fill(i, src, dst_ptr, cout, cin, nt, mode, cop, cip, ld_dst=0, dst2=None):
            d = tab[i]
            d.src, d.dst, d.dst2 = src.data_ptr(), dst_ptr, dst2
            d.Cout, d.Cin, d.ntaps, d.mode, d.Cout_pad, d.Cin_pad, d.ld_dst = cout, cin, nt, mode, cop, cip, ld_dst
            tiles = -(-cop // L.PACK_TILE) * -(-cip // L.PACK_TILE)
            blocks.extend((i, t_) for t_ in range(tiles))

        i = 0
        base = self.pack_arena.data_ptr()
        for g, mode, o, o2 in descs:
            n = g.cop * g.cip * g.ntaps
            view = self.pack_arena[o: o + n].view((g.cop, g.ntaps, g.cip) if mode == 0 else (g.cip, g.ntaps, g.cop))
            if mode == 0:
                g.pack0 = view
            else:
                g.pack1 = view
            if o2 is not None:
                g.pack1 = self.pack_arena[o2: o2 + n].view(g.cip, g.ntaps, g.cop)
            src = g.pack_view if g.pack_view is not None else g.w
            cout, cin, nt = (src.shape[0], src.shape[1], src.numel() // (src.shape[0] * src.shape[1]))
            fill(i, src, base + 2 * o, cout, cin, nt, mode, g.cop, g.cip, dst2=None if o2 is None else base + 2 * o2)
            i += 1
        # emb_layers: 30 [Cout_i, E] matrices side by side = one [sumC, E] operand (forward) and its transpose
        self.emb_all.pack0 = self.pack_arena[emb0: emb0 + self.sumC * self.E].view(self.sumC, 1, self.E)
        self.emb_all.pack1 = self.pack_arena[emb1: emb1 + self.sumC * self.E].view(self.E, 1, self.sumC)
        row = 0
        for w in emb_w:
            n = w.shape[0]
            fill(i, w, base + 2 * (emb0 + row * self.E), n, self.E, 1, 0, n, self.E)
            fill(i + 1, w, base + 2 * (emb1 + row), n, self.E, 1, 1, n, self.E, ld_dst=self.sumC)
            i += 2
            row += n
        self.max_ntaps = max(g.ntaps for g in self.gemms)
        self._descs = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(dev)
        self._blocks = torch.tensor(blocks, dtype=torch.int32).to(dev)
        self.nblocks = len(blocks)
        self.emb_bias = torch.zeros(self.sumC, dtype=f32, device=dev)
        self._emb_biases = [m.emb_layers[1].bias for m in self.res_blocks]
        self._pack_stamp = None
        self._ptr_sig = (self.params[0].data_ptr(), self.params[-1].data_ptr())

    def gview(self, prm):
        o = self.goff[id(prm)]
        return self.grad_arena[o: o + prm.numel()].view(prm.shape)

    # ------------------------------------------------------------------ weights -> bf16 operands
    def repack(self):
        """ONE launch refreshes every bf16 GEMM operand from the fp32 parameters (+ one cat of the emb biases)."""
        L.call("pddm_pack_weights_multi", L.ptr(self._descs), L.ptr(self._blocks), self.nblocks, self.max_ntaps,
               L.stream())
        torch.cat([b.detach() for b in self._emb_biases], out=self.emb_bias)

    def weights_stamp(self):
        return tuple(p_._version for p_ in self.params) + (_EPOCH[0],)

    def matches(self, model):
        """Still bound to this model's storage?  (``model.to(...)`` / ``.cuda()`` replace the parameters' memory.)"""
        ps = self.params
        return ps[0].data_ptr() == self._ptr_sig[0] and ps[-1].data_ptr() == self._ptr_sig[1] and \
            ps[0].device == self.dev

    def refresh_packs(self):
        """Re-pack if any parameter changed since the last pack (eager / sampling use; a captured training step calls
        ``repack`` itself every step).  Kernels that update parameters through raw pointers bump ``_EPOCH``."""
        st = self.weights_stamp()
        if st != self._pack_stamp:
            self.repack()
            self._pack_stamp = st

    # ------------------------------------------------------------------ launch helpers
    def lin(self, x, g, bias=None, out=None):
        M, K = x.shape
        y = F.tap_gemm(x.view(1, 1, M, K), g.pack0, T1, 1, 1, M, bias=g.b if bias is None else bias, out_dtype=f32,
                       out=None if out is None else out.view(1, 1, M, g.cout))
        return y.view(M, g.cout)

    def wgrad(self, x, dy, g, taps, B, H, W, x2=None):
        """Weight gradient straight into the arena; on the side stream while the backward chain continues."""
        cur = torch.cuda.current_stream(self.dev)
        st = None
        if self.overlap:
            if self.side is None:
                self.side = torch.cuda.Stream(device=self.dev)
            self.side.wait_stream(cur)
            st = self.side
            self._keep.extend((x, dy, x2))
        if x2 is None:
            F.tap_wgrad(x, dy, taps, B, H, W, g.cin, g.cout, None, out=g.gw, ws=self.wg_ws, launch_stream=st)
        else:
            ca = x.shape[-1]
            F.tap_wgrad(x, dy, taps, B, H, W, ca, g.cout, None, out=g.gw, ws=self.wg_ws, launch_stream=st,
                        dw_ldc=g.cin, dw_c0=0)
            F.tap_wgrad(x2, dy, taps, B, H, W, g.cin - ca, g.cout, None, out=g.gw, ws=self.wg_ws, launch_stream=st,
                        dw_ldc=g.cin, dw_c0=ca)

    def cs_view(self, col0, n):
        return self.ps[:, col0: col0 + n]

    def sink(self, node):
        """A fresh destination for the gradient of ``node``'s input; its column sums go to the producer's columns."""
        return Sink(self.cs_view(*node.in_cols))

    def gn_bwd(self, x, x2, dy, norm, mean, rstd, silu, gres, sink_a, sink_b, part_cols, xcat=None):
        """GroupNorm backward delivering dx (+gres) into the sinks, per-sample partials into ``ps``.
        Returns (dx_a, dx_b)."""
        B = x.shape[0]
        C_a = x.shape[-1]
        C_ = C_a + (x2.shape[-1] if x2 is not None else 0)
        HW = x.numel() // (B * C_a)
        G = norm.num_groups
        pg = self.ps[:, part_cols: part_cols + C_]
        pb = self.ps[:, part_cols + C_: part_cols + 2 * C_]
        ntens = 3 if gres is not None else 2
        if F.gn_pipe_slots(B, HW, C_, G, C_a if x2 is not None else 0, ntens) >= 2:
            dxa = sink_a.buf if sink_a.buf is not None else torch.empty(x.shape, dtype=bf16, device=x.device)
            dxb = None
            if x2 is not None:
                dxb = sink_b.buf if sink_b.buf is not None else torch.empty(x2.shape, dtype=bf16, device=x.device)
            F.gn_silu_bwd(x, dy, norm.weight, norm.bias, mean, rstd, G, silu, x2=x2, gres=gres, dx=dxa, dx2=dxb,
                          dx_accumulate=sink_a.accumulate, dx2_accumulate=bool(sink_b and sink_b.accumulate),
                          part_dgamma=pg, part_dbeta=pb, colsum=sink_a.cs, colsum2=sink_b.cs if sink_b else None,
                          colsum_accumulate=sink_a.accumulate, colsum2_accumulate=bool(sink_b and sink_b.accumulate))
            return dxa, dxb
        # ---- shapes the persistent kernel does not take: compose from simpler launches
        xin = xcat if xcat is not None else (x if x2 is None else F.concat_channels(x.contiguous(), x2.contiguous()))
        if F.gn_pipe_slots(B, HW, C_, G, 0, 2) >= 2:
            dx, _, _, _, _, _ = F.gn_silu_bwd(xin, dy, norm.weight, norm.bias, mean, rstd, G, silu, part_dgamma=pg,
                                              part_dbeta=pb)
        else:
            dx, dg, db, _, _, _ = F.gn_silu_bwd(xin, dy, norm.weight, norm.bias, mean, rstd, G, silu)
            self.ps[:, part_cols: part_cols + 2 * C_].zero_()
            self.ps[0, part_cols: part_cols + C_].copy_(dg)
            self.ps[0, part_cols + C_: part_cols + 2 * C_].copy_(db)
        if gres is not None:
            dx = F.add_bf16(dx, gres.contiguous())
        outs = []
        for sink, lo, hi in ((sink_a, 0, C_a),) + (((sink_b, C_a, C_),) if x2 is not None else ()):
            part = dx if (lo == 0 and hi == C_) else None
            if part is None:
                part = torch.empty(dx.shape[:-1] + (hi - lo,), dtype=bf16, device=dx.device)
                F.copy_channels(dx, lo, part, 0, hi - lo)
            if sink.buf is not None:
                if sink.accumulate:
                    part = F.add_bf16(sink.buf, part)
                sink.buf.copy_(part)
                part = sink.buf
            if sink.cs is not None:
                F.colsum_rows(part, sink.cs)  # total after the fan-in
            outs.append(part)
        return outs[0], (outs[1] if x2 is not None else None)

    # ------------------------------------------------------------------ forward
    def forward(self, x, t, save):
        """x: NCHW fp32 [B, C, H, W]; t: [B] int64 / float32 -> NCHW fp32 model output."""
        if x.shape[0] != self.B or x.shape[2] != self.H0 or x.shape[3] != self.W0:
            raise ValueError("UNetPlan.forward: input shape differs from the planned one")
        S = {} if save else None
        te = F.timestep_embedding(t.contiguous(), self.mc)
        h1 = self.lin(te, self.te0)
        a1 = F.silu_vec(h1)
        emb = self.lin(a1, self.te2)
        act = F.silu_vec(emb)
        emb_all = self.lin(act, self.emb_all, bias=self.emb_bias)
        if save:
            S["emb"] = (te, h1, a1, emb, act)
        self.emb_all_out = emb_all
        h = self.stem.fwd(x, S)
        hs = [h]
        for nodes in self.enc[1:]:
            for nd in nodes:
                h = nd.fwd(h, None, S)
            hs.append(h)
        for nd in self.mid:
            h = nd.fwd(h, None, S)
        for nodes in self.dec:
            skip = hs.pop()
            h = nodes[0].fwd(h, skip, S)
            for nd in nodes[1:]:
                h = nd.fwd(h, None, S)
        out = self.head.fwd(h, S)
        self.emb_all_out = None
        return (out, S) if save else out

    # ------------------------------------------------------------------ backward
    def backward(self, S, dout, overlap=True):
        """S: what ``forward(save=True)`` returned; dout: NCHW fp32 gradient of the model output.
        Fills the gradient arena (see ``assign_grads``)."""
        self.overlap = overlap
        g = self.head.bwd(dout, S)
        skip_g = {}
        n_enc = len(self.enc)
        for k in reversed(range(len(self.dec))):
            nodes = self.dec[k]
            for nd in reversed(nodes[1:]):
                g = nd.bwd(g, S, self.sink(nd))
            idx = n_enc - 1 - k
            enc_last = self.enc[idx][-1]
            g, gs = nodes[0].bwd(g, S, self.sink(nodes[0]), Sink(self.cs_view(enc_last.out_cs, enc_last.cout)))
            skip_g[idx] = gs
        for nd in reversed(self.mid[1:]):
            g = nd.bwd(g, S, self.sink(nd))
        last = n_enc - 1
        self.mid[0].bwd(g, S, Sink(self.cs_view(self.enc[last][-1].out_cs, self.enc[last][-1].cout), skip_g[last], True))
        b_down, b_mid, b_end = self.bucket_bounds
        if self.comm is not None:  # every gradient of the middle block, the decoder and the head has been issued
            self.comm.range_final(self, b_mid, b_end)
        for i in reversed(range(1, n_enc)):
            g = skip_g.pop(i)
            nodes = self.enc[i]
            for nd in reversed(nodes[1:]):
                g = nd.bwd(g, S, self.sink(nd))
            prev = self.enc[i - 1][-1]
            nodes[0].bwd(g, S, Sink(self.cs_view(prev.out_cs, prev.cout), skip_g[i - 1], True))
            if self.comm is not None and self._gi_down is not None and i == self._enc_down and b_down < b_mid:
                self.comm.range_final(self, b_down, b_mid)  # ... and of the encoder below full resolution
        self.stem.bwd(skip_g.pop(0), S)
        self._emb_backward(S)
        F.batch_fold(self.ps, self.src_of, self.n_fold, self.grad_arena[self.small_off:])
        if self.side is not None and overlap:
            torch.cuda.current_stream(self.dev).wait_stream(self.side)
        self._keep.clear()

    def _emb_backward(self, S):
        te, h1, a1, emb, act = S["emb"]
        B = self.B
        d_all = torch.empty((B, self.sumC), dtype=bf16, device=self.dev)
        F.convert_rows(self.cs_view(self.emb_col0, self.sumC), d_all)
        # all 30 emb_layers weight gradients = ONE [sumC, E] GEMM; their rows are contiguous in the arena
        F.tap_wgrad(act.view(1, 1, B, self.E), d_all.view(1, 1, B, self.sumC), T1, 1, 1, B, self.E, self.sumC, None,
                    out=self.emb_gw, ws=self.wg_ws_main)
        d_act = F.tap_gemm(d_all.view(1, 1, B, self.sumC), self.emb_all.pack1, T1, 1, 1, B, out_dtype=f32).view(B, self.E)
        d_emb = F.silu_vec_bwd(emb, d_act)
        d_emb_b = F.convert(d_emb, bf16)
        g2 = self.te2
        F.tap_wgrad(a1.view(1, 1, B, g2.cin), d_emb_b.view(1, 1, B, g2.cout), T1, 1, 1, B, g2.cin, g2.cout, None,
                    out=g2.gw.view(g2.cout, g2.cin, 1), ws=self.wg_ws_main)
        F.colsum(d_emb_b, g2.cout, out=self.gview(g2.b))
        d_a1 = F.tap_gemm(d_emb_b.view(1, 1, B, g2.cout), g2.pack1, T1, 1, 1, B, out_dtype=f32).view(B, g2.cin)
        d_h1 = F.convert(F.silu_vec_bwd(h1, d_a1), bf16)
        g0 = self.te0
        F.tap_wgrad(te.view(1, 1, B, g0.cin), d_h1.view(1, 1, B, g0.cout), T1, 1, 1, B, g0.cin, g0.cout, None,
                    out=g0.gw.view(g0.cout, g0.cin, 1), ws=self.wg_ws_main)
        F.colsum(d_h1, g0.cout, out=self.gview(g0.b))

    def assign_grads(self):
        """Point every ``p.grad`` at its slice of the gradient arena (views: no copy, no kernel)."""
        for prm in self.params:
            prm.grad = self.gview(prm)


_EPOCH = [0]


def bump_weight_epoch():
    """Parameters were changed through raw pointers (fused Adam / EMA kernels): cached packs are stale."""
    _EPOCH[0] += 1


# ================================================================================================ nodes
class _StemNode:
    """First conv (Cin <= 4, src/modules/unet.py:353): im2col to K = 32, then a one-tap tensor-core GEMM."""

    def __init__(self, plan, conv, H, W, kp):
        self.plan, self.H, self.W, self.kp = plan, H, W, kp
        self.cout = conv.weight.shape[0]
        self.g = plan._gemm(conv.weight, conv.bias, 1, cin_pad=kp, need_dgrad=False,
                            pack_view=conv.weight.view(self.cout, -1, 1))
        self.out_cs = plan._cols(self.cout)
        plan._small(conv.bias, self.out_cs)
        plan._need_wgrad_ws(plan.B, H, W, kp, self.cout, 1, main=True)

    def fwd(self, x, S):
        patches = F.im2col3x3(x.float().contiguous())
        if S is not None:
            S[id(self)] = patches
        return F.tap_gemm(patches, self.g.pack0, T1, self.plan.B, self.H, self.W, bias=self.g.b)

    def bwd(self, g, S):
        patches = S[id(self)]
        pl = self.plan
        dwp = F.tap_wgrad(patches, g, T1, pl.B, self.H, self.W, self.kp, self.cout, (self.cout, self.kp, 1),
                          ws=pl.wg_ws_main)
        k = self.g.w.shape[1] * 9
        self.g.gw.view(self.cout, k).copy_(dwp.view(self.cout, self.kp)[:, :k])


class _ResNode:
    """ResBlock (src/modules/unet.py:111-201) with a one- or two-source input."""

    def __init__(self, plan, mod, cin_a, cin_b, H, W):
        self.plan, self.mod, self.H, self.W = plan, mod, H, W
        self.ca, self.cb, self.cin, self.cout = cin_a, cin_b, cin_a + cin_b, mod.out_channels
        assert self.cin == mod.channels
        self.n1, self.n2 = mod.in_layers[0], mod.out_layers[0]
        c1, c2 = mod.in_layers[2], mod.out_layers[3]
        self.c1 = plan._gemm(c1.weight, c1.bias, 9)
        self.c2 = plan._gemm(c2.weight, c2.bias, 9)
        self.sk = None
        if not isinstance(mod.skip_connection, torch.nn.Identity):
            sc = mod.skip_connection
            self.sk = plan._gemm(sc.weight, sc.bias, 1)
        self.emb_o = plan.emb_off[id(mod)]
        lin = mod.emb_layers[1]
        self.out_cs = plan._cols(self.cout)
        self.p1 = plan._cols(2 * self.cin)
        self.p2 = plan._cols(2 * self.cout)
        plan._small(self.n1.weight, self.p1)
        plan._small(self.n1.bias, self.p1 + self.cin)
        plan._small(self.n2.weight, self.p2)
        plan._small(self.n2.bias, self.p2 + self.cout)
        ecol = plan.emb_col0 + self.emb_o
        plan._small(c1.bias, ecol)
        plan._small(lin.bias, ecol)
        plan._small(c2.bias, self.out_cs)
        if self.sk is not None:
            plan._small(mod.skip_connection.bias, self.out_cs)
        B = plan.B
        plan._need_wgrad_ws(B, H, W, self.cin, self.cout, 9)
        plan._need_wgrad_ws(B, H, W, self.cout, self.cout, 9)
        if self.sk is not None:  # the 1x1 skip conv: whole input, or its two sources one after the other
            for c_ in {self.cin, cin_a, cin_b} - {0}:
                plan._need_wgrad_ws(B, H, W, c_, self.cout, 1)
        # can the two-source input stay un-concatenated?  (GroupNorm chunks must not straddle the seam)
        self.two_src = cin_b > 0 and F.gn_pipe_slots(B, H * W, self.cin, self.n1.num_groups, cin_a, 1) >= 2 and \
            cin_a % (64 if self.cin % 64 == 0 else 32) == 0

    def fwd(self, x, skip, S):
        pl, B, H, W = self.plan, self.plan.B, self.H, self.W
        x2 = skip
        if skip is not None and not self.two_src:
            x, x2 = F.concat_channels(x, skip), None
        n1, m1, r1 = F.gn_silu_fwd(x, self.n1.weight, self.n1.bias, self.n1.num_groups, self.n1.eps, True, x2=x2)
        emb_out = pl.emb_all_out[:, self.emb_o: self.emb_o + self.cout]
        c1 = F.tap_gemm(n1, self.c1.pack0, T3, B, H, W, bias=self.c1.b, bcast=emb_out)
        n2, m2, r2 = F.gn_silu_fwd(c1, self.n2.weight, self.n2.bias, self.n2.num_groups, self.n2.eps, True)
        if self.mod.dropout > 0 and self.mod.training:
            raise RuntimeError("UNetPlan does not implement dropout > 0 in training mode")
        if self.sk is None:
            res = x
        else:
            res = F.tap_gemm(x, self.sk.pack0, T1, B, H, W, bias=self.sk.b, x2=x2)
        out = F.tap_gemm(n2, self.c2.pack0, T3, B, H, W, bias=self.c2.b, residual=res)
        if S is not None:
            S[id(self)] = (x, x2, n1, m1, r1, c1, n2, m2, r2)
        return out

    def bwd(self, g, S, sink_a, sink_b=None):
        """g: gradient of the block output (its per-sample column sums are already in ps[:, out_cs]).
        Returns the input gradient (two tensors for a two-source block)."""
        pl, B, H, W = self.plan, self.plan.B, self.H, self.W
        x, x2, n1, m1, r1, c1, n2, m2, r2 = S.pop(id(self))
        pl.wgrad(n2, g, self.c2, T3, B, H, W)
        dn2 = F.tap_gemm(g, self.c2.pack1, T3, B, H, W)
        ecs = pl.cs_view(pl.emb_col0 + self.emb_o, self.cout)
        dc1, _ = pl.gn_bwd(c1, None, dn2, self.n2, m2, r2, True, None, Sink(ecs), None, self.p2)
        pl.wgrad(n1, dc1, self.c1, T3, B, H, W)
        dn1 = F.tap_gemm(dc1, self.c1.pack1, T3, B, H, W)
        if self.sk is None:
            gres = g
        else:
            pl.wgrad(x, g, self.sk, T1, B, H, W, x2=x2)
            gres = F.tap_gemm(g, self.sk.pack1, T1, B, H, W)
        if self.cb and x2 is None:
            # the concat was materialised in the forward pass (its GroupNorm chunks straddle the seam)
            return pl.gn_bwd(x[..., : self.ca], x[..., self.ca:], dn1, self.n1, m1, r1, True, gres, sink_a, sink_b,
                             self.p1, xcat=x)
        dxa, dxb = pl.gn_bwd(x, x2, dn1, self.n1, m1, r1, True, gres, sink_a, sink_b, self.p1)
        return (dxa, dxb) if self.cb else dxa


class _AttnNode:
    """AttentionBlock (src/modules/unet.py:204-234)."""

    def __init__(self, plan, mod, H, W):
        self.plan, self.mod, self.H, self.W = plan, mod, H, W
        self.C = self.cout = mod.channels
        self.norm = mod.norm
        self.qkv = plan._gemm(mod.qkv.weight, mod.qkv.bias, 1)
        self.proj = plan._gemm(mod.proj_out.weight, mod.proj_out.bias, 1)
        self.out_cs = plan._cols(self.C)
        self.pn = plan._cols(2 * self.C)
        self.q_cs = plan._cols(3 * self.C)
        plan._small(self.norm.weight, self.pn)
        plan._small(self.norm.bias, self.pn + self.C)
        plan._small(mod.qkv.bias, self.q_cs)
        plan._small(mod.proj_out.bias, self.out_cs)
        plan._need_wgrad_ws(plan.B, H, W, self.C, 3 * self.C, 1)
        plan._need_wgrad_ws(plan.B, H, W, self.C, self.C, 1)

    def fwd(self, x, _skip, S):
        pl, B, H, W, Cc = self.plan, self.plan.B, self.H, self.W, self.C
        n, m, r = F.gn_silu_fwd(x, self.norm.weight, self.norm.bias, self.norm.num_groups, self.norm.eps, False)
        qkv = F.tap_gemm(n, self.qkv.pack0, T1, B, H, W, bias=self.qkv.b)
        a, lse = F.attn_fwd(qkv.view(B, H * W, 3 * Cc), self.mod.num_heads)
        out = F.tap_gemm(a.view(B, H, W, Cc), self.proj.pack0, T1, B, H, W, bias=self.proj.b, residual=x)
        if S is not None:
            S[id(self)] = (x, n, m, r, qkv, a, lse)
        return out

    def bwd(self, g, S, sink):
        pl, B, H, W, Cc = self.plan, self.plan.B, self.H, self.W, self.C
        x, n, m, r, qkv, a, lse = S.pop(id(self))
        av = a.view(B, H, W, Cc)
        pl.wgrad(av, g, self.proj, T1, B, H, W)
        da = F.tap_gemm(g, self.proj.pack1, T1, B, H, W)
        dqkv = F.attn_bwd(qkv.view(B, H * W, 3 * Cc), a, da.view(B, H * W, Cc), lse, self.mod.num_heads)
        dq4 = dqkv.view(B, H, W, 3 * Cc)
        F.colsum_rows(dq4, pl.cs_view(self.q_cs, 3 * Cc))
        pl.wgrad(n, dq4, self.qkv, T1, B, H, W)
        dn = F.tap_gemm(dq4, self.qkv.pack1, T1, B, H, W)
        dx, _ = pl.gn_bwd(x, None, dn, self.norm, m, r, False, g, sink, None, self.pn)
        return dx


class _DownNode:
    """Downsample = conv3x3 stride 2 (src/modules/unet.py:85-108) as a tap GEMM over the phase-split input."""

    def __init__(self, plan, mod, H, W):
        self.plan, self.H, self.W = plan, H, W
        self.C = self.cout = mod.channels
        self.g = plan._gemm(mod.op.weight, mod.op.bias, 9)
        self.out_cs = plan._cols(self.C)
        plan._small(mod.op.bias, self.out_cs)
        plan._need_wgrad_ws(plan.B, H // 2, W // 2, self.C, self.C, 9, x_NB=4 * plan.B)

    def fwd(self, x, _skip, S):
        B = self.plan.B
        xs = F.phase_split(x)
        if S is not None:
            S[id(self)] = xs
        return F.tap_gemm(xs, self.g.pack0, F.taps_stride2(B), B, self.H // 2, self.W // 2, bias=self.g.b)

    def bwd(self, g, S, sink):
        pl, B, H, W = self.plan, self.plan.B, self.H // 2, self.W // 2
        xs = S.pop(id(self))
        pl.wgrad(xs, g, self.g, F.taps_stride2(B), B, H, W)
        dx = sink.buf if sink.buf is not None else torch.empty((B, 2 * H, 2 * W, self.C), dtype=bf16, device=g.device)
        res = dx if (sink.buf is not None and sink.accumulate) else None  # in place: each element read, then written
        for a in range(2):
            for b in range(2):
                F.tap_gemm(g, self.g.pack1, F.taps_stride2_dgrad(a, b), B, H, W, out=dx, out_hw=(2 * H, 2 * W),
                           out_map=(2, 2, a, b), residual=res)
        if sink.cs is not None:
            F.colsum_rows(dx, sink.cs)
        return dx


class _UpNode:
    """Upsample = nearest x2, conv3x3 (src/modules/unet.py:54-82)."""

    def __init__(self, plan, mod, H, W):
        self.plan, self.H, self.W = plan, H, W
        self.C = self.cout = mod.channels
        self.g = plan._gemm(mod.conv.weight, mod.conv.bias, 9)
        self.out_cs = plan._cols(self.C)
        plan._small(mod.conv.bias, self.out_cs)
        plan._need_wgrad_ws(plan.B, 2 * H, 2 * W, self.C, self.C, 9)

    def fwd(self, x, _skip, S):
        xu = F.upsample2x(x)
        if S is not None:
            S[id(self)] = xu
        return F.tap_gemm(xu, self.g.pack0, T3, self.plan.B, 2 * self.H, 2 * self.W, bias=self.g.b)

    def bwd(self, g, S, sink):
        pl, B, H, W = self.plan, self.plan.B, 2 * self.H, 2 * self.W
        xu = S.pop(id(self))
        pl.wgrad(xu, g, self.g, T3, B, H, W)
        dx = F.upsample2x_bwd(F.tap_gemm(g, self.g.pack1, T3, B, H, W))
        if sink.cs is not None:
            F.colsum_rows(dx, sink.cs)
        return dx


class _HeadNode:
    """GroupNorm -> SiLU -> conv3x3 to Cout <= 8 channels, NCHW fp32 out (src/modules/unet.py:437-441)."""

    def __init__(self, plan, norm, conv, H, W):
        self.plan, self.norm, self.conv, self.H, self.W = plan, norm, conv, H, W
        self.C = conv.weight.shape[1]
        self.co = conv.weight.shape[0]
        self.g = plan._gemm(conv.weight, conv.bias, 9, cout_pad=8, need_dgrad=False)
        self.gd = plan._gemm(conv.weight, None, 9, cout_pad=32, need_fwd=False)  # dgrad operand [Cin, 9, 32]
        self.pn = plan._cols(2 * self.C)
        plan._small(norm.weight, self.pn)
        plan._small(norm.bias, self.pn + self.C)
        plan.direct_small.append(conv.bias)
        plan._need_wgrad_ws(plan.B, H, W, self.C, 32, 9, main=True)

    def fwd(self, x, S):
        pl, B, H, W = self.plan, self.plan.B, self.H, self.W
        n, m, r = F.gn_silu_fwd(x, self.norm.weight, self.norm.bias, self.norm.num_groups, self.norm.eps, True)
        if not hasattr(self, "bias8") or self.bias8.device != x.device:
            self.bias8 = torch.zeros(8, dtype=f32, device=x.device)
        self.bias8[: self.co].copy_(self.conv.bias.detach())
        y8 = F.tap_gemm(n, self.g.pack0, T3, B, H, W, bias=self.bias8, out_dtype=f32)
        y = torch.empty((B, self.co, H, W), dtype=f32, device=x.device)
        L.call("pddm_nhwc_slice_to_nchw", L.ptr(y8), L.ptr(y), B, self.co, H * W, 8, L.stream())
        if S is not None:
            S[id(self)] = (x, n, m, r)
        return y

    def bwd(self, dout, S):
        pl, B, H, W = self.plan, self.plan.B, self.H, self.W
        x, n, m, r = S.pop(id(self))
        dyp = torch.empty((B, H, W, 32), dtype=bf16, device=x.device)
        L.call("pddm_nchw_to_nhwc_padded", L.ptr(dout.float().contiguous()), L.ptr(dyp), B, self.co, H * W, 32, L.stream())
        dn = F.tap_gemm(dyp, self.gd.pack1, T3, B, H, W)
        dw32 = F.tap_wgrad(n, dyp, T3, B, H, W, self.C, 32, (32, self.C, 3, 3), ws=pl.wg_ws_main)
        self.g.gw.copy_(dw32[: self.co])
        pl.gview(self.conv.bias).copy_(F.colsum(dyp, 32)[: self.co])
        dx, _ = pl.gn_bwd(x, None, dn, self.norm, m, r, True, None, pl.sink(self), None, self.pn)
        return dx


# ================================================================================================ autograd seam
class PlanFunction(torch.autograd.Function):
    """``model(x, t)`` under autograd: the plan's forward, and its hand-scheduled backward as ONE autograd node.
    Parameter gradients are returned as views of a fresh copy of the gradient arena (one copy kernel), so repeated
    backward passes accumulate into ``p.grad`` exactly as with any other autograd node."""

    @staticmethod
    def forward(ctx, plan, x, t, *params):
        plan.refresh_packs()
        out, S = plan.forward(x.detach(), t, save=True)
        ctx.plan, ctx.S = plan, S
        return out

    @staticmethod
    def backward(ctx, dout):
        plan, S = ctx.plan, ctx.S
        ctx.S = None
        if S is None:
            raise RuntimeError("PlanFunction: backward called twice (retain_graph is not supported on this path)")
        comm, plan.comm = plan.comm, None  # the overlapped all-reduce belongs to the captured step, not to autograd
        try:
            plan.backward(S, dout.contiguous(), overlap=OVERLAP_EAGER[0])
        finally:
            plan.comm = comm
        flat = plan.grad_arena.clone()
        grads = []
        for prm in plan.params:
            o = plan.goff[id(prm)]
            grads.append(flat[o: o + prm.numel()].view(prm.shape))
        return (None, None, None, *grads)


OVERLAP_EAGER = [True]
_PLANS = __import__("weakref").WeakKeyDictionary()
ENABLED = [__import__("os").environ.get("PDDM_NO_PLAN") != "1"]


def plan_for(model, x):
    """The cached plan of ``model`` for this input shape / device, or None if the model takes the op-by-op path."""
    if not ENABLED[0] or not x.is_cuda:
        return None
    cache = _PLANS.get(model)
    if cache is None:
        cache = _PLANS[model] = {"ok": supported(model)}
    if not cache["ok"]:
        return None
    key = (x.shape[0], x.shape[2], x.shape[3], x.device.index)
    plan = cache.get(key)
    if plan is not None and not plan.matches(model):
        plan = None
    if plan is None:
        if any(p_.device != x.device for p_ in model.parameters()):
            return None
        with torch.cuda.device(x.device):
            plan = cache[key] = UNetPlan(model, x.shape[0], x.shape[2], x.shape[3], x.device)
    return plan
