#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: train img/s (+ 1000-step samples/s) of the CIFAR-10-shape
improved-diffusion UNet (config/model/unet.yaml, cosine schedule, learned variance, L_hybrid, bf16 compute).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one optimisation step (q_sample, UNet forward, L_hybrid, backward, Adam) on 128 synthetic images per
GPU (weak scaling; the gradient all-reduce is part of the step for N > 1).  The timed region replays the captured
CUDA graph K times between CUDA events; rank 0 prints ONE JSON line.  ``--impl reference`` times the CPU oracle
port of the reference's own PyTorch path (oracle/) on the host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# --config: BASELINE.json configs[1] (default; the configuration the metric is quoted on) and configs[4]
CONFIGS = {
    "cifar": dict(model="unet", res=32, batch=128,
                  metric="train img/s (+ 1000-step samples/s), CIFAR-10 32x32 improved-diffusion UNet",
                  workload="BASELINE configs[1]: CIFAR-10-shape 3x32x32 UNet (config/model/unet.yaml), cosine "
                           "schedule, learned variance, L_hybrid, Adam, bf16 compute / fp32 accumulate"),
    "celeba64": dict(model="unet_celeba", res=64, batch=64,
                     metric="train img/s (+ 1000-step samples/s), CelebA 64x64 UNet with attention at 16x16 / 8x8",
                     workload="BASELINE configs[4]: CelebA-shape 3x64x64 UNet (config/model/unet_celeba.yaml), cosine "
                              "schedule, learned variance, L_hybrid, Adam, bf16 compute / fp32 accumulate"),
}
METRIC, MODEL, RES, PER_GPU_BATCH = (CONFIGS["cifar"][k] for k in ("metric", "model", "res", "batch"))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), p["hbm_gbs"], "measured"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # the sampler is started before the warm-up steps (nvidia-smi takes a second to come up, longer than the
        # timed region itself): keep the samples taken inside the timed region; if the region was shorter than the
        # sampling period, the samples of the (identical, back-to-back) warm-up steps right before it stand in
        inside = [r for (ts, r) in self.rows if self.t0 is not None and self.t0 <= ts <= (self.t1 or ts)]
        scope = "timed region"
        if not inside:
            inside = [r for (ts, r) in self.rows if self.t0 is None or ts <= (self.t1 or ts)][-10:]
            scope = "warm-up steps immediately before the timed region"
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed summary of
    an ``ncu --set full`` capture on this tree (profiles/r2_conv_fwd_ncu_traffic.json); None if no capture exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_conv_fwd_ncu_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_launch"]), t.get("note", "")
    except Exception:
        return None, "no ncu --set full capture of this kernel committed for this tree"


def dominant_kernel_roofline(dev, B, burst_tflops, src, census):
    """Time the dominant kernel (conv_fwd_kernel: the tcgen05 tap-GEMM behind every conv forward and data gradient)
    live with CUDA events on the model's own 3x3 conv census (configs.conv3x3_census; stride-2 convs counted at their
    output resolution): one launch per layer, distinct input/weight tensors per layer (> L2 in total).
    achieved = algorithmic FLOPs of the census / time of the census."""
    import torch

    from probabilisticdeepdiffusionmodels_b200 import _lib, ops
    P = torch.ops.pddm
    g = torch.Generator(device=dev).manual_seed(0)
    layers, flops = [], 0.0
    for H, cin, cout, count in census:
        for _ in range(count):
            x = torch.randn((B, H, H, cin), generator=g, device=dev).to(torch.bfloat16)
            w = torch.randn((cout, cin, 3, 3), generator=g, device=dev) * 0.02
            b = torch.zeros(cout, device=dev)
            layers.append((x, w, b))
            flops += 2.0 * B * H * H * cout * cin * 9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream(dev)
    with torch.no_grad(), ops.frozen_weights(), torch.cuda.stream(stream):
        for _ in range(3):  # warm-up (weight packs cached)
            for x, w, b in layers:
                P.conv2d(x, w, b, None, None, 1, False)
        torch.cuda.synchronize(dev)
        # one pass over the census captured in a CUDA graph, as the product path runs it (eager launches of the
        # 4x4 / 8x8 layers would time the host-side tensor-map encode, not the kernel)
        graph = torch.cuda.CUDAGraph()
        k0 = _lib.KERNELS[0]
        with torch.cuda.graph(graph, stream=stream):
            for x, w, b in layers:
                P.conv2d(x, w, b, None, None, 1, False)
        launches = _lib.KERNELS[0] - k0
        graph.replay()
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            graph.replay()
        e1.record(stream)
        torch.cuda.synchronize(dev)
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    ach = flops / sec / 1e12
    traffic, note = measured_traffic()
    return {"bound": "tensor", "achieved": ach, "peak": burst_tflops, "unit": "TFLOP/s", "frac": ach / burst_tflops,
            "traffic": traffic, "traffic_note": note,
            "kernel": "pddm::conv_fwd_kernel (conv_fwd_swap_kernel, its swapped-operand variant, for <=128 output channels)",
            "launches_timed": int(launches), "us_per_launch": sec / launches * 1e6,
            "flops_per_census": flops, "peak_source": f"{src} (burst bf16, kernel timed alone)",
            "workload": f"forward 3x3 conv census of the {MODEL} model at B={B} ({int(launches)} launches), CUDA events"}


def cpu_reference_arm(steps, warmup, batch=None, threads=None):
    """The reference's own PyTorch path (CPU oracle port, fp32): training step with Adam + a few reverse steps."""
    import numpy as np
    import torch

    from oracle import engine_ref as E
    from oracle.diffusion_ref import DiffusionRef
    from oracle.unet_ref import MODEL_CONFIGS, arch_from_config, make_params

    batch = batch or PER_GPU_BATCH
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = MODEL_CONFIGS[MODEL]
    arch = arch_from_config(RES, **{k: v for k, v in cfg.items() if k != "name"}, learn_sigma=True)
    P = {k: v.requires_grad_(True) for k, v in make_params(arch, seed=1).items()}
    diff = DiffusionRef(1000, mode="cosine")
    opt = torch.optim.Adam(list(P.values()), lr=1e-4)
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand((batch, 3, RES, RES), generator=g) * 2 - 1

    def step():
        t = torch.randint(1, 1001, (batch,), generator=g)
        noise = torch.randn(x0.shape, generator=g)
        opt.zero_grad(set_to_none=True)
        loss, _ = E.train_loss(P, arch, diff, x0, t, noise, learn_sigma=True)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    # sampling: a few reverse steps, extrapolated to 1000-step chains (stated in `sample`)
    xs = torch.randn((batch, 3, RES, RES), generator=g)
    nst = 3
    with torch.no_grad():
        E.denoising_step(P, arch, diff, xs, 1000, torch.randn(xs.shape, generator=g), True, learn_sigma=True)
        t1 = time.perf_counter()
        for k in range(nst):
            xs = E.denoising_step(P, arch, diff, xs, 999 - k, torch.randn(xs.shape, generator=g), True,
                                  learn_sigma=True)
        ds = (time.perf_counter() - t1) / nst
    return {"train_img_s": batch / dt, "ms_per_step": dt * 1e3, "image_steps_s": batch / ds,
            "samples_s": batch / ds / 1000.0, "cores": threads, "batch": batch,
            "sample": f"{steps} timed train steps (+{warmup} warm-up) at B={batch} fp32 on {threads} host threads; "
                      f"sampling = {nst} reverse steps at B={batch} extrapolated x1000/{nst} to 1000-step chains"}


def unmodified_reference_arm(steps, warmup, batch, threads=None):
    """The reference's OWN Engine (staged under baseline/_ref by __graft_entry__.build(); git-ignored) timed on the
    host cores: ``Engine.training_step`` + backward + torch Adam on the same architecture, schedule and batch.  The
    reference hard-codes learn_sigma=False (src/modules/__init__.py:34), so this is L_simple with the 3-channel head --
    what the reference actually trains; the oracle port (``cpu_reference_arm``) adds the learned-variance L_hybrid of
    the measured arm.  Returns None when the staged copy is absent."""
    root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(root, "src")):
        return None
    import torch

    os.environ["PDDM_REFERENCE_ROOT"] = root
    from oracle import ref_shims
    from oracle.unet_ref import MODEL_CONFIGS
    ref_shims.REFERENCE_ROOT = root
    ref_shims.install()
    from src.engine import Engine  # the reference's, not this package's

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    with __import__("contextlib").redirect_stdout(sys.stderr):  # the reference prints its settings: keep stdout = one JSON line
        eng = Engine(dict(MODEL_CONFIGS[MODEL]), {"lr": 1e-4}, diffusion_steps=1000, mode="cosine", resolution=RES)
    opt = torch.optim.Adam(eng.parameters(), lr=1e-4)
    x0 = torch.rand((batch, 3, RES, RES)) * 2 - 1

    def step():
        opt.zero_grad(set_to_none=True)
        loss = eng.training_step((x0, None), 0)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"train_img_s": batch / dt, "ms_per_step": dt * 1e3, "cores": threads, "batch": batch,
            "what": "unmodified reference Engine.training_step + backward + torch.optim.Adam (L_simple, fixed variance: "
                    "the reference has no learned-variance path), fp32, CPU"}


def micro_diffusion():
    """HBM evidence for the fused elementwise diffusion kernels (north_star group 4): time
    q_sample / p_sample / vlb / sq_err at the training batch (B=128: 1.5 MB tensors, launch-bound) and at a
    bandwidth-meaningful size (B=16384: 201 MB tensors, > L2), 20 launches in a CUDA graph, CUDA events; GB/s =
    algorithmic bytes (every fp32 tensor read or written once) / time, against the measured HBM peak."""
    import torch

    from probabilisticdeepdiffusionmodels_b200 import functional as F
    from probabilisticdeepdiffusionmodels_b200 import schedules
    burst, sustained, hbm, src = peaks()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    tabs = F.DeviceTables(schedules.make_tables(schedules.get_betas(diffusion_steps=1000, mode="cosine").float()), dev)
    out = {"metric": "GB/s of the fused diffusion kernels (fp32, 3x32x32 images)", "unit": "GB/s", "peak": hbm,
           "peak_source": src, "n_gpus": 1, "dtype": "f32", "data": "synthetic", "kernels": {}}
    g = torch.Generator(device=dev).manual_seed(0)
    for B in (128, 16384):
        x0 = torch.rand((B, 3, 32, 32), generator=g, device=dev) * 2 - 1
        noise = torch.randn((B, 3, 32, 32), generator=g, device=dev)
        t = torch.randint(1, 1001, (B,), generator=g, device=dev)
        mo = torch.randn((B, 6, 32, 32), generator=g, device=dev)  # eps | v (learned variance)
        z = torch.randn((B, 3, 32, 32), generator=g, device=dev)
        mo3 = mo[:, :3].contiguous()
        x_t = F.q_sample(x0, noise, t, tabs)
        gsc = torch.ones(B, device=dev)
        n = x0.numel() * 4  # bytes of one [B, 3, 32, 32] fp32 tensor
        cases = {
            "q_sample": (lambda: F.q_sample(x0, noise, t, tabs), 3 * n),                        # x0, noise -> x_t
            "p_sample(learned sigma)": (lambda: F.p_sample_step(x_t, mo, z, 500, tabs, True, "learned"), 5 * n),
            "p_sample(fixed sigma)": (lambda: F.p_sample_step(x_t, mo3, z, 500, tabs, True, "beta"), 4 * n),
            "vlb(L_vlb, learned sigma, +grad_v)": (lambda: F.vlb_terms(x0, x_t, mo, t, tabs, 1, "learned", True), 5 * n),
            "sq_err(+grad)": (lambda: F.sq_err(mo, noise, gsc, True), 5 * n),                  # pred(6ch), noise -> grad(6ch)
        }
        for name, (fn, nbytes) in cases.items():
            try:
                fn()
            except Exception as e:  # a sigma mode this build does not name the same way: report, do not die
                out["kernels"][f"{name} B={B}"] = {"error": str(e)[:80]}
                continue
            s_ = torch.cuda.Stream(dev)
            with torch.cuda.stream(s_):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=s_):
                    for _ in range(20):
                        fn()
                gr.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s_)
                for _ in range(3):
                    gr.replay()
                e1.record(s_)
            e1.synchronize()
            us = e0.elapsed_time(e1) / 60 * 1e3
            gbs = nbytes / us * 1e-3
            out["kernels"][f"{name} B={B}"] = {"us": round(us, 2), "GB/s": round(gbs, 1), "frac": round(gbs / hbm, 3),
                                              "bytes": nbytes}
    best = max((v.get("GB/s", 0) for v in out["kernels"].values()), default=0)
    out["value"] = best
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bucketed-allreduce", action="store_true",
                    help="N>1: all-reduce bucket by bucket from inside backward on a communication stream")
    ap.add_argument("--allreduce", default="overlap", choices=["overlap", "flat"],
                    help="N > 1: arena all-reduce hidden under the backward pass (default) or one call after it")
    ap.add_argument("--overlap-comm", default="own", choices=["own", "own-default", "same"],
                    help="communicator of the overlapped ranges: own (max_ctas = sm reserve), own-default, same")
    ap.add_argument("--sm-reserve", type=int, default=8,
                    help="SMs left to NCCL while an overlapped all-reduce is in flight (= its max_ctas)")
    ap.add_argument("--config", default="cifar", choices=sorted(CONFIGS),
                    help="cifar = BASELINE configs[1] (the metric's configuration); celeba64 = configs[4]")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch images per GPU; strong: --batch images in total, split over the GPUs")
    ap.add_argument("--micro", default=None, choices=["diffusion"],
                    help="diffusion: GB/s of the fused q_sample / p_sample / vlb / sq_err kernels instead of the step")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (weak) / global batch (strong)")
    ap.add_argument("--sample-steps", type=int, default=1000, help="length of the timed reverse chain")
    ap.add_argument("--no-sampling", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="keep weight-gradient GEMMs on the main stream")
    args = ap.parse_args()
    if os.environ.get("PDDM_BENCH_WATCHDOG"):  # debugging aid: dump every thread's Python stack and exit if we hang
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["PDDM_BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    global METRIC, MODEL, RES, PER_GPU_BATCH
    METRIC, MODEL, RES = (CONFIGS[args.config][k] for k in ("metric", "model", "res"))
    if args.batch is None:
        args.batch = CONFIGS[args.config]["batch"]
    if args.scaling == "strong":
        if args.batch % max(world, 1):
            raise SystemExit("--scaling strong: the global batch must divide by the number of GPUs")
        global_batch, args.batch = args.batch, args.batch // max(world, 1)
    else:
        global_batch = args.batch * max(world, 1)
    PER_GPU_BATCH = args.batch

    cfg_json = {"workload": CONFIGS[args.config]["workload"],
                "per_gpu_batch": args.batch, "global_batch": global_batch, "parallelism": f"dp{world}",
                "l2_policy": "activations + weights + grads per step (>2 GB) exceed the 126 MB L2; no explicit flush"}

    if args.micro == "diffusion":
        if rank == 0:
            micro_diffusion()
        return

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's CPU path on the SAME configuration (batch included); every step is a whole batch, the
        # run is bounded by timing at most 3 steps (~4 s each at B=128 on 16 host threads)
        steps = max(1, min(args.steps, 3))
        warm = max(1, min(args.warmup, 1))
        r = cpu_reference_arm(steps, warm, batch=args.batch)
        try:
            unmod = unmodified_reference_arm(max(1, steps - 1), 1, args.batch)
        except Exception as e:  # the staged copy is optional evidence; the port is the arm
            unmod = {"error": str(e)[:200]}
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["train_img_s"], "unit": "img/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_json,
            "sampling": {"value": r["samples_s"], "unit": "samples/s (1000-step chains)",
                         "image_steps_per_s": r["image_steps_s"]},
            "cpu_baseline": {"value": r["train_img_s"], "unit": "img/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "unmodified_reference": unmod,
            "e2e": {"value": r["train_img_s"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist

    # (nothing on this arm imports oracle/: configurations, the synthetic initialisation and the FLOP counter
    #  live in the package; the oracle is only executed by the cpu_baseline / --impl reference leg above)
    from probabilisticdeepdiffusionmodels_b200 import Engine, _lib, parallel
    from probabilisticdeepdiffusionmodels_b200.configs import (MODEL_CONFIGS, conv3x3_census, synthetic_init_,
                                                               unet_fwd_flops_per_image)

    rank, world, local_rank = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = MODEL_CONFIGS[MODEL]
    torch.manual_seed(1)
    eng = Engine(dict(cfg), {"lr": 1e-4}, diffusion_steps=1000, mode="cosine", resolution=RES,
                 clip_while_generating=True, learn_sigma=True, log_loss_per_t=False)
    synthetic_init_(eng.model, seed=1)  # random init incl. the reference's zero-init convs
    fwd_flops = unet_fwd_flops_per_image(eng.model, RES)
    census = conv3x3_census(eng.model, RES)
    eng = eng.to(dev)
    parallel.broadcast_parameters(eng.model)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    host_x = (torch.rand((B, 3, RES, RES), generator=g) * 2 - 1).pin_memory()
    x_dev = host_x.to(dev)
    # gradient averaging: one flat all-reduce between backward and Adam.  --bucketed-allreduce issues it bucket by
    # bucket from inside backward on a communication stream; measured equal at 2 GPUs (the persistent GEMM kernels
    # leave NCCL no SM to overlap on), so the simpler form is the default.
    hook, optimizer = None, None
    if world > 1:
        from probabilisticdeepdiffusionmodels_b200 import plan as _plan
        from probabilisticdeepdiffusionmodels_b200.optim import FusedAdam
        if args.bucketed_allreduce:
            hook = parallel.BucketedGradAllReduce(eng.model.parameters())
        elif _plan.plan_for(eng.model, x_dev) is not None:
            # gradients live in one flat arena: all-reduce it in place, fold 1/W into the Adam kernel; by default the
            # tail ranges of the arena are reduced under the backward pass (--allreduce flat: one call after it)
            optimizer = FusedAdam(eng.model.parameters(), lr=1e-4)
            if args.allreduce == "overlap":
                hook = parallel.OverlappedArenaAllReduce(optimizer, sm_reserve=args.sm_reserve, comm=args.overlap_comm)
            else:
                hook = parallel.ArenaGradAllReduce(optimizer)
        else:
            hook = parallel.FlatGradAllReduce()

    step = eng.capture_train_step((B, 3, RES, RES), optimizer=optimizer, grad_hook=hook,
                                  overlap_wgrad=not args.no_overlap)
    kernels_per_step = step.state["kernels_per_step"]  # this library's kernels inside one captured step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        for _ in range(max(args.warmup, 3)):
            step(x_dev)
        # keep the GPU under the same load until the sampler has produced its first rows -- on EVERY rank, and every
        # rank leaves the loop after the same number of steps (a step contains collectives: ranks that replayed the
        # graph a different number of times would pair mismatched all-reduces)
        deadline = time.perf_counter() + 4.0
        while not parallel.all_ranks(clk.proc is None or bool(clk.rows) or time.perf_counter() >= deadline, dev):
            step(x_dev)
        barrier()
        clk.mark_start()
        e0.record()
        for _ in range(args.steps):
            eng._train_graph["graph"].replay()
        e1.record()
        barrier()
        clk.mark_end()
    ms = e0.elapsed_time(e1) / args.steps

    # end-to-end through the public step(x): pinned host batch -> device each step, loss read back each step
    for _ in range(2):
        float(step(host_x))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        loss_val = float(step(host_x))  # H2D copy (non_blocking, pinned) + graph replay + D2H of the loss (sync)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    _ = t0

    t_ms = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_ms[0]), float(t_ms[1])
    img_s = B * world / (ms * 1e-3)
    img_s_e2e = B * world / (ms_e2e * 1e-3)

    # ---- sampling: one full reverse chain per GPU (batch-sharded, no collective)
    sampling = None
    if not args.no_sampling:
        gen = torch.Generator(device=dev).manual_seed(99 + rank)
        xT = torch.randn((B, 3, RES, RES), generator=gen, device=dev)
        with torch.no_grad():
            eng.sample_from_step(xT, 5, generator=gen)  # warm-up + graph capture
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            out = eng.sample_from_step(xT, args.sample_steps, generator=gen)
            s1.record()
            barrier()
        chain_ms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(chain_ms, op=dist.ReduceOp.MAX)
        chain_s = float(chain_ms[0]) * 1e-3
        steps_s = B * world * args.sample_steps / chain_s
        sampling = {"value": steps_s / 1000.0, "unit": "samples/s (1000-step chains)", "image_steps_per_s": steps_s,
                    "chain_steps_timed": args.sample_steps, "chain_seconds": chain_s, "batch_per_gpu": B,
                    "finite": bool(torch.isfinite(out).all())}

    def finish():
        # Multi-rank teardown.  dist.destroy_process_group() and a plain interpreter exit were both observed to hang
        # here once the captured step (NCCL all-reduce node + side-stream weight-gradient branch) had been replayed,
        # so every rank leaves through os._exit.  The other ranks first wait (on the rendezvous store, not on NCCL)
        # until rank 0 has printed its line: a rank that disappears while rank 0 is still timing the roofline census
        # can get rank 0 torn down by the NCCL watchdog.
        if world > 1:
            torch.cuda.synchronize(dev)
            try:
                store = dist.distributed_c10d._get_default_store()
                if rank == 0:
                    store.set("pddm_bench_done", "1")
                    time.sleep(0.5)  # let the waiters read the key before the store's server (this process) exits
                else:
                    store.wait(["pddm_bench_done"], __import__("datetime").timedelta(seconds=600))
            except Exception:
                pass
            sys.stdout.flush()
            sys.stderr.flush()
            try:
                __import__("ctypes").CDLL(None).fflush(None)  # C stdio buffers (NCCL's banner) are not flushed by _exit
            except Exception:
                pass
            os._exit(0)

    if rank != 0:
        finish()
        return
    burst, sustained, hbm, src = peaks()
    fwd = fwd_flops
    train_flops = 3 * fwd
    achieved = img_s / world * train_flops / 1e12  # per GPU
    roof_step = {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                 "frac": achieved / sustained, "peak_source": f"{src} (sustained bf16; burst {burst})",
                 "scope": "whole train step per GPU (algorithmic 3 x fwd FLOPs/img x img/s)"}
    roof = dominant_kernel_roofline(dev, B, burst, src, census)
    if sampling is not None:
        s_ach = sampling["image_steps_per_s"] / world * fwd / 1e12
        sampling["roofline_frac"] = s_ach / sustained
        sampling["achieved_tflops_per_gpu"] = s_ach
    line = {"metric": METRIC, "value": img_s, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg_json, "clocks": clk.summary(),
            "e2e": {"value": img_s_e2e, "unit": "img/s", "h2d_bytes_per_step": host_x.numel() * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "last_loss": loss_val},
            "gpu_launches": int(kernels_per_step * args.steps), "kernels_per_step": int(kernels_per_step),
            "roofline": roof, "roofline_step": roof_step, "sampling": sampling}
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_arm(2, 1, batch=min(B, 32))
        line["cpu_baseline"] = {"value": r["train_img_s"], "unit": "img/s", "cores": r["cores"], "kind": "port",
                                "sample": r["sample"], "samples_per_s": r["samples_s"]}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
