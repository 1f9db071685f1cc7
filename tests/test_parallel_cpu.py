"""Data-parallel host logic over gloo, world_size 2, on CPU: batch sharding and the flat-bucket gradient
all-reduce (the N>1 path of bench.py uses the same code over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from probabilisticdeepdiffusionmodels_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    if rank == 1:  # ranks start different: the broadcast must fix that
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    parallel.broadcast_parameters(net)
    x = torch.arange(7 * 6, dtype=torch.float32).reshape(7, 6) / 10.0
    y = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3) / 5.0
    xs, ys = parallel.shard_batch(x, rank, world), parallel.shard_batch(y, rank, world)
    assert xs.shape[0] == (4 if rank == 0 else 3)
    # per-rank loss scaled so that AVERAGING the rank gradients reproduces the global-batch mean loss
    loss = ((net(xs) - ys) ** 2).sum() / x.shape[0] * world
    loss.backward()
    params = list(net.parameters())
    parallel.FlatGradAllReduce()(params)
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        ref.load_state_dict(net.state_dict())
        (((ref(x) - y) ** 2).sum() / x.shape[0]).backward()
        errs = [float((a.grad - b.grad).abs().max()) for a, b in zip(params, ref.parameters())]
        torch.save({"errs": errs}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_equals_single_process_gradient(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    errs = torch.load(out)["errs"]
    assert max(errs) < 1e-5, errs


def _worker_bucketed(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from probabilisticdeepdiffusionmodels_b200 import parallel
    parallel.init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 50), torch.nn.Tanh(), torch.nn.Linear(50, 40), torch.nn.Tanh(),
                              torch.nn.Linear(40, 3))
    params = list(net.parameters())
    hook = parallel.BucketedGradAllReduce(params, bucket_bytes=4000)  # several buckets for this tiny net
    x = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10.0
    y = torch.arange(8 * 3, dtype=torch.float32).reshape(8, 3) / 5.0
    xs, ys = parallel.shard_batch(x, rank, world), parallel.shard_batch(y, rank, world)
    errs, nb = [], 0
    for it in range(3):  # pass 0 learns the order, passes 1-2 reduce bucket by bucket from inside backward
        for p in params:
            p.grad = None
        (((net(xs) - ys) ** 2).sum() / x.shape[0] * world).backward()
        hook(params)
        nb = len(hook.buckets)
        ref = torch.nn.Sequential(torch.nn.Linear(6, 50), torch.nn.Tanh(), torch.nn.Linear(50, 40), torch.nn.Tanh(),
                                  torch.nn.Linear(40, 3))
        ref.load_state_dict(net.state_dict())
        (((ref(x) - y) ** 2).sum() / x.shape[0]).backward()
        errs.append(max(float((a.grad - b.grad).abs().max()) for a, b in zip(params, ref.parameters())))
        with torch.no_grad():
            for p in params:
                p.add_(p.grad, alpha=-0.05)  # move, so that every pass has different gradients
    if rank == 0:
        torch.save({"errs": errs, "nb": nb}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_overlapped_allreduce_equals_single_process_gradient(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_bucketed, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["nb"] >= 2
    assert max(res["errs"]) < 1e-5, res


def test_shard_batch_covers_everything():
    from probabilisticdeepdiffusionmodels_b200.parallel import shard_batch
    x = torch.arange(10)
    for world in (1, 2, 3, 4, 8):
        parts = [shard_batch(x, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts), x)
        assert max(p.numel() for p in parts) - min(p.numel() for p in parts) <= 1


def test_single_process_is_a_noop():
    from probabilisticdeepdiffusionmodels_b200.parallel import FlatGradAllReduce
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    FlatGradAllReduce()([p])
    assert p.grad.tolist() == [2.0, 2.0, 2.0]


def _worker_arena(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from probabilisticdeepdiffusionmodels_b200 import parallel
    parallel.init_from_env("gloo")

    class Plan:  # what ArenaGradAllReduce needs of plan.UNetPlan: the flat fp32 gradient arena
        grad_arena = torch.arange(5000, dtype=torch.float32) * (rank + 1)

    class Opt:
        grad_scale = 1.0

    opt = Opt()
    for buckets in (1, 3):
        Plan.grad_arena = torch.arange(5000, dtype=torch.float32) * (rank + 1)
        parallel.ArenaGradAllReduce(opt, buckets=buckets)(Plan)
        assert opt.grad_scale == 1.0 / world  # the 1/W lives in the optimizer kernel
        want = torch.arange(5000, dtype=torch.float32) * sum(r + 1 for r in range(world))
        assert torch.equal(Plan.grad_arena, want)
    Plan.grad_arena = torch.ones(100) * (rank + 1)
    parallel.ArenaGradAllReduce(None)(Plan)  # no optimizer to carry the scale: averaged in place
    assert torch.allclose(Plan.grad_arena, torch.full((100,), 1.5))
    # overlapped form: the backward pass reports finished tail ranges, the hook reduces the head and joins
    opt2 = Opt()
    hook = parallel.OverlappedArenaAllReduce(opt2)
    want = torch.arange(5000, dtype=torch.float32) * sum(r + 1 for r in range(world))
    for _ in range(2):  # reusable step after step
        Plan.grad_arena = torch.arange(5000, dtype=torch.float32) * (rank + 1)
        Plan.comm = None
        hook.attach(Plan)
        assert Plan.comm is hook
        hook.range_final(Plan, 3000, 5000)
        assert torch.equal(Plan.grad_arena[3000:], want[3000:]) and torch.equal(
            Plan.grad_arena[:3000], torch.arange(3000, dtype=torch.float32) * (rank + 1))
        hook.range_final(Plan, 1024, 3000)
        hook(Plan)
        assert opt2.grad_scale == 1.0 / world and torch.equal(Plan.grad_arena, want)
    try:  # ranges must be contiguous, tail first
        hook.range_final(Plan, 3000, 5000)
        hook.range_final(Plan, 0, 1000)
        raise AssertionError("non-contiguous range accepted")
    except RuntimeError:
        hook._done_from = None
    if rank == 0:
        torch.save({"ok": True}, out)
    dist.barrier()
    dist.destroy_process_group()


def _worker_all_ranks(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from probabilisticdeepdiffusionmodels_b200 import parallel
    parallel.init_from_env("gloo")
    assert parallel.all_ranks(True) is True
    assert parallel.all_ranks(rank == 0) is False  # one rank not ready: nobody leaves
    assert parallel.all_ranks(False) is False
    # a loop with a collective in its body, left on all_ranks: every rank runs it the same number of times although
    # the local condition turns true at different iterations
    n, total = 0, torch.zeros(1)
    while not parallel.all_ranks(n >= 2 + 3 * rank):
        dist.all_reduce(total.add_(1.0))
        n += 1
    assert n == 2 + 3 * (world - 1)
    if rank == 0:
        torch.save({"ok": True, "n": n}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_loops_with_collectives_leave_together(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_all_ranks, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out)["ok"]
    from probabilisticdeepdiffusionmodels_b200.parallel import all_ranks
    assert all_ranks(True) is True and all_ranks(False) is False  # single process: the local flag


def test_arena_grad_allreduce(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_arena, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out)["ok"]


def test_shard_timesteps_covers_every_term_once():
    from probabilisticdeepdiffusionmodels_b200.parallel import shard_timesteps
    for T in (2, 3, 20, 1000):
        for world in (1, 2, 3, 8):
            got = sorted(t for r in range(world) for t in shard_timesteps(T, r, world))
            assert got == list(range(2, T + 1))


def _worker_nll(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from probabilisticdeepdiffusionmodels_b200 import parallel
    parallel.init_from_env("gloo")
    T, B = 50, 4
    f = lambda t: torch.arange(B, dtype=torch.float32) * 0.01 + 1.0 / t  # per-sample term of timestep t
    mine = parallel.shard_timesteps(T, rank, world)
    parts = {"L_int": sum((f(t) for t in mine), torch.zeros(B)),
             "L_0": torch.full((B,), 3.0) if rank == 0 else torch.zeros(B),
             "L_T": torch.full((B,), 0.5) if rank == 0 else torch.zeros(B),
             "mse_sum": torch.tensor(float(sum(mine)), dtype=torch.float64), "mse_count": float(len(mine))}
    tot = parallel.all_reduce_nll(parts)
    want = sum((f(t) for t in range(2, T + 1)), torch.zeros(B))
    assert torch.allclose(tot["L_int"], want, rtol=1e-6)
    assert torch.equal(tot["L_0"], torch.full((B,), 3.0)) and torch.equal(tot["L_T"], torch.full((B,), 0.5))
    assert float(tot["mse_sum"]) == float(sum(range(2, T + 1))) and tot["mse_count"] == float(T - 1)
    if rank == 0:
        torch.save({"ok": True}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_nll_reduction(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_nll, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out)["ok"]
