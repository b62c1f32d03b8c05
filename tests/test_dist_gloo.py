"""world_size-2 gloo test of the N>1 host logic: contiguous batch sharding + the usage-histogram /
loss all-reduce reproduce the single-process result exactly (integer counts) — SURVEY.md §8e."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from paintmind_b200 import dist as pmdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_tokenize(lo, hi, n_e=64, tokens=16):
    """Deterministic stand-in for encode(): indices depend only on the GLOBAL image index."""
    g = torch.Generator().manual_seed(1234)
    all_idx = torch.randint(0, n_e, (40, tokens), generator=g)
    all_sse = torch.rand(40, generator=g).double()
    return all_idx[lo:hi], all_sse[lo:hi]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pmdist.shard_range(total, rank, world)
    idx, sse = _fake_tokenize(lo, hi)
    hist = torch.bincount(idx.reshape(-1), minlength=64)
    sums = torch.tensor([float(sse.sum()), float(idx.numel() * 32)], dtype=torch.float64)
    pmdist.allreduce_usage(hist, sums)
    q.put((rank, lo, hi, hist.tolist(), sums.tolist()))      # plain lists: no shared-memory handles outlive the worker
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 8192):
        for world in (1, 2, 3, 8):
            spans = [pmdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_allreduce_usage_world2_equals_single_process():
    total, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx, sse = _fake_tokenize(0, total)
    want_hist = torch.bincount(idx.reshape(-1), minlength=64)
    want_sums = torch.tensor([float(sse.sum()), float(idx.numel() * 32)], dtype=torch.float64)
    for rank, lo, hi, hist, sums in outs:
        hist, sums = torch.tensor(hist, dtype=torch.int64), torch.tensor(sums, dtype=torch.float64)
        assert torch.equal(hist, want_hist)                       # integer counts: bit-exact
        torch.testing.assert_close(sums, want_sums, atol=1e-12, rtol=1e-12)
        assert abs(pmdist.global_loss(sums) - 1.25 * float(want_sums[0] / want_sums[1])) < 1e-15


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.LayerNorm(7), torch.nn.Linear(7, 3))
    net[2].bias.requires_grad_(False)                                       # a parameter without a gradient is skipped
    g = torch.Generator().manual_seed(100 + rank)
    for p in net.parameters():
        if p.requires_grad:
            p.grad = torch.randn(p.shape, generator=g)
    bucket = pmdist.allreduce_gradients(net)
    bucket2 = pmdist.allreduce_gradients(net, bucket=bucket, average=False)  # second call reuses the bucket; SUM of the means
    q.put((rank, [p.grad.tolist() for p in net.parameters() if p.grad is not None], bucket2.data_ptr() == bucket.data_ptr()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_gradients_world2():
    """One flattened all-reduce averages every gradient over the ranks (data-parallel generator training)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shapes = [(7, 5), (7,), (7,), (7,), (3, 7)]
    per_rank = []
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        per_rank.append([torch.randn(s, generator=g) for s in shapes])
    mean = [(a + b) / 2 for a, b in zip(*per_rank)]
    for rank, grads, reused in outs:
        assert reused
        assert len(grads) == len(shapes)
        for got, want in zip(grads, mean):
            torch.testing.assert_close(torch.tensor(got), want * world, atol=1e-6, rtol=1e-6)   # mean, then summed again over 2 ranks
