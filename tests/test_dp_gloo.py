"""world_size-2 gloo tests (CPU) of the data-parallel logic of the train step: the statistics
each rank contributes combine, after the all-reduces ``_train.py`` issues, to exactly the
single-device full-batch statistics."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zenflow_b200._train import _allreduce, _dist_world

    rng = np.random.default_rng(0)
    M, D, C = 1000, 6, 2
    x = rng.normal(size=(M, D))
    c = rng.uniform(size=(M, C))
    lo, hi = rank * M // world, (rank + 1) * M // world
    xs, cs = x[lo:hi], c[lo:hi]
    d = D // 2
    h = np.hstack([xs[:, d:], cs])
    sums = torch.tensor(np.concatenate([h.sum(0), (h * h).sum(0)]))
    _allreduce(sums, None, "sum")
    mm_min, mm_max = torch.tensor(xs.min(0)), torch.tensor(xs.max(0))
    _allreduce(mm_min, None, "min")
    _allreduce(mm_max, None, "max")
    cnt = torch.tensor([hi - lo])
    _allreduce(cnt, None, "sum")
    # gradient of a sum-over-samples loss: per-rank partial sums add up
    g = torch.tensor((xs[:, :1] * xs).sum(0))
    _allreduce(g, None, "sum")
    if rank == 0:
        hf = np.hstack([x[:, d:], c])
        ok = (np.allclose(sums.numpy(), np.concatenate([hf.sum(0), (hf * hf).sum(0)]), rtol=1e-12)
              and np.array_equal(mm_min.numpy(), x.min(0)) and np.array_equal(mm_max.numpy(), x.max(0))
              and int(cnt) == M and np.allclose(g.numpy(), (x[:, :1] * x).sum(0), rtol=1e-12)
              and _dist_world(None)[1] == world)
        out.put(bool(ok))
    dist.destroy_process_group()


def test_statistics_combine_across_two_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_single_process_is_world_one():
    sys.path.insert(0, ROOT)
    from zenflow_b200._train import _allreduce, _dist_world

    assert _dist_world(None)[1] == 1
    t = torch.ones(3)
    _allreduce(t, None, "sum")  # no-op
    assert t.tolist() == [1, 1, 1]


def _steps_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zenflow_b200.train import _global_step_counts

    res = {}
    # equal step counts, ragged last batch of different sizes: per-step global row counts
    n_rows = 1000 if rank == 0 else 900
    res["counts"] = _global_step_counts(n_rows, 256, torch.device("cpu"), None, world)
    # shards that disagree on the number of steps: ValueError on EVERY rank (no rank is left in a collective)
    try:
        _global_step_counts(700 if rank == 0 else 300, 256, torch.device("cpu"), None, world)
        res["raised"] = False
    except ValueError as e:
        res["raised"] = "disagree" in str(e)
    out.put((rank, res))
    dist.destroy_process_group()


def test_train_loop_step_counts_are_global_and_validated():
    """data-parallel train() must not let ranks run different numbers of collective steps."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_steps_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in (0, 1):
        assert got[r]["counts"] == [512, 512, 512, 232 + 132]
        assert got[r]["raised"] is True
