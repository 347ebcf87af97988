"""N>1 host logic on CPU: world_size-2 gloo run of the shard / all-gather / trim loop (the device sampler is replaced
by a rank-tagged stub; the real one is exercised on GPUs by bench.py --gpus N)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from panopticdiffusionmodels_b200 import distributed as D
    assert D.world() == (rank, world) and D.rank_seed(1234) == 1234 + rank
    calls = []

    def sample_fn(b):
        calls.append(b)
        k = len(calls) - 1
        z = torch.full((b, 4, 2, 2), float(100 * k + 10 * rank)) + torch.arange(b).float().view(b, 1, 1, 1)
        return z, -z[:, :1].repeat(1, 8, 1, 1)

    z, pm = D.sample_all(sample_fn, n_samples=10, mini_batch_size=3)   # global batches of 6: 6 + 4 (trimmed)
    q.put((rank, z[:, 0, 0, 0].tolist(), pm[:, 0, 0, 0].tolist(), calls))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_sampling():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, z, pm, calls in res:
        assert calls == [3, 3]                                   # each rank samples only its own shard
        # rank-major gather per global batch, trimmed to 10 samples; identical on every rank
        assert z == [0, 1, 2, 10, 11, 12, 100, 101, 102, 110]
        assert pm == [-v for v in z]


def test_single_process_passthrough():
    from panopticdiffusionmodels_b200 import distributed as D
    z, pm = D.sample_all(lambda b: (torch.ones(b, 4, 2, 2), None), n_samples=5, mini_batch_size=2)
    assert z.shape == (5, 4, 2, 2) and pm is None


def test_context_indices_follow_the_gather_order():
    """sample_t2i draws captions in contiguous per-rank blocks inside each global batch, so the rank-major all-gather puts the
    sample of caption i at output index i (the reference gathers the same way, utils.py:585-588)."""
    from panopticdiffusionmodels_b200.sample_t2i import context_indices
    mbs, n = 3, 2
    order = []
    for batch in range(2):                       # global batches, gathered rank-major
        for rank in range(n):
            order += context_indices(batch, mbs, rank, n, total=100)
    assert order == list(range(2 * mbs * n))
    assert context_indices(0, 4, 1, 2, total=6) == [4, 5, 0, 1]     # wraps around a short caption list
