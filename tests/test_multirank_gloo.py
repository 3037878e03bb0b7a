"""N>1 host logic of bench.py on CPU: two gloo ranks, the same reduction bench.py uses under
torchrun (MAX over ranks for times, SUM for counts), independent shards per rank, and the
reference arm's "rank 0 only" rule.  No GPU, no CUDA library calls."""
import importlib.util
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    sys.path.insert(0, ROOT)
    spec.loader.exec_module(m)
    return m


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = _bench()
    # rank r: wall = 1+r s, e2e = 2-r s, device = 0.5 s; counts 10+r bursts, 7 frames, 100 launches
    times, counts = b.reduce_over_ranks(torch, dist, torch.device("cpu"),
                                        [1.0 + rank, 2.0 - rank, 0.5], [10 + rank, 7, 100])
    out[rank] = (times, counts, b.shard_seed(world, rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_reduce_over_two_gloo_ranks():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for r in (0, 1):
        times, counts, _ = res[r]
        assert times == [2.0, 2.0, 0.5]          # max over ranks
        assert counts == [21, 14, 200]           # sum over ranks
    assert res[0][2] != res[1][2]                # ranks decode different recordings


def test_single_process_reduce_is_identity():
    b = _bench()
    t, c = b.reduce_over_ranks(torch, None, torch.device("cpu"), [0.25, 1.5], [3, 4])
    assert t == [0.25, 1.5] and c == [3, 4]
    assert b.shard_seed(1, 0) == 2
