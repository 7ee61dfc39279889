"""N > 1 host path on CPU: world_size-2 gloo processes shard the environments and all-reduce the aggregate
infraction metrics (the only collective of the hot path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchdrivesim_b200 import distributed as D


def test_shard_ranges_cover_the_batch():
    for n, w in ((1024, 8), (10, 3), (2, 4), (0, 2)):
        ranges = [D.shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    D.init_process_group("gloo")
    torch.manual_seed(0)
    B, A = 10, 4
    coll, off = torch.rand(B, A), torch.rand(B, A) * (torch.rand(B, A) > 0.5)
    present = torch.rand(B, A) > 0.3
    idx = D.shard_indices(B, rank, world)
    local = D.infraction_metrics(coll[idx], off[idx], present[idx])
    D.allreduce_metrics(local)
    ref = D.infraction_metrics(coll, off, present)
    out[rank] = bool(torch.allclose(local, ref))
    dist.destroy_process_group()


def test_allreduce_of_metrics_equals_the_unsharded_metrics():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_single_process_is_a_noop():
    v = torch.ones(6, dtype=torch.float64)
    assert D.allreduce_metrics(v) is None and D.rank_world()[1] >= 1
    assert set(D.metrics_dict(v)) == set(D.METRIC_NAMES)
