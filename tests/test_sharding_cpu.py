"""Host-side multi-GPU logic on CPU: contiguous partitions, the packed gradient buffer, and a real world_size-2
all-reduce over gloo (the GPU path uses the same call over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vgpmp_b200.utils import sharding as S


def test_shard_range_partitions_everything():
    for n in (1, 7, 55, 275, 1024, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [S.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.shard_range(10, 4, 4)
    assert S.partition_problems(list(range(55)), 1, 8) == list(range(7, 14))


def test_packed_views_alias_one_buffer():
    lay = S.packed_layout(3, 5, 7)
    n = lay["_total"][0]
    assert n == 3 * (5 * 7 + 7 * 25 + 7 + 7 + 1)
    flat = torch.zeros(n, dtype=torch.float64)
    v = S.packed_views(flat, 3, 5, 7)
    assert v["d_q_sqrt"].shape == (3, 7, 5, 5) and v["elbo"].shape == (3,)
    v["d_variances"].fill_(2.0)
    v["elbo"].fill_(-1.0)
    assert flat.sum().item() == 2.0 * 21 - 3.0
    with pytest.raises(ValueError):
        S.packed_views(flat[:-1], 3, 5, 7)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Bp, M, D = 2, 4, 6
        flat = torch.zeros(S.packed_layout(Bp, M, D)["_total"][0], dtype=torch.float64)
        v = S.packed_views(flat, Bp, M, D)
        # every rank contributes its sample-slice partial sums; KL-type terms carry weight 1/world
        v["d_q_mu"].fill_(float(rank + 1))
        v["elbo"].copy_(torch.tensor([10.0 * (rank + 1), -3.0 / world]))
        S.allreduce_packed(flat)
        lo, hi = S.shard_range(11, rank, world)
        q.put((rank, v["d_q_mu"][0, 0, 0].item(), v["elbo"].tolist(), lo, hi))
    finally:
        dist.destroy_process_group()


def test_allreduce_packed_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, elbo, lo, hi in res:
        assert g == 3.0                               # 1 + 2
        assert np.allclose(elbo, [30.0, -3.0])        # likelihood parts add, the split KL term adds back to 1x
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 6, 6, 11)
