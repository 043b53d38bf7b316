"""Host-side logic of the multi-GPU path on CPU: contiguous query shards, one broadcast per index
buffer, results gathered in rank order.  world_size 2, gloo backend, 127.0.0.1 rendezvous."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diskann_b200 import sharding


def test_partition_matches_partition_iter():
    # diskann-benchmark-core/src/search/api.rs:410-419: contiguous, first n % parts one longer
    assert sharding.partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sharding.partition(10000, 8)[0] == (0, 1250) and sharding.partition(10000, 8)[-1] == (8750, 10000)
    assert sharding.partition(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sharding.partition(0, 2) == [(0, 0), (0, 0)]
    for n in (0, 1, 7, 64, 1001):
        for parts in (1, 2, 3, 8):
            r = sharding.partition(n, parts)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    with pytest.raises(ValueError):
        sharding.partition(5, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        n, d, stride, nq, k = 50, 8, 5, 11, 3
        truth_vec = rng.normal(size=(n, d)).astype(np.float32)
        truth_adj = rng.integers(0, n, (n, stride)).astype(np.uint32)
        # only rank 0 holds the index; the others pass zero-filled buffers of the same shape
        arrays = {"vectors": truth_vec if rank == 0 else np.zeros_like(truth_vec),
                  "adj": truth_adj if rank == 0 else np.zeros_like(truth_adj)}
        got = sharding.broadcast_arrays(arrays, src=0)
        assert np.array_equal(got["vectors"].numpy(), truth_vec)
        assert np.array_equal(got["adj"].numpy().view(np.uint32), truth_adj)
        # each rank "searches" its contiguous shard; results come back in query order on rank 0
        lo, hi = sharding.shard_of(nq, rank, world)
        local_ids = np.arange(lo, hi, dtype=np.uint32)[:, None] * 10 + np.arange(k, dtype=np.uint32)[None, :]
        local_d = local_ids.astype(np.float32) / 2
        ids, dists = sharding.gather_results(local_ids, local_d, nq, k, dst=0)
        if rank == 0:
            want = np.arange(nq, dtype=np.uint32)[:, None] * 10 + np.arange(k, dtype=np.uint32)[None, :]
            assert np.array_equal(ids, want) and np.array_equal(dists, want.astype(np.float32) / 2)
        else:
            assert ids is None
        # device-side timing is reported as the max over ranks
        assert sharding.max_over_ranks(1.0 + rank) == float(world)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_broadcast_shard_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: "ok", 1: "ok"}
