"""world_size-2 gloo test of the N>1 host path: contiguous batch sharding + the single all-gather of token ids."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vcoder_b200.dist import all_gather_ids, shard_batch, split_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = torch.arange(n_total * 3, dtype=torch.int64).view(n_total, 3)
        pix = torch.arange(n_total, dtype=torch.float32).view(n_total, 1)
        my_ids, my_pix, none = shard_batch(world, rank, ids, pix, None)
        assert none is None
        b, e = split_range(n_total, world, rank)
        assert my_ids.shape[0] == e - b == my_pix.shape[0]
        # "generate": new token = f(prompt) so the gathered result is checkable
        new = (my_ids.sum(dim=1, keepdim=True) + torch.arange(4)[None]).to(torch.int32)
        full = all_gather_ids(new, n_total)
        want = (ids.sum(dim=1, keepdim=True) + torch.arange(4)[None]).to(torch.int32)
        ret[rank] = bool(torch.equal(full, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 5])
def test_shard_and_allgather_world2(n_total):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_split_range_covers_everything():
    for n in range(1, 40):
        for w in (1, 2, 3, 4, 8):
            spans = [split_range(n, w, r) for r in range(w)]
            got = [i for b, e in spans for i in range(b, e)]
            assert got == list(range(n))
