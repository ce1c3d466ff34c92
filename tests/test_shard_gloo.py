"""world_size-2 gloo test of the N>1 host path: weight broadcast, contiguous clip shards, pose
gather in clip order.  The per-rank compute is a deterministic stand-in (row-wise function of the
inputs) because the CUDA engine needs a GPU; the sharding / collective plumbing is the unit under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from amuse_b200.shard import broadcast_state_dict, gather_clips, shard_range
    g = torch.Generator().manual_seed(100 + rank)          # different on every rank before the broadcast
    sd = {"w": torch.randn(4, 4, generator=g), "b": torch.randn(4, generator=g)}
    broadcast_state_dict(sd, src=0)
    full = torch.arange(n_global * 4, dtype=torch.float32).view(n_global, 4)     # rank-0-style global inputs
    a, b = shard_range(n_global, rank, world)
    local = full[a:b] @ sd["w"] + sd["b"]                   # per-clip independent "compute"
    out = gather_clips(local.view(b - a, 2, 2), n_global, dst=0)
    if rank == 0:
        ret["out"] = out
        ret["w"] = sd["w"]
        ret["b"] = sd["b"]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [7, 1])               # ragged: 4 + 3 clips; 1 + 0 clips (an empty shard)
def test_broadcast_shard_gather_two_ranks(n_global):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_global, ret), nprocs=world, join=True)
    full = torch.arange(n_global * 4, dtype=torch.float32).view(n_global, 4)
    expect = (full @ ret["w"] + ret["b"]).view(n_global, 2, 2)
    assert torch.equal(ret["out"], expect)                  # same result as the 1-rank computation
