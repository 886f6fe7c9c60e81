"""world_size-2 gloo test (CPU) of the N>1 host path: shard bounds, per-shard parameters keyed by the
global sample index, and the optional final gather.  No collective sits on the data path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ultrasoundaugmentation_b200 as us
    from ultrasoundaugmentation_b200 import sharding
    ch = us.PhysicsChain.full(seed=21)
    p = sharding.shard_params(ch, total, rank, world, (64, 64))
    lo, hi = sharding.shard_bounds(total, rank, world)
    # stand-in for the per-rank augmented shard: rows tagged with their global index and drawn d
    local = torch.stack([torch.arange(lo, hi, dtype=torch.float32), p["d"]], dim=1)
    full = sharding.gather_batch(local, total)
    ref = ch.get_params(total, image_size=(64, 64))["d"]
    ok = full.shape == (total, 2) and torch.equal(full[:, 0], torch.arange(total, dtype=torch.float32)) \
        and torch.equal(full[:, 1], ref)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(float(t.item()))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_ragged():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, ret)) for r in range(2)]   # 7 = ragged split 4 + 3
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1.0


def test_gather_is_identity_without_process_group():
    from ultrasoundaugmentation_b200 import sharding
    x = torch.arange(6.0).reshape(3, 2)
    assert sharding.gather_batch(x, 3) is x
