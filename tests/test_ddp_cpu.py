"""Host-side logic of the N>1 path on CPU: two gloo ranks, the flat-gradient hook averages once, parameters broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from univtg_b200 import ddp

    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.full((5,), float(rank + 1)))

    m = Dummy()
    ddp.broadcast_parameters(m)
    ok_bcast = bool((m.w.data == 1.0).all())
    ddp.attach_flat_allreduce(m)
    flat = torch.arange(6, dtype=torch.float32) * (rank + 1)  # rank r holds (r+1) * [0..5]
    m._flat_grad_hook(flat)
    expect = torch.arange(6, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    q.put((rank, ok_bcast, bool(torch.allclose(flat, expect))))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "broadcast_parameters did not replicate rank 0"
    assert all(r[2] for r in res), "flat all-reduce did not average"
