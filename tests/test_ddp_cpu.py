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
    # stage-sliced exchange (the overlap path) on host tensors: same result as the single collective
    ex = ddp.OverlappedGradExchange.__new__(ddp.OverlappedGradExchange)
    ex.group, ex.world, ex.backend = None, world, "gloo"
    ex.stages = [(0, [(4, 6)]), (1, [(2, 4)]), (2, [(0, 1), (1, 2)])]
    flat2 = torch.arange(6, dtype=torch.float32) * (rank + 1)
    ex.after_backward(flat2)
    q.put((rank, ok_bcast, bool(torch.allclose(flat, expect)) and bool(torch.allclose(flat2, expect))))
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


def test_backward_stage_slices_cover_the_flat_gradient_buffer_once():
    """univtg_backward_stages (host function of the C-ABI library) x the plugin's flat layout: every float of every
    parameter view belongs to exactly one stage slice."""
    from univtg_b200 import build_model, ddp, synth

    for name in ("tiny", "cfg2"):
        cfg = synth.CONFIGS[name]
        model, _ = build_model(synth.reference_args(cfg, device="cpu"))
        stages = ddp.grad_stage_slices(model)
        assert len(stages) == cfg["enc_layers"] + 3
        offs = model._grad_offsets()
        cover = torch.zeros(offs[-1], dtype=torch.int32)
        for _, sl in stages:
            for lo, hi in sl:
                assert 0 <= lo < hi <= offs[-1] and lo % 4 == 0
                cover[lo:hi] += 1
        assert int(cover.min()) == 1 and int(cover.max()) == 1
        # completion order: heads first, encoder layers from the last to the first, projectors last
        firsts = [sl[0][0] for _, sl in stages]
        assert firsts[0] > firsts[1] > firsts[-2] > firsts[-1] == 0
