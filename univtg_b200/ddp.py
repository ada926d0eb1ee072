"""Data-parallel training of the UniVTG path: batches shard by video-query pair, parameters stay replicated, and the only
cross-GPU exchange is ONE all-reduce (average) of the flat gradient buffer per step (reference: DDP's bucketed gradient
all-reduce, main/train_vlp_ddp.py:272-275).

Usage (one process per GPU, torchrun):
    dist.init_process_group("nccl")
    model, criterion = build_model(args); model.to(device)
    ddp.broadcast_parameters(model)              # what the DDP constructor does (rank 0 -> all)
    ddp.attach_flat_allreduce(model)             # one NCCL all-reduce issued at the end of the fused backward
    ... the usual loop: outputs = model(**inputs); loss = ...; loss.backward(); optimizer.step()

The reference's own script wraps the model in torch DistributedDataParallel(find_unused_parameters=True); that also works
unchanged with this model (gradients reach param.grad through autograd), with DDP's 25 MB buckets instead of one buffer.
"""
import torch
import torch.distributed as dist


def broadcast_parameters(model, src=0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers."""
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def make_flat_allreduce_hook(group=None):
    """Returns hook(flat): in-place average of `flat` over the process group with a single collective."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)

    def hook(flat):
        if world == 1:
            return
        if backend == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)  # NVLink/NVSwitch; NVLS in-switch reduction when available
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(world)

    return hook


def attach_flat_allreduce(model, group=None):
    """Install the single-collective gradient exchange on a univtg_b200 model (runs inside its fused backward)."""
    model._flat_grad_hook = make_flat_allreduce_hook(group)
    model.direct_grad = True  # gradients are handed to param.grad as views of the flat buffer (no autograd accumulation copies)
    return model


def detach_flat_allreduce(model):
    model._flat_grad_hook = None
    model.direct_grad = False
    return model
