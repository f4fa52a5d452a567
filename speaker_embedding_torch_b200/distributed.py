"""Data-parallel plumbing: drop-in for the reference's ``distributed.py`` on one NVSwitch box.

Same three entry points as /root/reference/distributed.py:
  * ``init_distributed(rank, num_gpus, dist_backend)``        (distributed.py:28-36)
  * ``apply_gradient_allreduce(module) -> module``            (distributed.py:73-125)
  * ``reduce_tensor(tensor, num_gpus)``                        (distributed.py:22-26)

What changes underneath:
  * the encoder's backward writes every parameter gradient into ONE flat fp32 arena
    (``GE2E._grad_arena``); ``param.grad`` tensors are views of it, so the per-step collective is a
    single in-place ``all_reduce(AVG)`` over NCCL / NVLink with no flatten ``cat`` and no copy-back;
  * initial weights are broadcast as one flat message instead of 44;
  * like the reference, the loss is rank-local (each rank owns its speakers): only
    parameter gradients are exchanged, there is no embedding all-gather.
One process per GPU (torchrun / ``python -m torch.distributed.run``), env:// rendezvous.
"""
import os

import torch
import torch.distributed as dist
from torch.autograd import Variable


# data_ptr of the loss whose rank-mean already sits behind the gradients in the arena -> that arena element
_PIGGYBACK = {}


def reduce_tensor(tensor, num_gpus):
    """Mean of ``tensor`` over ranks (logging only).  The training loss of the step whose gradients were just averaged
    needs no collective of its own: its mean came back with the gradient arena."""
    hit = _PIGGYBACK.pop(tensor.data_ptr(), None) if tensor.numel() == 1 else None
    if hit is not None and hit.device == tensor.device:
        return hit.clone().view_as(tensor)
    rt = tensor.clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= num_gpus
    return rt


def init_distributed(rank, num_gpus, dist_backend):
    if not torch.cuda.is_available():
        raise AssertionError("Distributed mode requires CUDA.")
    print("> initializing distributed for rank {} out of {}".format(rank, num_gpus))
    local = int(os.environ.get("LOCAL_RANK", rank % torch.cuda.device_count()))
    torch.cuda.set_device(local)
    dist.init_process_group(backend=dist_backend or "nccl", device_id=torch.device("cuda", local))


def _flat_broadcast(tensors, src=0):
    """Broadcast a list of same-device tensors as one message."""
    groups = {}
    for t in tensors:
        groups.setdefault((t.dtype, t.device), []).append(t)
    for (_, _), ts in groups.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, src)
        off = 0
        for t in ts:
            n = t.numel()
            t.detach().copy_(flat[off:off + n].view_as(t))
            off += n


def _average_inplace(buf, world):
    backend = dist.get_backend()
    if backend == "nccl":
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    else:                       # gloo (CPU tests) has no AVG
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        buf /= world


def allreduce_gradients(module):
    """Average ``param.grad`` over ranks.  One in-place collective when the grads alias the flat arena."""
    world = dist.get_world_size()
    params = [p for p in module.parameters() if p.requires_grad and p.grad is not None]
    if not params:
        return 0
    arena = getattr(module, "_arena", None)
    if arena is not None and arena.is_cuda == params[0].grad.is_cuda:
        lo, hi = arena.data_ptr(), arena.data_ptr() + arena.numel() * arena.element_size()
        if all(lo <= p.grad.data_ptr() < hi for p in params):
            from .Modules import LAST_TRAINING_LOSS
            loss = LAST_TRAINING_LOSS.pop(arena.device, None)
            numel = getattr(module, "_arena_numel", None)
            _PIGGYBACK.clear()
            if loss is not None and numel is not None and arena.numel() > numel:
                arena[numel].copy_(loss.float())               # the step's loss rides behind the gradients
                _average_inplace(arena, world)
                _PIGGYBACK[loss.data_ptr()] = arena[numel]
            else:
                _average_inplace(arena, world)
            return 1
    # generic path (gradients produced by plain autograd, e.g. CPU tests): one flat bucket per dtype
    buckets = {}
    for p in params:
        buckets.setdefault(p.grad.dtype, []).append(p.grad)
    for grads in buckets.values():
        flat = torch.cat([g.reshape(-1) for g in grads])
        _average_inplace(flat, world)
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    return len(buckets)


def apply_gradient_allreduce(module):
    """Broadcast rank 0's state, then average gradients at the end of every backward.
    Returns the same module object (no wrapper), like the reference."""
    _flat_broadcast([t for t in module.state_dict().values() if torch.is_tensor(t)], 0)
    module.needs_reduction = False

    def allreduce_params():
        if module.needs_reduction:
            module.needs_reduction = False
            allreduce_gradients(module)

    def allreduce_hook(*unused):
        Variable._execution_engine.queue_callback(allreduce_params)

    for param in list(module.parameters()):
        if param.requires_grad:
            param.register_hook(allreduce_hook)

    def set_needs_reduction(self, input, output):
        self.needs_reduction = True

    module.register_forward_hook(set_needs_reduction)
    return module
