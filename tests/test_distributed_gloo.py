"""world_size-2 gloo test of the data-parallel plumbing (distributed.py) on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _ArenaFn(torch.autograd.Function):
    """Mimics the encoder's backward: all parameter grads are views of one flat arena."""

    @staticmethod
    def forward(ctx, x, module, *params):
        ctx.module = module
        ctx.save_for_backward(x, *params)
        return (x @ params[0]).sum() + (params[1] ** 2).sum() * x.mean()

    @staticmethod
    def backward(ctx, g):
        x, w, b = ctx.saved_tensors
        arena = torch.zeros(w.numel() + b.numel() + 4)        # + the spare slots behind the gradients (loss piggyback)
        ctx.module._arena = arena
        ctx.module._arena_numel = w.numel() + b.numel()
        gw = arena[:w.numel()].view_as(w)
        gb = arena[w.numel():w.numel() + b.numel()].view_as(b)
        gw.copy_(x.sum(0).unsqueeze(1).expand_as(w) * g)
        gb.copy_(2 * b * x.mean() * g)
        return None, None, gw, gb


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.randn(4, 3))
        self.b = torch.nn.Parameter(torch.randn(5))
        self.register_buffer("buf", torch.randn(2))
        self._arena = None

    def forward(self, x):
        return _ArenaFn.apply(x, self, self.w, self.b)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from speaker_embedding_torch_b200.distributed import apply_gradient_allreduce, reduce_tensor
    torch.manual_seed(100 + rank)          # different init per rank: broadcast must fix it
    toy = _Toy()
    same = apply_gradient_allreduce(toy)
    assert same is toy
    w_all = [torch.zeros_like(toy.w) for _ in range(world)]
    dist.all_gather(w_all, toy.w.detach())
    buf_all = [torch.zeros_like(toy.buf) for _ in range(world)]
    dist.all_gather(buf_all, toy.buf)
    assert all(torch.equal(w_all[0], t) for t in w_all) and all(torch.equal(buf_all[0], t) for t in buf_all)

    # arena path; the step's loss rides behind the gradients (what GE2E_Loss.forward registers)
    from speaker_embedding_torch_b200 import distributed as D
    from speaker_embedding_torch_b200.Modules import LAST_TRAINING_LOSS
    x = torch.randn(6, 4, generator=torch.Generator().manual_seed(7 + rank))
    loss = toy(x)
    LAST_TRAINING_LOSS[loss.device] = loss.detach()
    loss.backward()
    losses = [torch.zeros(()) for _ in range(world)]
    dist.all_gather(losses, loss.detach())
    assert loss.data_ptr() in D._PIGGYBACK
    mean_loss = reduce_tensor(loss.data, world)                    # no collective: answered from the arena
    torch.testing.assert_close(mean_loss, sum(losses) / world)
    assert not D._PIGGYBACK
    torch.testing.assert_close(reduce_tensor(loss.data, world), sum(losses) / world)     # second call: real collective
    assert toy.w.grad.data_ptr() == toy._arena.data_ptr()          # grads alias the arena
    local_gw = x.sum(0).unsqueeze(1).expand(4, 3)
    gw_all = [torch.zeros(4, 3) for _ in range(world)]
    dist.all_gather(gw_all, local_gw.contiguous())
    torch.testing.assert_close(toy.w.grad, sum(gw_all) / world)
    local_gb = 2 * toy.b.detach() * x.mean()
    gb_all = [torch.zeros(5) for _ in range(world)]
    dist.all_gather(gb_all, local_gb)
    torch.testing.assert_close(toy.b.grad, sum(gb_all) / world)

    # generic path (plain autograd module)
    lin = torch.nn.Linear(3, 2)
    apply_gradient_allreduce(lin)
    xin = torch.randn(5, 3, generator=torch.Generator().manual_seed(20 + rank))
    lin(xin).sum().backward()
    g_all = [torch.zeros(2, 3) for _ in range(world)]
    dist.all_gather(g_all, xin.sum(0).unsqueeze(0).expand(2, 3).contiguous())
    torch.testing.assert_close(lin.weight.grad, sum(g_all) / world)

    # a second backward without a forward must not reduce again (needs_reduction protocol)
    r = reduce_tensor(torch.tensor(float(rank + 1)), world)
    assert abs(r.item() - (sum(range(1, world + 1)) / world)) < 1e-6
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(out.get(timeout=5) for _ in range(2)) == [0, 1]
