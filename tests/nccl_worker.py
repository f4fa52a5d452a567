"""Worker of tests/test_gpu_nccl.py: one process per GPU (torchrun), NCCL, the real encoder.

Checks, on every rank:
  1. apply_gradient_allreduce broadcasts rank 0's weights (ranks start from different states);
  2. after one backward (eval mode, so the oracle can follow) the gradient arena holds the MEAN over ranks of the
     per-rank oracle gradients (each rank owns its own speakers; the loss is rank-local);
  3. reduce_tensor(loss) is the mean of the rank losses (answered from the arena, no second collective);
  4. after 3 train-mode steps (dropout on, different masks per rank) with the fused RAdam the weights are still
     bit-identical on all ranks.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ge2e_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from speaker_embedding_torch_b200 import GE2E, GE2E_Loss
    from speaker_embedding_torch_b200 import distributed as D
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    from speaker_embedding_torch_b200.Radam import RAdam

    nspk, utt, T = 3, 2, 40
    state0 = synth.make_state(71)                       # rank 0's weights
    mine = synth.make_state(71 + rank)                  # every other rank starts somewhere else
    model = GE2E(default_hyper_parameters())
    model.load_state_dict({k: torch.as_tensor(v) for k, v in mine.items()}, strict=True)
    model = model.to(dev)
    assert D.apply_gradient_allreduce(model) is model
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), torch.as_tensor(state0[k])), "broadcast: %s differs from rank 0" % k

    crit = GE2E_Loss().to(dev)
    model.eval()
    mels = [synth.make_mel(900 + r, nspk * utt, T) for r in range(world)]
    loss = crit(model(torch.as_tensor(mels[rank]).to(dev)), utt)
    loss.backward()
    arena = model._arena
    lo, hi = arena.data_ptr(), arena.data_ptr() + arena.numel() * 4
    assert all(lo <= p.grad.data_ptr() < hi for p in model.parameters()), "gradients do not alias the arena"
    refs = [O.train_step_grads(state0, mels[r], utt) for r in range(world)]
    num = den = 0.0
    for name, p in model.named_parameters():
        want = sum(refs[r][2][name] for r in range(world)) / world
        got = p.grad.detach().cpu().numpy().astype(np.float64)
        num += ((got - want) ** 2).sum()
        den += (want ** 2).sum()
    grad_rel = (num / den) ** 0.5
    assert grad_rel <= 1e-3, grad_rel
    mean_loss = D.reduce_tensor(loss.data, world).item()
    want_loss = sum(refs[r][0] for r in range(world)) / world
    assert abs(mean_loss - want_loss) <= 1e-3 * abs(want_loss), (mean_loss, want_loss)
    assert abs(loss.item() - refs[rank][0]) <= 1e-3 * abs(refs[rank][0])

    model.train()
    torch.manual_seed(1000 + rank)                      # different dropout masks per rank
    opt = RAdam(model.parameters(), lr=2e-3, eps=1e-6, max_grad_norm=1.0)
    for step in range(3):
        opt.zero_grad()
        crit(model(torch.as_tensor(synth.make_mel(950 + 10 * step + rank, nspk * utt, T)).to(dev)), utt).backward()
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], t) for t in gathered), "weights diverged across ranks"
    assert not torch.equal(flat.cpu(), torch.cat([torch.as_tensor(state0[n]).reshape(-1)
                                                  for n, _ in model.named_parameters()])), "weights did not move"
    dist.barrier()
    if rank == 0:
        print("NCCL_WORKER_OK " + json.dumps({"world": world, "grad_rel": grad_rel, "mean_loss": mean_loss}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
