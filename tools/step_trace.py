"""Per-step device time of the bench's training loop with its variable frame lengths (allocator / clock hiccups)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from speaker_embedding_torch_b200 import GE2E, GE2E_Loss  # noqa: E402
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters  # noqa: E402
from speaker_embedding_torch_b200.Radam import RAdam  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = GE2E(default_hyper_parameters()).to(dev).train()
crit = GE2E_Loss().to(dev)
opt = RAdam(model.parameters(), lr=2e-3, eps=1e-6, max_grad_norm=1.0)
gen = torch.Generator(device=dev).manual_seed(1234)
rs = np.random.RandomState(0)
lengths = [int(rs.randint(bench.T_MIN, bench.T_MAX + 1)) for _ in range(steps)]
mels = [bench.synth_mel(gen, 960, T, dev) for T in lengths]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
torch.cuda.synchronize()
t0 = time.perf_counter()
cpu = []
ev[0].record()
for i in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = crit(model(mels[i]), 15)
    loss.backward()
    opt.step()
    ev[i + 1].record()
    cpu.append(time.perf_counter() - t0)
torch.cuda.synchronize()
for i in range(steps):
    print("step %2d T=%3d gpu %.2f ms  cpu_issue_done %.1f ms  reserved %.1f GB" % (
        i, lengths[i], ev[i].elapsed_time(ev[i + 1]), cpu[i] * 1e3, torch.cuda.memory_reserved() / 2**30))
