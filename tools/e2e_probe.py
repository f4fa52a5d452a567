"""End-to-end loop variants: synchronous H2D on the compute stream vs Device_Prefetcher (debug aid)."""
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
from speaker_embedding_torch_b200.Prefetch import Device_Prefetcher  # noqa: E402
from speaker_embedding_torch_b200.Radam import RAdam  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = GE2E(default_hyper_parameters()).to(dev).train()
crit = GE2E_Loss().to(dev)
opt = RAdam(model.parameters(), lr=2e-3, eps=1e-6, max_grad_norm=1.0)
gen = torch.Generator(device=dev).manual_seed(1234)
rs = np.random.RandomState(0)
lengths = [180, 140] + [int(rs.randint(bench.T_MIN, bench.T_MAX + 1)) for _ in range(K)]


def step(mel):
    opt.zero_grad(set_to_none=True)
    loss = crit(model(mel), 15)
    loss.backward()
    opt.step()
    return loss


for T in lengths[:4]:
    step(bench.synth_mel(gen, 960, T, dev))
torch.cuda.synchronize()
host = [bench.synth_mel(gen, 960, T, dev).cpu().pin_memory() for T in lengths[2:]]
print("pinned:", host[0].is_pinned())

# raw H2D bandwidth
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for h in host:
    d = h.to(dev, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print("H2D alone: %.2f ms per batch (%.1f MB)" % (e0.elapsed_time(e1) / len(host), host[0].numel() * 4 / 1e6))


def run(name, it):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per = []
    it = iter(it)
    nx = []
    while True:
        tn = time.perf_counter()
        try:
            mel = next(it)
        except StopIteration:
            break
        t1 = time.perf_counter()
        nx.append((t1 - tn) * 1e3)
        step(mel).item()
        per.append((time.perf_counter() - t1) * 1e3)
    print("   next(): " + " ".join("%.1f" % p for p in nx))
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3
    print("%-28s %.2f ms/step   per-step: %s" % (name, tot / len(host), " ".join("%.1f" % p for p in per)))


run("sync H2D on compute stream", (h.to(dev, non_blocking=True) for h in host))
run("Device_Prefetcher", Device_Prefetcher(host, dev, reserve_bytes=960 * 80 * 180 * 4))
run("sync H2D again", (h.to(dev, non_blocking=True) for h in host))
run("Device_Prefetcher again", Device_Prefetcher(host, dev, reserve_bytes=960 * 80 * 180 * 4))
