"""Drift-balanced A/B of a library option on the training step: python tools/train_ab_balanced.py <option> v0 v1
Order v0 v1 v1 v0 v0 v1 v1 v0 (each 30 steps), so a monotonic clock / power drift cancels in the means."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from speaker_embedding_torch_b200 import GE2E, GE2E_Loss, _native
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
from speaker_embedding_torch_b200.Radam import RAdam
opt_name, v0, v1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = GE2E(default_hyper_parameters()).to(dev).train()
crit = GE2E_Loss().to(dev)
opt = RAdam(model.parameters(), lr=2e-3, eps=1e-6, max_grad_norm=1.0)
gen = torch.Generator(device=dev).manual_seed(1234)
mels = [bench.synth_mel(gen, 960, T, dev) for T in (160, 144, 176, 152, 168)]
def step(mel):
    opt.zero_grad(set_to_none=True)
    loss = crit(model(mel), 15)
    loss.backward()
    opt.step()
for v in (v0, v1):
    _native.set_option(opt_name, v)
    for _ in range(4):
        for m in mels: step(m)
torch.cuda.synchronize()
res = {v0: [], v1: []}
for v in (v0, v1, v1, v0, v0, v1, v1, v0):
    _native.set_option(opt_name, v)
    for m in mels[:2]: step(m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): step(mels[i % 5])
    e1.record(); torch.cuda.synchronize()
    res[v].append(e0.elapsed_time(e1) / 30)
    print("%s=%d: %.3f ms/step" % (opt_name, v, res[v][-1]), flush=True)
for v in (v0, v1):
    print("mean %s=%d: %.4f ms/step" % (opt_name, v, sum(res[v]) / len(res[v])))
_native.set_option(opt_name, v0)
