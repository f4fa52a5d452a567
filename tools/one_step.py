"""Two training steps at the BASELINE shape (64 x 15 x T frames) -- the command profiled under ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from speaker_embedding_torch_b200 import GE2E, GE2E_Loss  # noqa: E402
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters  # noqa: E402
from speaker_embedding_torch_b200.Radam import RAdam  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T = int(sys.argv[2]) if len(sys.argv) > 2 else 160
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = GE2E(default_hyper_parameters()).to(dev).train(os.environ.get("SPK_EVAL") != "1")
crit = GE2E_Loss().to(dev)
opt = RAdam(model.parameters(), lr=2e-3, eps=1e-6, max_grad_norm=1.0)
gen = torch.Generator(device=dev).manual_seed(1234)
mel = bench.synth_mel(gen, 960, T, dev)
from speaker_embedding_torch_b200 import _native as N  # noqa: E402
for i in range(steps):
    if i == steps - 1 and os.environ.get("SPK_PROF") == "1":
        torch.cuda.synchronize()
        N.prof_enable(True)
    opt.zero_grad(set_to_none=True)
    loss = crit(model(mel), 15)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
if os.environ.get("SPK_PROF") == "1":
    rep = N.prof_report()
    N.prof_enable(False)
    tot = sum(v["ms"] for v in rep.values())
    print("profiled step %.3f ms" % tot)
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:int(os.environ.get("SPK_PROF_TOP", "24"))]:
        print("  %-28s %7.4f ms  x%d" % (k, v["ms"], v["launches"]))
