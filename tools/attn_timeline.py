"""Phase timeline of one CTA of the fused attention backward kernel (spk_set_debug_buffer): prints, per (i, j) unit,
the SM-clock deltas between the phase boundaries.  Usage: python tools/attn_timeline.py [T]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from speaker_embedding_torch_b200 import GE2E, GE2E_Loss, _native  # noqa: E402
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 160
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = GE2E(default_hyper_parameters()).to(dev).train()
crit = GE2E_Loss().to(dev)
gen = torch.Generator(device=dev).manual_seed(1234)
mel = bench.synth_mel(gen, 960, T, dev)
for _ in range(2):
    model.zero_grad(set_to_none=True)
    crit(model(mel), 15).backward()
torch.cuda.synchronize()
units = 32
buf = torch.zeros(units * 8, dtype=torch.int64, device=dev)
_native.lib().spk_set_debug_buffer(ctypes.c_void_p(buf.data_ptr()), buf.numel() * 8)
model.zero_grad(set_to_none=True)
crit(model(mel), 15).backward()
torch.cuda.synchronize()
_native.lib().spk_set_debug_buffer(None, 0)
t = buf.cpu().view(units, 8)          # the LAST launch that ran (layer 0's backward) wrote last
base = int(t[0, 0])
names = ["ops_ready", "mma1_issued", "bar_p_seen", "mma2_issued", "scores_seen", "chunks_done", "tile_seen", "drained"]
print("unit  " + "  ".join("%12s" % n for n in names))
for u in range(units):
    print("%4d  " % u + "  ".join("%12d" % (int(v) - base if int(v) else -1) for v in t[u]))
