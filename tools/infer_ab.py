"""A/B of a library option on the inference path: python tools/infer_ab.py <option>   (d-vectors identical? time per batch)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from speaker_embedding_torch_b200 import GE2E, _native
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
opt = sys.argv[1]
dev = torch.device('cuda', 0)
torch.manual_seed(0)
m = GE2E(default_hyper_parameters()).to(dev).eval()
gen = torch.Generator(device=dev).manual_seed(1)
for B, T in ((960, 160), (4000, 64), (960, 100), (7, 33), (960, 200)):
    mel = bench.synth_mel(gen, B, T, dev)
    outs = {}
    for on in (0, 1):
        _native.set_option(opt, on)
        with torch.no_grad():
            for _ in range(3): d = m(mel)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): d = m(mel)
            e1.record(); torch.cuda.synchronize()
        outs[on] = d
        print('%s=%d  B %d T %d  ms per batch %.3f' % (opt, on, B, T, e0.elapsed_time(e1) / 10))
    a, b = outs[0], outs[1]
    print('   min cos between variants %.7f  max abs diff %.2e' % (float((a * b).sum(1).min()), float((a - b).abs().max())))
_native.set_option(opt, 1)
