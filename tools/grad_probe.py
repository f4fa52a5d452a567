"""Per-tensor gradient errors against the fp64 oracle for a few small cases, under precision / fused-attention /
gradient-scale variations.  Usage: python tools/grad_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ge2e_oracle as O, synth  # noqa: E402
from speaker_embedding_torch_b200 import GE2E, GE2E_Loss, _native  # noqa: E402
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters  # noqa: E402


def run(nspk, utt, T, precision, fused, gs, seed=37):
    _native.set_option("fused_training_attention", fused)
    _native.set_option("grad_scale_log2", gs)
    state = synth.make_state(seed)
    m = GE2E(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=True)
    m = m.cuda().eval()
    m.train_precision = precision
    crit = GE2E_Loss().cuda()
    mel = synth.make_mel(700 + T, nspk * utt, T)
    loss = crit(m(torch.as_tensor(mel).cuda()), utt)
    loss.backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().cpu().numpy().astype(np.float64) for n, p in m.named_parameters()}, state, mel


for (nspk, utt, T) in ((2, 2, 1024), (3, 2, 16), (4, 3, 160)):
    ref = None
    for precision, fused, gs in ((2, 1, 8), (2, 0, 8), (3, 1, 8), (2, 1, 12), (2, 0, 12), (2, 1, 14)):
        g, state, mel = run(nspk, utt, T, precision, fused, gs)
        if ref is None:
            ref = O.train_step_grads(state, mel, utt)[2]
        num = sum(((g[n] - ref[n]) ** 2).sum() for n in g)
        den = sum((ref[n] ** 2).sum() for n in g)
        per = sorted(((np.linalg.norm(g[n] - ref[n]) / max(np.linalg.norm(ref[n]), 1e-30), n) for n in g), reverse=True)[:3]
        print("%dx%dx%d P=%d fused=%d gs=2^%d: grad_rel %.3e | worst %s" % (
            nspk, utt, T, precision, fused, gs, (num / den) ** 0.5,
            ", ".join("%s %.1e" % (n.replace("transformer.layers.", "L"), r) for r, n in per)), flush=True)
_native.set_option("fused_training_attention", 1)
_native.set_option("grad_scale_log2", 8)
