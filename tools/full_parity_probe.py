"""Where does the gradient error at the benchmark size come from?  On the GPU box: the staged, unmodified reference
(baseline/_ref) in fp64 and in fp32 (TF32 off) on the BASELINE batch (64 x 15 x T), next to this framework's step on
the same weights and input.  Prints the relative L2 error of every comparison, globally and for the worst tensors:

    ref32 vs ref64   what an fp32 implementation of the same graph achieves (ReLU gates that flip, fp32 sums)
    ours  vs ref64   the parity figure the tests assert
    ours  vs ref32

Usage: python tools/full_parity_probe.py [T] [state_seed] [mel_seed]      (diagnostic, not part of the product path)
"""
import os
import sys
import warnings

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402


def load_reference():
    warnings.filterwarnings("ignore")
    import importlib.util

    def mod(name):
        spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    modules, args = mod("Modules"), mod("Arg_Parser")
    hp = args.Recursive_Parse(yaml.load(open(os.path.join(REF, "Hyper_Parameters.yaml")), Loader=yaml.Loader))
    return modules.GE2E, modules.GE2E_Loss, hp


def reference_grads(GE2E, GE2E_Loss, hp, state, mel, utt, dtype):
    m = GE2E(hp)
    m.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=True)
    m = m.to(dtype).cuda().eval()
    crit = GE2E_Loss().to(dtype).cuda()
    d = m(torch.as_tensor(mel).to(dtype).cuda())
    loss = crit(d, utt)
    loss.backward()
    torch.cuda.synchronize()
    g = {n: p.grad.detach().double().cpu().numpy() for n, p in m.named_parameters()}
    return float(loss), d.detach().double().cpu().numpy(), g


def ours_grads(state, mel, utt, precision):
    from speaker_embedding_torch_b200 import GE2E, GE2E_Loss
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    m = GE2E(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=True)
    m = m.cuda().eval()
    m.train_precision = precision
    crit = GE2E_Loss().cuda()
    d = m(torch.as_tensor(mel).cuda())
    loss = crit(d, utt)
    loss.backward()
    torch.cuda.synchronize()
    g = {n: p.grad.detach().double().cpu().numpy() for n, p in m.named_parameters()}
    return float(loss), d.detach().double().cpu().numpy(), g


def compare(tag, a, b):
    num = sum(((a[n] - b[n]) ** 2).sum() for n in a)
    den = sum((b[n] ** 2).sum() for n in a)
    per = sorted(((np.linalg.norm(a[n] - b[n]) / max(np.linalg.norm(b[n]), 1e-30), n) for n in a), reverse=True)
    print("%-16s grad_rel %.3e | %s" % (tag, (num / den) ** 0.5, ", ".join(
        "%s %.1e" % (n.replace("transformer.layers.", "L"), r) for r, n in per[:6])), flush=True)
    if os.environ.get("SPK_PROBE_ALL"):
        for r, n in sorted(per, key=lambda t: list(a).index(t[1])):
            print("      %-48s %.2e" % (n, r))
    return {n: r for r, n in per}


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    ss = int(sys.argv[2]) if len(sys.argv) > 2 else 61
    ms = int(sys.argv[3]) if len(sys.argv) > 3 else 601
    nspk, utt = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (64, 15)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    GE2E, GE2E_Loss, hp = load_reference()
    state = synth.make_state(ss)
    mel = synth.make_mel(ms, nspk * utt, T)
    l64, d64, g64 = reference_grads(GE2E, GE2E_Loss, hp, state, mel, utt, torch.float64)
    l32, d32, g32 = reference_grads(GE2E, GE2E_Loss, hp, state, mel, utt, torch.float32)
    print("%d x %d x %d   loss ref64 %.9f ref32 %.9f" % (nspk, utt, T, l64, l32))
    compare("ref32 vs ref64", g32, g64)
    from speaker_embedding_torch_b200 import _native
    for precision, fused, gs, prune in ((2, 1, 12, 1),):
        _native.set_option("fused_training_attention", fused)
        _native.set_option("grad_scale_log2", gs)
        _native.set_option("prune_last_layer", prune)
        lo, do, go = ours_grads(state, mel, utt, precision)
        print("ours P=%d fused=%d gs=2^%d prune=%d loss %.9f  dvec rel err vs ref64 %.2e (ref32: %.2e)" % (
            precision, fused, gs, prune, lo, np.linalg.norm(do - d64) / np.linalg.norm(d64),
            np.linalg.norm(d32 - d64) / np.linalg.norm(d64)))
        compare("  vs ref64", go, g64)




def flips_main():
    """python tools/full_parity_probe.py flips [nspk utt T]: per encoder layer, the ReLU gates that differ from the fp64
    run and the error they alone put on dU (gradient at the FFN pre-activation): sqrt(sum over those elements of
    dU64^2 / sum dU64^2).  Ours runs with the last layer unpruned so that every layer's gates can be read back."""
    nspk, utt, T = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (64, 15, 160)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    GE2E, GE2E_Loss, hp = load_reference()
    state = synth.make_state(61)
    mel = synth.make_mel(601, nspk * utt, T)
    B = nspk * utt
    got = {}
    for dtype in (torch.float64, torch.float32):
        m = GE2E(hp)
        m.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=True)
        m = m.to(dtype).cuda().eval()
        crit = GE2E_Loss().to(dtype).cuda()
        keep = {}
        hs = []
        for l in range(3):
            def hook(mod, inp, out, l=l, keep=keep):
                out.retain_grad()
                keep[l] = out
            hs.append(m.transformer.layers[l].linear1.register_forward_hook(hook))
        loss = crit(m(torch.as_tensor(mel).to(dtype).cuda()), utt)
        loss.backward()
        for h in hs:
            h.remove()
        for l in range(3):
            u, du = keep[l].detach(), keep[l].grad.detach()
            if u.shape[0] == T and u.shape[1] == B:          # the reference runs its encoder sequence-first
                u, du = u.transpose(0, 1), du.transpose(0, 1)
            got[(dtype, l)] = (u.reshape(-1, 1024) > 0, u.double().reshape(-1, 1024).abs().cpu() if dtype == torch.float64 else None,
                               du.double().reshape(-1, 1024).cpu() if dtype == torch.float64 else None)
        del m, loss, keep
    from speaker_embedding_torch_b200 import GE2E as Ours, GE2E_Loss as OursLoss, _native as N
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    N.set_option("prune_last_layer", 0)
    m = Ours(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=True)
    m = m.cuda().eval()
    m._debug_keep_ws = True
    crit = OursLoss().cuda()
    loss = crit(m(torch.as_tensor(mel).cuda()), utt)
    loss.backward()
    torch.cuda.synchronize()
    lay = N.debug_layout(m._cfg, B, T, 1, m._last_meta[3], True)
    for l in (2, 1, 0):
        g64, au64, du64 = got[(torch.float64, l)]
        den = du64.pow(2).sum()
        ours = N.read_split(m._last_ws, lay, "L%d.f" % l, B * T, 1024, 2) > 0
        for tag, gate in (("ref32", got[(torch.float32, l)][0]), ("ours", ours)):
            flip = (gate != g64).cpu()
            print("layer %d %-5s gates that differ %4d of %d (max |u64| there %.2e) -> dU rel err from them alone %.3e" % (
                l, tag, int(flip.sum()), flip.numel(), float(au64[flip].max()) if flip.any() else 0.0,
                float((du64[flip].pow(2).sum() / den).sqrt())), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "flips":
    flips_main()
    sys.exit(0)

if __name__ == "__main__":
    main()
