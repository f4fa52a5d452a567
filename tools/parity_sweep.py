import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
os.environ.pop("SPK_PROBE_ALL", None)
import full_parity_probe as P
from oracle import synth
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
GE2E, GE2E_Loss, hp = P.load_reference()
rows = []
for ss, ms, T in ((71, 701, 160), (72, 702, 140), (73, 703, 180), (74, 704, 150), (75, 705, 170), (76, 706, 160), (77, 707, 144), (78, 708, 176)):
    state = synth.make_state(ss); mel = synth.make_mel(ms, 960, T)
    l64, d64, g64 = P.reference_grads(GE2E, GE2E_Loss, hp, state, mel, 15, torch.float64)
    l32, d32, g32 = P.reference_grads(GE2E, GE2E_Loss, hp, state, mel, 15, torch.float32)
    lo, do, go = P.ours_grads(state, mel, 15, 2)
    def rel(a, b):
        num = sum(((a[n] - b[n]) ** 2).sum() for n in a); den = sum((b[n] ** 2).sum() for n in a)
        return (num / den) ** 0.5
    cos = (do * d64).sum(1) / (np.linalg.norm(do, axis=1) * np.linalg.norm(d64, axis=1))
    print("seed %d T %d: ours grad_rel %.3e  ref32 grad_rel %.3e  loss_rel %.1e  min_cos %.10f" % (ss, T, rel(go, g64), rel(g32, g64), abs(lo - l64) / l64, cos.min()), flush=True)
