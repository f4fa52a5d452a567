import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import synth
from speaker_embedding_torch_b200 import GE2E, GE2E_Loss, _native
from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
state = synth.make_state(8)
def model():
    m = GE2E(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()})
    return m.cuda()
worst = 0.0
for B, utt, T in ((1, 1, 1), (1, 1, 5), (2, 1, 64), (1, 1, 192), (2, 2, 191), (300, 2, 17), (2, 1, 193)):
    mel = torch.as_tensor(synth.make_mel(100 + T, B, T)).cuda()
    outs = {}
    for fast in (1, 0):
        _native.set_option("inference_attention_two_ctas", fast)
        _native.set_option("training_attention_two_ctas", fast)
        _native.set_option("fused_layernorm", 2 if fast else 0)
        m = model().eval()
        with torch.no_grad():
            di = m(mel)
        m.train(False)
        d = m(mel)                      # training-precision path with grads
        loss = GE2E_Loss().cuda()(d, utt) if B // utt >= 2 else d.sum()
        loss.backward()
        g = torch.cat([p.grad.flatten() for p in m.parameters()])
        outs[fast] = (di, d.detach(), g)
    di = float((outs[1][0] - outs[0][0]).abs().max()); dt = float((outs[1][1] - outs[0][1]).abs().max())
    gr = float((outs[1][2] - outs[0][2]).double().norm() / outs[0][2].double().norm().clamp_min(1e-30))
    worst = max(worst, dt, gr if gr < 1 else 0)
    print("B %d T %d: infer max diff %.2e  train dvec max diff %.2e  grad rel %.2e" % (B, T, di, dt, gr), flush=True)
for o in ("inference_attention_two_ctas", "training_attention_two_ctas"): _native.set_option(o, 1)
_native.set_option("fused_layernorm", 2)
print("ok")
