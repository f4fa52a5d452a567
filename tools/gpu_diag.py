"""Bring-up diagnostics on a B200: each case runs in its own process (a device trap poisons the context).

    python tools/gpu_diag.py            # run every case, print a summary table
    python tools/gpu_diag.py CASE       # run one case in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gemm_case(planes, m, n, k, a_mn, b_mn, ksplit=1, block_n=0, relu_bias=False):
    import torch
    from speaker_embedding_torch_b200 import _native as N
    torch.manual_seed(m * 7 + n * 3 + k)
    dev = "cuda"
    A = torch.randn(m, k, device=dev)
    B = torch.randn(n, k, device=dev)
    a_store = A.t().contiguous() if a_mn else A
    b_store = B.t().contiguous() if b_mn else B
    a_s = N.split_pack(a_store, planes)
    b_s = N.split_pack(b_store, planes)
    a_eff = N.split_unpack(a_s)
    b_eff = N.split_unpack(b_s)
    a_eff = a_eff.t() if a_mn else a_eff
    b_eff = b_eff.t() if b_mn else b_eff
    ref = a_eff.double() @ b_eff.double().t()
    bias = torch.randn(n, device=dev) if relu_bias else None
    if relu_bias:
        ref = torch.relu(ref + bias.double())
    if ksplit > 1:
        out = torch.zeros(m, n, device=dev)
        N.gemm(a_s, b_s, planes, m, n, k, a_mn, b_mn, atomic_out=out, ksplit=ksplit, block_n=block_n)
    else:
        out = N.gemm(a_s, b_s, planes, m, n, k, a_mn, b_mn, bias=bias, relu=relu_bias, out_f32=True, block_n=block_n)
    torch.cuda.synchronize()
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print("gemm planes=%d m=%d n=%d k=%d a_mn=%d b_mn=%d ks=%d bn=%d: max_abs_err=%.3e (ref max %.3e)"
          % (planes, m, n, k, a_mn, b_mn, ksplit, block_n, err, scale))
    if err > 1e-2 * scale:
        print("  out[0,:8] =", out[0, :8].tolist())
        print("  ref[0,:8] =", ref[0, :8].tolist())
        print("  out[1,:4] =", out[1, :4].tolist(), " ref[1,:4] =", ref[1, :4].tolist())
        bad = ((out.double() - ref).abs() > 1e-2 * scale)
        print("  bad fraction %.4f; bad rows %s ; bad cols %s" % (
            bad.float().mean().item(), bad.any(1).nonzero().flatten()[:16].tolist(),
            bad.any(0).nonzero().flatten()[:16].tolist()))
    # two-plane products must be ~fp32-accurate wrt the *unsplit* operands too
    tol = 2e-2 if planes == 1 else 2e-4
    if planes == 3:
        ref32 = A.double() @ B.double().t()
        if relu_bias:
            ref32 = torch.relu(ref32 + bias.double())
        print('  vs unsplit fp32 operands: max_abs_err=%.3e' % (out.double() - ref32).abs().max().item())
    return err <= tol * max(scale, 1.0) * 0.05 + 1e-3 * (planes == 1) * scale


def ge2e_case(n, m, unit=True):
    import numpy as np
    import torch
    from oracle import ge2e_oracle as O, synth
    from speaker_embedding_torch_b200 import GE2E_Loss
    E = synth.make_embeddings(5 + n, n, m, unit_norm=unit)
    loss_ref, dE_ref, dw_ref, db_ref = O.ge2e_loss_and_grads_closed_form(E, m, 10.0, -5.0)
    crit = GE2E_Loss().cuda()
    e = torch.as_tensor(E).cuda().requires_grad_(True)
    loss = crit(e, m)
    loss.backward()
    torch.cuda.synchronize()
    dE = e.grad.cpu().numpy().astype(np.float64)
    rel = np.linalg.norm(dE - dE_ref) / max(np.linalg.norm(dE_ref), 1e-30)
    print("ge2e N=%d M=%d: loss %.7f ref %.7f | dE rel %.3e | dw %.6e ref %.6e | db %.3e"
          % (n, m, loss.item(), loss_ref, rel, crit.weight.grad.item(), dw_ref, crit.bias.grad.item()))
    return abs(loss.item() - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref)) and rel < 1e-4


def _model(seed):
    import torch
    from oracle import synth
    from speaker_embedding_torch_b200 import GE2E
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    state = synth.make_state(seed)
    m = GE2E(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=True)
    return m.cuda(), state


def enc_fwd_case(B, T, S, precision):
    import numpy as np
    import torch
    from oracle import ge2e_oracle as O, synth
    m, state = _model(31)
    m.eval()
    m.eval_precision = precision
    mel = synth.make_mel(400 + T, B, T)
    with torch.no_grad():
        d = m(torch.as_tensor(mel).cuda(), S)
    torch.cuda.synchronize()
    ref = O.encoder_forward(O.to_torch_state(state, torch.float64), torch.as_tensor(mel).double(), S).numpy()
    d = d.cpu().numpy().astype(np.float64)
    cos = (d * ref).sum(1) / (np.linalg.norm(d, axis=1) * np.linalg.norm(ref, axis=1))
    print("enc fwd B=%d T=%d S=%d P=%d: min cos %.7f, max abs err %.3e, |d| in [%.5f, %.5f]"
          % (B, T, S, precision, cos.min(), np.abs(d - ref).max(), np.linalg.norm(d, axis=1).min(),
             np.linalg.norm(d, axis=1).max()))
    return cos.min() >= 0.9999


def enc_bwd_case(Nspk, M, T, precision=3):
    import numpy as np
    import torch
    from oracle import ge2e_oracle as O, synth
    from speaker_embedding_torch_b200 import GE2E_Loss
    m, state = _model(33)
    m.eval()            # dropout off: parity is defined in eval mode
    m.train_precision = precision
    crit = GE2E_Loss().cuda()
    mel = synth.make_mel(500 + T, Nspk * M, T)
    d = m(torch.as_tensor(mel).cuda())
    loss = crit(d, M)
    loss.backward()
    torch.cuda.synchronize()
    loss_ref, d_ref, g_ref = O.train_step_grads(state, mel, M)
    num = den = 0.0
    worst = ("", 0.0)
    for name, p in m.named_parameters():
        g = p.grad.detach().cpu().numpy().astype(np.float64)
        r = g_ref[name]
        num += ((g - r) ** 2).sum()
        den += (r ** 2).sum()
        rel = np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-30)
        if rel > worst[1]:
            worst = (name, rel)
    rel_all = (num / den) ** 0.5
    if rel_all > 1e-3 or os.environ.get("DIAG_VERBOSE"):
        for name, p in m.named_parameters():
            g = p.grad.detach().cpu().numpy().astype(np.float64)
            r = g_ref[name]
            print("   %-50s rel %.3e  |ref| %.3e" % (name, np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-30),
                                                  np.linalg.norm(r)))
    print("enc bwd P=%d N=%d M=%d T=%d: loss %.6f ref %.6f | global grad rel-L2 %.3e | worst %s %.3e | dw %.4e ref %.4e"
          % (precision, Nspk, M, T, loss.item(), loss_ref, rel_all, worst[0], worst[1], crit.weight.grad.item(),
             g_ref["loss.weight"]))
    return abs(loss.item() - loss_ref) <= 1e-3 * abs(loss_ref) and rel_all <= 1e-3


def enc_stage_case(Nspk, M, T, layers=1, precision=3):
    """Single-layer model: compare every forward / backward stage buffer with an fp64 torch restatement."""
    import math
    import numpy as np
    import torch
    from oracle import synth
    from speaker_embedding_torch_b200 import GE2E, GE2E_Loss, _native as N
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    hp = default_hyper_parameters()
    hp.GE2E.Transformer.Num_Layers = layers
    N.set_option("prune_last_layer", 0)           # the stage buffers of the last layer are full-size only then
    N.set_option("fused_training_attention", int(os.environ.get("SPK_DIAG_FUSED", "0")))
    full = synth.make_state(33)
    m = GE2E(hp)
    sd = {k: torch.as_tensor(v) for k, v in full.items() if k in m.state_dict()}
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m._debug_keep_ws = True
    m.train_precision = precision
    crit = GE2E_Loss().cuda()
    B = Nspk * M
    mel = synth.make_mel(500 + T, B, T)
    d = m(torch.as_tensor(mel).cuda())
    loss = crit(d, M)
    loss.backward()
    torch.cuda.synchronize()
    ws = m._last_ws
    lay = N.debug_layout(m._cfg, B, T, 1, precision, True)
    D, H, Mt = 256, 4, B * T
    Tp = (T + 7) // 8 * 8

    # fp64 restatement with retained grads
    st = {k: v.double().clone().requires_grad_(not k.endswith(".pe")) for k, v in sd.items()}
    x = torch.as_tensor(mel).double().transpose(1, 2)
    inter = {}

    def keep(name, t):
        t.retain_grad()
        inter[name] = t
        return t
    u0 = keep("upre", x @ st["prenet.weight"][:, :, 0].t() + st["prenet.bias"])
    h = keep("h0", torch.relu(u0) + st["positional_encoding.alpha"] * st["positional_encoding.pe"][0, :, :T].t())
    l = layers - 1
    for li in range(layers):
        p = "transformer.layers.%d." % li
        qkv = keep("qkv%d" % li, h @ st[p + "self_attn.in_proj_weight"].t() + st[p + "self_attn.in_proj_bias"])
        q, k, v = [t.reshape(B, T, H, 64).transpose(1, 2) for t in (qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:])]
        sraw = keep("sraw%d" % li, q @ k.transpose(-1, -2))
        pr = keep("p%d" % li, torch.softmax(sraw / 8.0, dim=-1))
        att = keep("att%d" % li, (pr @ v).transpose(1, 2).reshape(B, T, D))
        z1 = keep("z1%d" % li, h + att @ st[p + "self_attn.out_proj.weight"].t() + st[p + "self_attn.out_proj.bias"])
        h1 = keep("h1%d" % li, torch.nn.functional.layer_norm(z1, (D,), st[p + "norm1.weight"], st[p + "norm1.bias"]))
        uu = keep("u%d" % li, h1 @ st[p + "linear1.weight"].t() + st[p + "linear1.bias"])
        f = keep("f%d" % li, torch.relu(uu))
        z2 = keep("z2%d" % li, h1 + f @ st[p + "linear2.weight"].t() + st[p + "linear2.bias"])
        h = keep("hout%d" % li, torch.nn.functional.layer_norm(z2, (D,), st[p + "norm2.weight"], st[p + "norm2.bias"]))
    h0n = torch.nn.functional.layer_norm(h[:, 0, :], (D,), st["transformer.norm.weight"], st["transformer.norm.bias"])
    e = h0n @ st["projection.weight"][:, :, 0].t() + st["projection.bias"]
    dv = e / e.norm(dim=1, keepdim=True)
    from oracle import ge2e_oracle as O
    lref = O.ge2e_loss(dv, M, 10.0, -5.0)
    lref.backward()

    def rel(a, b, fit=False):
        a = a.double().cpu().reshape(-1)
        b = b.detach().double().reshape(-1)
        if fit:     # backward stage buffers are carried at the pass's power-of-two gradient scale: fit it out
            a = a * (2.0 ** torch.round(torch.log2((a @ b) / (a @ a))))
        return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()

    def rd(name, rows, cols):
        bwd = name in ("dh_a", "dh_b", "dz", "dzd", "df", "datt", "dqkv", "scr", "ds")
        return N.read_split(ws, lay, name, rows, cols, min(precision, 2) if bwd else precision)
    print("loss %.7f ref %.7f" % (loss.item(), lref.item()))
    pre = "L%d." % l
    print("fwd  h0 %.2e qkv %.2e att %.2e z1 %.2e h1 %.2e f %.2e z2 %.2e hout %.2e" % (
        rel(rd("h0", Mt, D), inter["h0"]), rel(rd(pre + "qkv", Mt, 3 * D), inter["qkv%d" % l]),
        rel(rd(pre + "att", Mt, D), inter["att%d" % l]), rel(rd(pre + "z1", Mt, D), inter["z1%d" % l]),
        rel(rd(pre + "h1", Mt, D), inter["h1%d" % l]), rel(rd(pre + "f", Mt, 4 * D), inter["f%d" % l]),
        rel(rd(pre + "z2", Mt, D), inter["z2%d" % l]), rel(rd(pre + "hout", Mt, D), inter["hout%d" % l])))
    pmine = rd(pre + "p", B * H * T, Tp).view(B, H, T, Tp)[..., :T]
    print("fwd  p %.2e" % rel(pmine, inter["p%d" % l]))
    # backward buffers hold the values of the LAST processed layer (layer 0)
    ds_mine = rd("ds", B * H * T, Tp).view(B, H, T, Tp)[..., :T]
    print("bwd(layer0) dS %.2e dqkv %.2e datt %.2e dZ1 %.2e dU %.2e dH0 %.2e du0 %.2e" % (
        rel(ds_mine, inter["sraw0"].grad, True), rel(rd("dqkv", Mt, 3 * D), inter["qkv0"].grad, True),
        rel(rd("datt", Mt, D), inter["att0"].grad, True), rel(rd("dz", Mt, D), inter["z10"].grad, True),
        rel(rd("df", Mt, 4 * D), inter["u0"].grad, True),
        rel(rd("dh_a", Mt, D), inter["h0"].grad, True), rel(rd("dh_b", Mt, D), inter["upre"].grad, True)))
    # local check of one backward GEMM on this run's own operands: dATT = dZ1 Wo (K = 256, dense rows)
    dz_mine = rd("dz", Mt, D).double().cpu()
    datt_mine = rd("datt", Mt, D).double().cpu()
    wo = sd["transformer.layers.0.self_attn.out_proj.weight"].double()
    chk = dz_mine @ wo
    print("   local: dATT vs fp64(dZ1_mine @ Wo) %.2e   (both at the pass's scale)" % (
        ((datt_mine - chk).norm() / chk.norm()).item()))
    df_mine = rd("df", Mt, 4 * D).double().cpu()
    print("   local magnitudes: |dZ1| max %.3e rms %.3e   |dU| max %.3e rms %.3e  frac(|dU|<0.125, nonzero) %.3f" % (
        dz_mine.abs().max(), dz_mine.pow(2).mean().sqrt(), df_mine.abs().max(), df_mine.pow(2).mean().sqrt(),
        float(((df_mine.abs() < 0.125) & (df_mine != 0)).double().sum() / max(1.0, float((df_mine != 0).sum())))))

    def dist(name, mine, ref):
        mine = mine.double().cpu().reshape(ref.shape)
        ref = ref.detach().double()
        mine = mine * (2.0 ** torch.round(torch.log2((mine * ref).sum() / (mine * mine).sum())))
        err = (mine - ref).abs()
        mx = ref.abs().max().item()
        big = err > 1e-3 * mx
        mism = (mine != 0) != (ref != 0)
        idx = err.reshape(-1).topk(5).indices
        print("   [%s] max|ref| %.3e  n(err>1e-3 max) %d of %d  zero-pattern mismatches %d  err-energy in top5 %.3f"
              % (name, mx, int(big.sum()), err.numel(), int(mism.sum()),
                 float((err.reshape(-1)[idx] ** 2).sum() / (err ** 2).sum())))
        for i in idx.tolist():
            print("       flat %d mine %.6e ref %.6e" % (i, mine.reshape(-1)[i].item(), ref.reshape(-1)[i].item()))
    dist("dU", rd("df", Mt, 4 * D), inter["u0"].grad.reshape(Mt, 4 * D))
    dist("dZ1", rd("dz", Mt, D), inter["z10"].grad.reshape(Mt, D))
    fm = rd(pre + "f", Mt, 4 * D).double().cpu()
    ur = inter["u%d" % l].detach().reshape(Mt, 4 * D)
    flips = ((fm > 0) != (ur > 0))
    print("   forward ReLU gate flips (last layer): %d of %d ; min |u_ref| at flips %s" % (
        int(flips.sum()), flips.numel(), ur[flips].abs().max().item() if flips.any() else None))
    for sub in ("q", "k", "v"):
        i = "qkv".index(sub)
        a = rd("dqkv", Mt, 3 * D)[:, i * D:(i + 1) * D]
        b = inter["qkv0"].grad.reshape(Mt, 3 * D)[:, i * D:(i + 1) * D]
        print("   d%s %.2e" % (sub, rel(a, b, True)))
    worst = 0.0
    for name, p_ in m.named_parameters():
        r = rel(p_.grad, st[name].grad)
        worst = max(worst, r)
        print("   %-46s rel %.2e" % (name, r))
    return worst < 1e-3


CASES = {
    "gemm_kk_p1_tile": lambda: gemm_case(1, 128, 64, 64, 0, 0),
    "gemm_kk_p1_k256": lambda: gemm_case(1, 256, 256, 256, 0, 0),
    "gemm_kk_p1_bn128": lambda: gemm_case(1, 384, 128, 128, 0, 0),
    "gemm_kk_p1_bn192": lambda: gemm_case(1, 160, 192, 64, 0, 0),
    "gemm_kk_p2": lambda: gemm_case(2, 256, 256, 256, 0, 0),
    "gemm_kk_p2_big": lambda: gemm_case(2, 4000, 768, 256, 0, 0, relu_bias=True),
    "gemm_kk_p1_ragged": lambda: gemm_case(1, 300, 184, 80, 0, 0),
    "gemm_kmn_p1": lambda: gemm_case(1, 256, 256, 256, 0, 1),
    "gemm_kmn_p2": lambda: gemm_case(2, 300, 64, 160, 0, 1),
    "gemm_mnmn_p1": lambda: gemm_case(1, 256, 256, 256, 1, 1),
    "gemm_mnmn_p2_splitk": lambda: gemm_case(2, 256, 1024, 4096, 1, 1, ksplit=8),
    "gemm_mnmn_p1_n80": lambda: gemm_case(1, 256, 80, 1000, 1, 1, ksplit=3),
    "gemm_kk_p3": lambda: gemm_case(3, 384, 256, 256, 0, 0),
    "gemm_kmn_p3": lambda: gemm_case(3, 300, 320, 160, 0, 1, relu_bias=True),
    "gemm_mnmn_p3": lambda: gemm_case(3, 256, 128, 512, 1, 1),
    "gemm_mnmn_p2_att": lambda: gemm_case(2, 160, 64, 160, 1, 1),
    "gemm_kmn_p2_att": lambda: gemm_case(2, 160, 64, 160, 0, 1),
    "gemm_kk_p2_att": lambda: gemm_case(2, 160, 160, 64, 0, 0),
    "gemm_kk_p1_multitile": lambda: gemm_case(1, 128 * 400, 256, 256, 0, 0, relu_bias=True),
    "gemm_kmn_p2_multitile": lambda: gemm_case(2, 128 * 300, 1024, 256, 0, 1),
    "ge2e_small": lambda: ge2e_case(7, 5, True),
    "ge2e_64x15": lambda: ge2e_case(64, 15, True),
    "ge2e_nonunit": lambda: ge2e_case(33, 3, False),
    "ge2e_256": lambda: ge2e_case(256, 15, True),
    "enc_fwd_small_p2": lambda: enc_fwd_case(6, 24, 1, 2),
    "enc_fwd_small_p1": lambda: enc_fwd_case(6, 24, 1, 1),
    "enc_fwd_160_p1": lambda: enc_fwd_case(8, 160, 1, 1),
    "enc_fwd_177_p2": lambda: enc_fwd_case(4, 177, 1, 2),
    "enc_fwd_multislice": lambda: enc_fwd_case(10, 64, 5, 1),
    "stage_t160": lambda: enc_stage_case(4, 3, 160),
    "stage_t160_p2": lambda: enc_stage_case(4, 3, 160, precision=2),
    "enc_bwd_mid_p2": lambda: enc_bwd_case(4, 3, 160, 2),
    "enc_bwd_big_p3": lambda: enc_bwd_case(16, 5, 160, 3),
    "enc_bwd_big_p2": lambda: enc_bwd_case(16, 5, 160, 2),
    "stage_t24": lambda: enc_stage_case(4, 3, 24),
    "stage_t160_l3": lambda: enc_stage_case(4, 3, 160, layers=3),
    "stage_t160_l2": lambda: enc_stage_case(4, 3, 160, layers=2, precision=2),
    "enc_bwd_small": lambda: enc_bwd_case(3, 2, 24),
    "enc_bwd_mid": lambda: enc_bwd_case(4, 3, 160),
    "enc_bwd_t64": lambda: enc_bwd_case(2, 2, 64),
    "enc_bwd_t72": lambda: enc_bwd_case(2, 2, 72),
    "enc_bwd_t128": lambda: enc_bwd_case(2, 2, 128),
    "enc_bwd_t136": lambda: enc_bwd_case(2, 2, 136),
    "enc_bwd_t24_b40": lambda: enc_bwd_case(10, 4, 24),
}

if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] in CASES:
        ok = CASES[sys.argv[1]]()
        print("RESULT", sys.argv[1], "PASS" if ok else "FAIL")
        sys.exit(0 if ok else 1)
    names = sys.argv[1:] or list(CASES)
    summary = []
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=180)
            out = (p.stdout + p.stderr).strip().splitlines()
            status = "PASS" if p.returncode == 0 else "FAIL(rc=%d)" % p.returncode
        except subprocess.TimeoutExpired:
            out, status = ["timeout"], "TIMEOUT"
        print("=== %s: %s (%.1fs)" % (name, status, time.time() - t0))
        for line in out[-70:]:
            print("   ", line)
        summary.append((name, status))
        sys.stdout.flush()
    print("SUMMARY " + " ".join("%s=%s" % s for s in summary))
