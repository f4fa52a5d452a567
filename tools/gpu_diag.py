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


def enc_bwd_case(Nspk, M, T):
    import numpy as np
    import torch
    from oracle import ge2e_oracle as O, synth
    from speaker_embedding_torch_b200 import GE2E_Loss
    m, state = _model(33)
    m.eval()            # dropout off: parity is defined in eval mode
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
    print("enc bwd N=%d M=%d T=%d: loss %.6f ref %.6f | global grad rel-L2 %.3e | worst %s %.3e | dw %.4e ref %.4e"
          % (Nspk, M, T, loss.item(), loss_ref, rel_all, worst[0], worst[1], crit.weight.grad.item(),
             g_ref["loss.weight"]))
    return abs(loss.item() - loss_ref) <= 1e-3 * abs(loss_ref) and rel_all <= 1e-3


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
    "ge2e_small": lambda: ge2e_case(7, 5, True),
    "ge2e_64x15": lambda: ge2e_case(64, 15, True),
    "ge2e_nonunit": lambda: ge2e_case(33, 3, False),
    "ge2e_256": lambda: ge2e_case(256, 15, True),
    "enc_fwd_small_p2": lambda: enc_fwd_case(6, 24, 1, 2),
    "enc_fwd_small_p1": lambda: enc_fwd_case(6, 24, 1, 1),
    "enc_fwd_160_p1": lambda: enc_fwd_case(8, 160, 1, 1),
    "enc_fwd_177_p2": lambda: enc_fwd_case(4, 177, 1, 2),
    "enc_fwd_multislice": lambda: enc_fwd_case(10, 64, 5, 1),
    "enc_bwd_small": lambda: enc_bwd_case(3, 2, 24),
    "enc_bwd_mid": lambda: enc_bwd_case(4, 3, 160),
}

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
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
        for line in out[-14:]:
            print("   ", line)
        summary.append((name, status))
        sys.stdout.flush()
    print("SUMMARY " + " ".join("%s=%s" % s for s in summary))
