"""Fused optimiser parity, checkpoint / trace drop-in contracts, and a Train.py-shaped training loop."""
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import synth

pytestmark = pytest.mark.gpu


def test_fused_radam_matches_reference_trajectory(golden_dir):
    """tests/golden/radam.npz was produced by the reference's Radam.py + Modified_Noam_Scheduler."""
    from speaker_embedding_torch_b200.Noam_Scheduler import Modified_Noam_Scheduler
    from speaker_embedding_torch_b200.Radam import RAdam
    g = np.load(os.path.join(golden_dir, "radam.npz"))
    p = torch.nn.Parameter(torch.as_tensor(g["p0"], dtype=torch.float32).cuda())
    opt = RAdam([p], lr=2e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0)
    sch = Modified_Noam_Scheduler(opt, base=4000)
    for t, grad in enumerate(g["grads"]):
        p.grad = torch.as_tensor(grad, dtype=torch.float32).cuda()
        np.testing.assert_allclose(opt.param_groups[0]["lr"], g["lrs"][t], rtol=1e-12)
        opt.step()
        sch.step()
        np.testing.assert_allclose(p.detach().cpu().numpy(), g["traj"][t], rtol=0, atol=3e-6)
    st = opt.state[p]
    assert st["step"] == len(g["grads"]) and set(st) == {"step", "exp_avg", "exp_avg_sq"}


def test_fused_adamw_and_clip_match_torch():
    from speaker_embedding_torch_b200.Radam import FusedAdamW
    torch.manual_seed(0)
    shapes = [(256, 80, 1), (256,), (768, 256), (1,), (1024, 256), ()]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    o1 = FusedAdamW(ours, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, max_grad_norm=1.0)
    o2 = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01)
    for step in range(6):
        grads = [torch.randn(s, device="cuda") * (3.0 if step % 2 else 0.01) for s in shapes]
        for p, q, gr in zip(ours, ref, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(ref, max_norm=1.0)       # Train.py:154-159
        o2.step()
        o1.step()
        torch.cuda.synchronize()
        torch.testing.assert_close(o1.grad_norm(), total, rtol=1e-5, atol=1e-7)
        for p, q in zip(ours, ref):
            torch.testing.assert_close(p, q, rtol=2e-5, atol=2e-7)


def test_fused_step_many_tensors_groups_and_uneven_step_counts():
    """More than 64 tensors, two parameter groups with different learning rates, one tensor whose gradient is None
    on the first step (its step count lags): the clip norm is GLOBAL over everything that has a gradient, like
    clip_grad_norm_(all parameters), and every tensor gets the bias correction of its own step count; the norm is
    bit-reproducible (deterministic reduction)."""
    from speaker_embedding_torch_b200.Radam import FusedAdamW
    torch.manual_seed(1)
    shapes = [(37, 5)] * 40 + [(129,)] * 30 + [(256, 64)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    split = 35
    o1 = FusedAdamW([{"params": ours[:split], "lr": 1e-3}, {"params": ours[split:], "lr": 3e-3}], betas=(0.9, 0.999),
                    eps=1e-6, weight_decay=0.01, max_grad_norm=0.5)
    o2 = torch.optim.AdamW([{"params": ref[:split], "lr": 1e-3}, {"params": ref[split:], "lr": 3e-3}],
                           betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01)
    norms = []
    for step in range(4):
        for i, (p, q) in enumerate(zip(ours, ref)):
            if step == 0 and i == 3:
                p.grad = q.grad = None
                continue
            gr = torch.randn(p.shape, device="cuda") * 0.3
            p.grad, q.grad = gr.clone(), gr.clone()
        total = torch.nn.utils.clip_grad_norm_([q for q in ref if q.grad is not None], max_norm=0.5)
        o2.step()
        o1.step()
        torch.testing.assert_close(o1.grad_norm(), total, rtol=1e-5, atol=1e-7)
        norms.append(o1.grad_norm().clone())
        for p, q in zip(ours, ref):
            torch.testing.assert_close(p, q, rtol=3e-5, atol=3e-7)
    assert o1.state[ours[3]]["step"] == 3 and o1.state[ours[0]]["step"] == 4
    # same gradients again -> bit-identical norm (no atomics in the reduction)
    o1.step()
    n1 = o1.grad_norm().clone()
    o1.step()
    assert torch.equal(n1, o1.grad_norm())


def _hp_file(tmp_path):
    hp = {"Sound": {"Mel_Dim": 80},
          "GE2E": {"Embedding_Size": 256, "Positional_Encoding": {"Max_Position": 1024, "Dropout_Rate": 0.1},
                   "Transformer": {"Num_Layers": 3, "Head": 4, "Dropout_Rate": 0.1}}}
    path = os.path.join(tmp_path, "Hyper_Parameters.yaml")
    yaml.safe_dump(hp, open(path, "w"))
    return path


def test_checkpoint_roundtrip_and_traced_export(tmp_path):
    """Checkpoint dict of Train.py:298-308 -> Tracer (Trace.py:7-44) -> torch.jit.trace -> save/load."""
    from speaker_embedding_torch_b200 import GE2E
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    from speaker_embedding_torch_b200.Trace import Tracer
    tmp_path = str(tmp_path)
    m = GE2E(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.make_state(12).items()})
    ckpt = os.path.join(tmp_path, "S_100.pt")
    torch.save({"Model": m.state_dict(), "Optimizer": {}, "Scheduler": {}, "Steps": 100}, ckpt)
    tracer = Tracer(_hp_file(tmp_path), ckpt)
    assert tracer.steps == 100 and not any(p.requires_grad for p in tracer.parameters())
    x = torch.rand(1, 80, 400, device="cuda")
    lengths = torch.LongTensor([400]).cuda()
    traced = torch.jit.trace(tracer, (x, lengths), check_trace=False)
    out_path = os.path.join(tmp_path, "ge2e.pts")
    traced.save(out_path)
    loaded = torch.jit.load(out_path)
    eager = tracer(x, lengths)
    torch.testing.assert_close(loaded(x, lengths), eager, atol=0, rtol=0)
    xb = torch.as_tensor(synth.make_mel(3, 4, 200)).cuda()               # generalises over batch and T
    torch.testing.assert_close(loaded(xb, lengths), tracer(xb, lengths), atol=0, rtol=0)
    assert "spkemb::encoder_infer" in str(traced.inlined_graph)
    m2 = m.cuda().eval()
    with torch.no_grad():
        torch.testing.assert_close(m2(x), eager, atol=0, rtol=0)


def test_train_py_shaped_loop_reduces_loss():
    """The statements of Trainer.Train_Step (Train.py:140-168) with the stock torch optimiser."""
    from speaker_embedding_torch_b200 import GE2E, GE2E_Loss
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    torch.manual_seed(0)
    model = GE2E(default_hyper_parameters()).cuda()
    criterion = GE2E_Loss().cuda()
    optimizer = torch.optim.AdamW(params=model.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-6)
    scaler = torch.amp.GradScaler("cuda", enabled=False)
    spk, utt = 8, 4
    rng = np.random.default_rng(0)
    centers = rng.standard_normal((spk, 1, 80, 1)) * 1.5
    feats = torch.as_tensor((centers + rng.standard_normal((spk, utt, 80, 60))).reshape(spk * utt, 80, 60),
                            dtype=torch.float32)
    losses = []
    model.train()
    for _ in range(12):
        features = feats.to("cuda", non_blocking=True)
        with torch.autocast("cuda", enabled=False):
            embeddings = model(features)
            loss = criterion(embeddings, utt)
        optimizer.zero_grad()
        scaler.scale(loss).backward()
        scaler.unscale_(optimizer)
        torch.nn.utils.clip_grad_norm_(parameters=model.parameters(), max_norm=1.0)
        scaler.step(optimizer)
        scaler.update()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and np.mean(losses[-3:]) < np.mean(losses[:3])
    # named_parameters (TensorBoard histogram tags, Train.py:226-232) and hooks keep working
    assert len([n for n, _ in model.named_parameters()]) == 43
    seen = []
    h = model.register_forward_hook(lambda mod, i, o: seen.append(o.shape))
    model.eval()
    with torch.no_grad():
        model(feats.cuda())
    h.remove()
    assert seen == [torch.Size([spk * utt, 256])]


def test_arena_export_serves_the_same_dvectors_and_eer_harness_runs(tmp_path):
    """N4: export -> model_from_arena on the device gives bit-identical d-vectors; the EER harness scores them."""
    from speaker_embedding_torch_b200 import Export, GE2E
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    from speaker_embedding_torch_b200.Verification import evaluate_eer
    m = GE2E(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.make_state(23).items()}, strict=True)
    m = m.cuda().eval()
    path = str(tmp_path / "encoder.spkw")
    Export.export_arena(m, path, steps=7)
    served = Export.model_from_arena(path, device="cuda")
    rng = np.random.default_rng(5)
    spk, utt = 12, 6
    centres = rng.standard_normal((spk, 1, 80, 1)) * 1.5
    feats = torch.as_tensor((-5.0 + centres + 0.7 * rng.standard_normal((spk, utt, 80, 64))).reshape(spk * utt, 80, 64),
                            dtype=torch.float32).cuda()
    with torch.no_grad():
        a, b = m(feats), served(feats)
    assert torch.equal(a, b)
    labels = np.repeat(np.arange(spk), utt)
    eer = evaluate_eer(a, labels, num_trials=4000, seed=0)
    assert 0.0 <= eer <= 0.5          # an untrained encoder still separates these synthetic speakers above chance
