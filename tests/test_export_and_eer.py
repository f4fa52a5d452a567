"""Versioned weight-arena export and the EER harness -- host logic, no GPU needed."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import synth


def _model(seed):
    from speaker_embedding_torch_b200 import GE2E
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    m = GE2E(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.make_state(seed).items()}, strict=True)
    return m


def test_arena_round_trip_and_rejections(tmp_path):
    from speaker_embedding_torch_b200 import Export
    m = _model(3)
    path = str(tmp_path / "encoder.spkw")
    header = Export.export_arena(m, path, steps=1234)
    assert header["floats"] >= 2456321 and len(header["tensors"]) == 43 and header["steps"] == 1234
    assert all(t["offset"] % 4 == 0 for t in header["tensors"])                 # 16-byte aligned tensors
    h2, state = Export.load_arena(path)
    assert h2 == header and len(state) == 44
    for k, v in m.state_dict().items():
        assert torch.equal(state[k], v.cpu()), k
    m2 = Export.model_from_arena(path)
    assert not m2.training and all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    raw = bytearray(open(path, "rb").read())
    bad = str(tmp_path / "bad.spkw")
    flipped = bytearray(raw); flipped[-5] ^= 0x40
    open(bad, "wb").write(flipped)
    with pytest.raises(RuntimeError, match="checksum"):
        Export.load_arena(bad)
    open(bad, "wb").write(raw[:-8])
    with pytest.raises(RuntimeError, match="payload"):
        Export.load_arena(bad)
    newer = bytearray(raw); newer[8:12] = struct.pack("<I", Export.FORMAT_VERSION + 1)
    open(bad, "wb").write(newer)
    with pytest.raises(RuntimeError, match="version"):
        Export.load_arena(bad)
    open(bad, "wb").write(b"not an arena at all")
    with pytest.raises(RuntimeError, match="not a speaker-embedding"):
        Export.load_arena(bad)


def test_equal_error_rate_known_answers():
    from speaker_embedding_torch_b200.Verification import equal_error_rate
    # perfectly separated
    eer, thr = equal_error_rate([0.9, 0.8, 0.7, 0.2, 0.1, 0.0], [1, 1, 1, 0, 0, 0])
    assert eer == 0.0 and 0.2 < thr <= 0.7
    # fully inverted
    assert equal_error_rate([0.1, 0.2, 0.8, 0.9], [1, 1, 0, 0])[0] == 1.0
    # one swap in 4 + 4: FRR = FAR = 0.25 at the threshold between them
    eer, _ = equal_error_rate([0.9, 0.8, 0.7, 0.4, 0.6, 0.3, 0.2, 0.1], [1, 1, 1, 1, 0, 0, 0, 0])
    assert abs(eer - 0.25) < 1e-12
    # two Gaussians one sigma' apart: EER = Phi(-d/2)
    rng = np.random.default_rng(0)
    n, d = 200000, 2.0
    s = np.r_[rng.standard_normal(n) + d, rng.standard_normal(n)]
    t = np.r_[np.ones(n), np.zeros(n)]
    eer, thr = equal_error_rate(s, t)
    from math import erf, sqrt
    want = 0.5 * (1 + erf(-d / 2 / sqrt(2)))
    assert abs(eer - want) < 3e-3 and abs(thr - d / 2) < 0.05
    # all scores tied: no threshold separates anything
    assert abs(equal_error_rate([0.5] * 6, [1, 1, 1, 0, 0, 0])[0] - 0.5) < 1e-12


def test_trials_and_scoring():
    from speaker_embedding_torch_b200.Verification import cosine_scores, evaluate_eer, make_trials
    labels = ["a"] * 5 + ["b"] * 4 + ["c"] * 3 + ["d"]
    a, b, t = make_trials(labels, 400, seed=1)
    lab = np.asarray(labels)
    assert len(a) == 400 and t.sum() == 200
    assert (lab[a[t == 1]] == lab[b[t == 1]]).all() and (a[t == 1] != b[t == 1]).all()
    assert (lab[a[t == 0]] != lab[b[t == 0]]).all()
    rng = np.random.default_rng(2)
    centres = rng.standard_normal((4, 256))
    idx = np.asarray([0] * 5 + [1] * 4 + [2] * 3 + [3])
    tight = torch.as_tensor(centres[idx] + 0.05 * rng.standard_normal((13, 256)), dtype=torch.float32)
    s = cosine_scores(tight, a, b)
    assert s.shape == (400,) and float(s[torch.as_tensor(t == 1)].min()) > float(s[torch.as_tensor(t == 0)].max())
    assert evaluate_eer(tight, labels, 2000) == 0.0
    noise = torch.as_tensor(rng.standard_normal((13, 256)), dtype=torch.float32)
    assert 0.3 < evaluate_eer(noise, labels, 4000) < 0.7
    with pytest.raises(RuntimeError):
        make_trials(["x", "y"], 10)
