"""Mel front-end: oracle vs the reference's own `mel_spectrogram` (tests/golden/mel.npz, made by
oracle/make_golden_mel.py from /root/reference/meldataset.py:73-96) on CPU; the CUDA kernel vs both on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle
from oracle.make_golden_mel import make_audio

SOUND = dict(n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=None)


def test_oracle_matches_the_reference_front_end(golden_dir):
    g = np.load(os.path.join(golden_dir, "mel.npz"))
    for i in range(int(g["num_cases"])):
        seed, batch, samples = [int(v) for v in g["case%d_meta" % i]]
        want = g["case%d_mel" % i]
        got = mel_oracle.mel_spectrogram_np(make_audio(seed, batch, samples), **SOUND)
        assert got.shape == want.shape == (batch, 80, samples // 256)
        np.testing.assert_allclose(got, want, atol=1e-4, rtol=0)      # fp64 restatement vs the reference's fp32 run


def test_mel_filters_are_slaney_triangles():
    from speaker_embedding_torch_b200.meldataset import mel_filters
    w = mel_filters(22050, 1024, 80, 0.0, None)
    assert w.shape == (80, 513) and w.dtype == np.float32 and (w >= 0).all()
    assert np.array_equal(w, mel_oracle.mel_filterbank(22050, 1024, 80, 0.0, None))
    peaks = w.argmax(axis=1)
    assert (np.diff(peaks) > 0).all()                                  # centre frequencies increase
    freqs = np.linspace(0, 11025, 513)
    area = (w * (freqs[1] - freqs[0])).sum(axis=1)
    np.testing.assert_allclose(area[5:], 1.0, atol=0.08)               # Slaney area normalisation
    for row in w:                                                      # one contiguous band per filter
        nz = np.nonzero(row)[0]
        assert nz.size and nz[-1] - nz[0] + 1 == nz.size


@pytest.mark.gpu
def test_device_front_end_matches_the_reference(golden_dir):
    from speaker_embedding_torch_b200.meldataset import mel_spectrogram
    g = np.load(os.path.join(golden_dir, "mel.npz"))
    for i in range(int(g["num_cases"])):
        seed, batch, samples = [int(v) for v in g["case%d_meta" % i]]
        y = make_audio(seed, batch, samples)
        got = mel_spectrogram(torch.as_tensor(y).cuda(), **SOUND)
        torch.cuda.synchronize()
        assert got.shape == (batch, 80, samples // 256) and got.dtype == torch.float32
        np.testing.assert_allclose(got.cpu().numpy(), g["case%d_mel" % i], atol=2e-4, rtol=0)
        np.testing.assert_allclose(got.cpu().numpy(), mel_oracle.mel_spectrogram_np(y, **SOUND), atol=2e-4, rtol=0)
        half = mel_spectrogram(torch.as_tensor(y).cuda(), out_dtype=torch.float16, **SOUND)
        assert half.dtype == torch.float16
        assert torch.equal(half, got.half())
    # other FFT sizes / a window shorter than n_fft, against the oracle
    y = make_audio(9, 2, 5000)
    for n_fft, hop, win in ((512, 128, 512), (2048, 512, 1200)):
        got = mel_spectrogram(torch.as_tensor(y).cuda(), n_fft, 40, 16000, hop, win, 50.0, 7000.0)
        want = mel_oracle.mel_spectrogram_np(y, n_fft, 40, 16000, hop, win, 50.0, 7000.0)
        np.testing.assert_allclose(got.cpu().numpy(), want, atol=3e-4, rtol=0)
    with pytest.raises(RuntimeError):
        mel_spectrogram(torch.as_tensor(y), **SOUND)                    # host tensor: no CPU path


@pytest.mark.gpu
def test_wav_to_dvector_on_the_device(golden_dir):
    """Inference.py:59-85 + 157-159 without the host: audio -> mel (device) -> 5 x 64 / 32 slices -> d-vector equals
    the encoder fed with the reference's own mel."""
    from oracle import synth
    from speaker_embedding_torch_b200 import GE2E
    from speaker_embedding_torch_b200.Arg_Parser import default_hyper_parameters
    from speaker_embedding_torch_b200.meldataset import mel_spectrogram
    g = np.load(os.path.join(golden_dir, "mel.npz"))
    seed, batch, samples = [int(v) for v in g["case2_meta"]]           # 86 frames
    m = GE2E(default_hyper_parameters())
    m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.make_state(17).items()}, strict=True)
    m = m.cuda().eval()
    mel = mel_spectrogram(torch.as_tensor(make_audio(seed, batch, samples)).cuda(), **SOUND)
    ref = torch.as_tensor(g["case2_mel"]).cuda()
    frame, overlap, samples_per_utt = 24, 12, 5                         # required length 72 <= 86
    need = samples_per_utt * (frame - overlap) + overlap
    with torch.no_grad():
        a = m.embed_windows(mel[:, :, :need].contiguous(), frame, overlap)
        b = m.embed_windows(ref[:, :, :need].contiguous(), frame, overlap)
    cos = torch.nn.functional.cosine_similarity(a, b, dim=1)
    assert a.shape == (batch, 256) and cos.min().item() >= 0.9999
