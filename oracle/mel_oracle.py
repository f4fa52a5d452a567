"""CPU restatement of the reference's mel front-end.

TEST INFRASTRUCTURE -- never imported by the product.

  * ``mel_spectrogram_np`` follows /root/reference/meldataset.py:73-96 (HiFi-GAN front-end): reflect-pad by
    (n_fft - hop) / 2, STFT with a periodic Hann window (``torch.hann_window`` default), ``center=False``, magnitude
    ``sqrt(re^2 + im^2 + 1e-9)``, mel-basis matmul, ``log(clamp(., 1e-5))``.  Pinned against the reference's own
    function by oracle/make_golden_mel.py -> tests/golden/mel.npz.
  * ``mel_filterbank`` restates ``librosa.filters.mel`` (Slaney-style mel scale, Slaney area normalisation, the
    defaults ``htk=False, norm='slaney'`` the reference calls it with, meldataset.py:81).  librosa is a third-party
    dependency that is NOT present in this image (the reference pins no version; the algorithm below is the one
    published in librosa >= 0.8): this one function is restated from the published algorithm and is UNPINNED; the
    golden fixture feeds the SAME basis to the reference's ``mel_spectrogram``, so everything else is pinned.
"""
import numpy as np


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """[n_mels, 1 + n_fft // 2] float32 triangular filters (librosa.filters.mel defaults)."""
    if fmax is None:
        fmax = sr / 2.0
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def mel_spectrogram_np(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, basis=None):
    """y [B, samples] -> log-mel [B, num_mels, frames] in fp64 (meldataset.py:73-96, center=False)."""
    y = np.asarray(y, dtype=np.float64)
    if basis is None:
        basis = mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)
    basis = np.asarray(basis, dtype=np.float64)
    pad = int((n_fft - hop_size) / 2)
    yp = np.pad(y, [[0, 0], [pad, pad]], mode="reflect")
    n = np.arange(win_size)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_size)          # periodic Hann (torch.hann_window)
    if win_size < n_fft:                                             # torch.stft centres a short window in n_fft
        left = (n_fft - win_size) // 2
        window = np.pad(window, [left, n_fft - win_size - left])
    frames = 1 + (yp.shape[1] - n_fft) // hop_size
    idx = np.arange(n_fft)[None, :] + hop_size * np.arange(frames)[:, None]
    seg = yp[:, idx] * window                                        # [B, frames, n_fft]
    spec = np.fft.rfft(seg, axis=-1)                                 # [B, frames, n_fft/2 + 1]
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)
    mel = np.einsum("mk,bfk->bmf", basis, mag)
    return np.log(np.maximum(mel, 1e-5))
