"""Generate tests/golden/mel.npz with the UNMODIFIED reference front-end (/root/reference/meldataset.py:73-96).

TEST INFRASTRUCTURE.  Authoring container only.  librosa is absent here, so ``librosa.filters.mel`` is provided by
oracle/mel_oracle.mel_filterbank (restated from librosa's published algorithm -- the one unpinned piece); the
reference's own ``mel_spectrogram`` then runs on seeded synthetic audio.

    python oracle/make_golden_mel.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("SPK_REFERENCE", "/root/reference")

from oracle import mel_oracle  # noqa: E402

CASES = [  # (seed, batch, samples)  -- Sound block of Hyper_Parameters.yaml: 1024 / 256 / 1024, 80 mels, 22050 Hz
    (1, 3, 256 * 40),
    (2, 1, 256 * 7),
    (3, 2, 22050),
]


def make_audio(seed, batch, samples):
    rng = np.random.default_rng(4000 + seed)
    t = np.arange(samples) / 22050.0
    f0 = rng.uniform(90.0, 300.0, size=(batch, 1))
    y = sum(rng.uniform(0.05, 0.3) * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28)) for h in range(1, 9))
    y = y + 0.02 * rng.standard_normal((batch, samples))
    y = y * np.minimum(1.0, np.linspace(0.0, 4.0, samples))[None, :]
    return (0.95 * y / np.abs(y).max(axis=1, keepdims=True)).astype(np.float32)


def main():
    lib = types.ModuleType("librosa")
    lib.util = types.ModuleType("librosa.util")
    lib.util.normalize = None
    lib.filters = types.ModuleType("librosa.filters")
    lib.filters.mel = lambda sr, n_fft, n_mels, fmin, fmax: mel_oracle.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    for m in (lib, lib.util, lib.filters):
        sys.modules[m.__name__] = m
    sys.path.insert(0, REF)
    import meldataset                      # the reference's own module
    out = {}
    for i, (seed, batch, samples) in enumerate(CASES):
        y = make_audio(seed, batch, samples)
        meldataset.mel_basis.clear()
        meldataset.hann_window.clear()
        mel = meldataset.mel_spectrogram(torch.from_numpy(y), 1024, 80, 22050, 256, 1024, 0, None, center=False)
        out["case%d_mel" % i] = mel.numpy()
        out["case%d_meta" % i] = np.array([seed, batch, samples], dtype=np.int64)
    out["num_cases"] = np.array(len(CASES))
    path = os.path.join(ROOT, "tests", "golden", "mel.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
