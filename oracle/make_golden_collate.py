"""Generate tests/golden/collate.npz with the UNMODIFIED reference collater (Datasets.py:9-19,72-86).

TEST INFRASTRUCTURE.  Authoring container only (needs /root/reference).  The reference's Datasets.py imports
Pattern_Generator -> librosa / pysptk, which this image does not have; they are stubbed in sys.modules (nothing on the
collater's path calls into them).  Inputs are regenerated from the seeds at test time.

    python oracle/make_golden_collate.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("SPK_REFERENCE", "/root/reference")

CASES = [  # (seed, speakers, utterances per speaker, min_frame_length, max_frame_length, min L, max L)
    (1, 4, 3, 40, 48, 20, 90),
    (2, 3, 2, 140, 180, 1, 400),        # very short patterns: reflect padding wraps around several times
    (3, 5, 2, 64, 64, 64, 65),          # L == T (neither branch crops) and L == T + 1
]


def make_batch(seed, speakers, utts, lo, hi):
    rng = np.random.default_rng(1000 + seed)
    batch = []
    for s in range(speakers):
        pattern = []
        for u in range(utts):
            length = int(rng.integers(lo, hi + 1))
            mel = np.clip(-5.0 + 2.0 * rng.standard_normal((80, length)), np.log(1e-5), 2.0).astype(np.float16)
            pattern.append((mel, "S%d" % s))
        batch.append(pattern)
    return batch


def main():
    for name in ("librosa", "librosa.util", "librosa.filters", "pysptk", "pysptk.sptk"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["librosa.util"].normalize = None
    sys.modules["librosa.filters"].mel = None
    sys.modules["librosa"].util = sys.modules["librosa.util"]
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    sys.modules["pysptk.sptk"].rapt = None
    sys.path.insert(0, REF)
    from Datasets import Collater        # the reference's own collater
    out = {}
    for i, (seed, spk, utt, tmin, tmax, lo, hi) in enumerate(CASES):
        batch = make_batch(seed, spk, utt, lo, hi)
        np.random.seed(seed)
        feats = Collater(tmin, tmax)(batch)
        out["case%d_out" % i] = feats.numpy()
        out["case%d_meta" % i] = np.array([seed, spk, utt, tmin, tmax, lo, hi], dtype=np.int64)
    out["num_cases"] = np.array(len(CASES))
    path = os.path.join(ROOT, "tests", "golden", "collate.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
