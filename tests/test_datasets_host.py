"""Host half of the device collater (speaker_embedding_torch_b200/Datasets.py) against the reference's numpy collater
(tests/golden/collate.npz, made by oracle/make_golden_collate.py from /root/reference/Datasets.py:9-19,72-86)."""
import os

import numpy as np
import pytest
import torch

from oracle.make_golden_collate import make_batch


def test_collater_draws_and_dense_equal_the_reference(golden_dir):
    from speaker_embedding_torch_b200.Datasets import Collater, RaggedMel
    g = np.load(os.path.join(golden_dir, "collate.npz"))
    for i in range(int(g["num_cases"])):
        seed, spk, utt, tmin, tmax, lo, hi = [int(v) for v in g["case%d_meta" % i]]
        batch = make_batch(seed, spk, utt, lo, hi)
        np.random.seed(seed)                               # same stream as the reference collater consumed
        ragged = Collater(tmin, tmax)(batch)
        assert isinstance(ragged, RaggedMel) and ragged.data.dtype == torch.float16
        want = g["case%d_out" % i]
        assert tuple(ragged.shape) == want.shape
        assert np.array_equal(ragged.dense().numpy(), want)           # bit-identical collation
        assert ragged.data.size(1) == sum(f.shape[1] for p in batch for f, _ in p)   # nothing padded on the host


def test_ragged_table_is_validated_on_the_host():
    from speaker_embedding_torch_b200.Datasets import RaggedMel
    data = torch.zeros(80, 10, dtype=torch.float16)
    RaggedMel(data, torch.tensor([[0, 4, 0], [4, 6, 1]], dtype=torch.int32), 5)
    with pytest.raises(RuntimeError):
        RaggedMel(data, torch.tensor([[0, 4, 0], [4, 7, 0]], dtype=torch.int32), 5)      # runs past the array
    with pytest.raises(RuntimeError):
        RaggedMel(data, torch.tensor([[0, 8, 4]], dtype=torch.int32), 5)                  # crop past the utterance
    with pytest.raises(RuntimeError):
        RaggedMel(data, torch.tensor([[0, 0, 0]], dtype=torch.int32), 5)                  # empty utterance
