"""Speaker-verification metric for the d-vectors: trial lists, cosine scoring, equal error rate.

The reference publishes only t-SNE plots (README.md:97-113); an EER over same / different-speaker trials is the number a
d-vector model is normally judged by.  Everything here is plain tensor code on whatever device the embeddings live on
(scoring 10^6 trials of 256-d unit vectors is one gather and one row-wise dot product); it is not part of the hot path.
"""
import numpy as np
import torch


def make_trials(labels, num_trials, seed=0):
    """Balanced trial list from per-embedding speaker labels: ``(index_a, index_b, target)`` int64 arrays with as many
    target (same speaker, different utterance) as non-target trials (up to ``num_trials`` in total)."""
    labels = np.asarray(labels)
    _, inverse = np.unique(labels, return_inverse=True)
    rng = np.random.default_rng(seed)
    by_speaker = [np.nonzero(inverse == s)[0] for s in range(inverse.max() + 1)]
    multi = [ix for ix in by_speaker if ix.size >= 2]
    if not multi or len(by_speaker) < 2:
        raise RuntimeError("trials need at least two speakers and one speaker with two utterances")
    half = max(1, num_trials // 2)
    a, b, t = [], [], []
    for _ in range(half):
        ix = multi[rng.integers(len(multi))]
        i, j = rng.choice(ix, size=2, replace=False)
        a.append(i); b.append(j); t.append(1)
    for _ in range(half):
        s1, s2 = rng.choice(len(by_speaker), size=2, replace=False)
        a.append(rng.choice(by_speaker[s1])); b.append(rng.choice(by_speaker[s2])); t.append(0)
    return np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64), np.asarray(t, dtype=np.int64)


def cosine_scores(embeddings, index_a, index_b):
    """Cosine similarity of the trial pairs; ``embeddings`` [U, D] (unit norm or not)."""
    e = torch.nn.functional.normalize(embeddings.float(), dim=1)
    ia = torch.as_tensor(index_a, device=e.device)
    ib = torch.as_tensor(index_b, device=e.device)
    return (e[ia] * e[ib]).sum(dim=1)


def equal_error_rate(scores, targets):
    """EER and its threshold: the operating point where the false-rejection rate (targets scored below the threshold)
    equals the false-acceptance rate (non-targets at or above it), linearly interpolated between the two neighbouring
    thresholds.  Ties are handled by evaluating thresholds only between distinct score values."""
    s = torch.as_tensor(scores).detach().double().cpu().numpy()
    t = np.asarray(torch.as_tensor(targets).cpu().numpy()).astype(bool)
    n_tar, n_non = int(t.sum()), int((~t).sum())
    if n_tar == 0 or n_non == 0:
        raise RuntimeError("EER needs target and non-target trials")
    order = np.argsort(-s, kind="stable")                   # descending: accept the top-k
    s, t = s[order], t[order]
    tar_acc = np.cumsum(t)                                  # targets accepted when the top-k are accepted
    non_acc = np.cumsum(~t)
    last = np.r_[s[1:] != s[:-1], True]                     # only cut between distinct scores
    k = np.nonzero(last)[0]
    frr = np.r_[1.0, 1.0 - tar_acc[k] / n_tar]              # k = 0: accept nothing
    far = np.r_[0.0, non_acc[k] / n_non]
    thr = np.r_[np.inf, s[k]]
    d = far - frr                                           # increasing from -1 to +1
    j = int(np.argmax(d >= 0))
    if d[j] == 0 or j == 0:
        return float((far[j] + frr[j]) / 2), float(thr[j])
    w = -d[j - 1] / (d[j] - d[j - 1])                        # linear interpolation of the crossing
    eer = frr[j - 1] + w * (frr[j] - frr[j - 1])
    return float(eer), float(thr[j])


def evaluate_eer(embeddings, labels, num_trials=100000, seed=0):
    """EER of a set of d-vectors under a balanced random trial list."""
    a, b, t = make_trials(labels, num_trials, seed)
    return equal_error_rate(cosine_scores(embeddings, a, b), t)[0]
