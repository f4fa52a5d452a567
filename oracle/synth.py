"""Deterministic synthetic weights and inputs shared by the oracle, the tests and bench.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Everything is generated from
numpy's PCG64 so that the same seed gives the same tensors in the authoring
container (where the golden vectors are made with the reference) and on the
GPU box (where /root/reference does not exist).

State-dict names and shapes follow the reference model (SURVEY.md Appendix A;
/root/reference/Modules.py:6-44).  Every tensor is randomised, including the
biases / LayerNorm affines that the reference initialises to 0 / 1 and the
three encoder layers that ``nn.TransformerEncoder`` deep-copies to identical
values -- otherwise a parity test could not see a swapped
layer index or a dropped bias.
"""
import math

import numpy as np

MEL_DIM = 80
EMB = 256
HEADS = 4
FFN = 1024
LAYERS = 3
MAX_POS = 1024


def state_shapes(mel_dim=MEL_DIM, emb=EMB, ffn=FFN, layers=LAYERS, max_pos=MAX_POS):
    """Ordered (name, shape) list == reference ``GE2E.state_dict()`` (Modules.py:10-44)."""
    out = [
        ("prenet.weight", (emb, mel_dim, 1)),
        ("prenet.bias", (emb,)),
        ("positional_encoding.alpha", (1,)),
        ("positional_encoding.pe", (1, emb, max_pos)),
    ]
    for l in range(layers):
        p = "transformer.layers.%d." % l
        out += [
            (p + "self_attn.in_proj_weight", (3 * emb, emb)),
            (p + "self_attn.in_proj_bias", (3 * emb,)),
            (p + "self_attn.out_proj.weight", (emb, emb)),
            (p + "self_attn.out_proj.bias", (emb,)),
            (p + "linear1.weight", (ffn, emb)),
            (p + "linear1.bias", (ffn,)),
            (p + "linear2.weight", (emb, ffn)),
            (p + "linear2.bias", (emb,)),
            (p + "norm1.weight", (emb,)),
            (p + "norm1.bias", (emb,)),
            (p + "norm2.weight", (emb,)),
            (p + "norm2.bias", (emb,)),
        ]
    out += [
        ("transformer.norm.weight", (emb,)),
        ("transformer.norm.bias", (emb,)),
        ("projection.weight", (emb, emb, 1)),
        ("projection.bias", (emb,)),
    ]
    return out


def positional_table(max_pos=MAX_POS, emb=EMB):
    """The ``pe`` buffer, [1, emb, max_pos] fp32 (Modules.py:86-92).

    Computed in fp32 with the same operation order as the reference
    (position * exp(arange * (-ln(1e4)/emb)), then sin / cos in fp32).
    """
    import torch  # torch's fp32 sin/cos/exp are what the reference buffer holds

    pe = torch.zeros(max_pos, emb)
    position = torch.arange(0, max_pos, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, emb, 2).float() * (-math.log(10000.0) / emb))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(2, 1).contiguous().numpy()


def make_state(seed=0, scale=1.0):
    """Random but well-conditioned encoder state (dict name -> fp32 ndarray).

    Weights ~ N(0, 1/fan_in) * scale, biases ~ N(0, 0.1^2), LayerNorm weights
    1 + N(0, 0.1^2), alpha = 1 + N(0, 0.1^2).
    """
    rng = np.random.default_rng(seed)
    state = {}
    for name, shape in state_shapes():
        if name.endswith(".pe"):
            state[name] = positional_table()
            continue
        if name.endswith("alpha"):
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif "norm" in name and name.endswith("weight"):
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith("bias"):
            v = 0.1 * rng.standard_normal(shape)
        else:
            fan_in = shape[1]
            v = rng.standard_normal(shape) * (scale / math.sqrt(fan_in))
        state[name] = v.astype(np.float32)
    return state


def make_mel(seed, batch, frames, mel_dim=MEL_DIM):
    """Log-mel-like input [batch, mel_dim, frames] fp32.

    The reference's features are log(clamp(mel, 1e-5)) (meldataset.py:51-52), so
    values live in [ln 1e-5, ~2].
    """
    rng = np.random.default_rng(seed)
    x = -5.0 + 2.0 * rng.standard_normal((batch, mel_dim, frames))
    return np.clip(x, math.log(1e-5), 2.0).astype(np.float32)


def make_embeddings(seed, speakers, utterances, emb=EMB, unit_norm=True, spread=0.35):
    """Clustered embeddings [speakers*utterances, emb] fp32, speaker-major rows."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((speakers, 1, emb))
    e = centers + spread * math.sqrt(emb) * rng.standard_normal((speakers, utterances, emb)) / math.sqrt(emb) * 3.0
    e = e.reshape(speakers * utterances, emb)
    if unit_norm:
        e = e / np.linalg.norm(e, axis=1, keepdims=True)
    else:
        e = e * (0.5 + rng.random((speakers * utterances, 1)))
    return e.astype(np.float32)
