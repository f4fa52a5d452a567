"""Versioned export of the encoder's flat weight arena.

``Train.py`` checkpoints are ``torch.save`` pickles of ``state_dict`` + optimiser state (Train.py:292-310); the traced
export (``Trace.py``) needs torch at load time.  For serving, the 43 parameters are also written as ONE flat fp32
arena -- the layout the kernels' gradient arena and the data-parallel broadcast already use -- behind a small
self-describing header:

    offset 0   magic   b"SPKEMBW\\0"
           8   u32     format version (FORMAT_VERSION)
          12   u32     length of the JSON header that follows
          16   JSON    {"config": {...encoder hyper-parameters...}, "steps": n, "dtype": "f32", "floats": total,
                        "crc32": of the payload, "tensors": [{"name", "shape", "offset", "numel"}, ...]}
          ..   zero padding to a multiple of 64 bytes
          ..   payload little-endian fp32, tensors in ``named_parameters()`` order, each 16-byte aligned

``load_arena`` checks magic, version and checksum and returns a ``state_dict`` (with the positional table rebuilt) that
loads with ``strict=True`` into this package's ``GE2E`` or the reference's.
"""
import json
import struct
import zlib

import numpy as np
import torch

MAGIC = b"SPKEMBW\0"
FORMAT_VERSION = 1


def _config_of(model):
    hp = model.hp
    return {"Mel_Dim": int(hp.Sound.Mel_Dim), "Embedding_Size": int(hp.GE2E.Embedding_Size),
            "Head": int(hp.GE2E.Transformer.Head), "Num_Layers": int(hp.GE2E.Transformer.Num_Layers),
            "Max_Position": int(hp.GE2E.Positional_Encoding.Max_Position),
            "PE_Dropout_Rate": float(hp.GE2E.Positional_Encoding.Dropout_Rate),
            "Dropout_Rate": float(hp.GE2E.Transformer.Dropout_Rate)}


def export_arena(model, path, steps=0):
    """Write ``model``'s parameters to ``path``; returns the header dict."""
    tensors, chunks, offset = [], [], 0
    for name, p in model.named_parameters():
        a = p.detach().to("cpu", torch.float32).contiguous().numpy().reshape(-1)
        tensors.append({"name": name, "shape": list(p.shape), "offset": offset, "numel": int(a.size)})
        pad = (-a.size) % 4
        chunks.append(a)
        if pad:
            chunks.append(np.zeros(pad, dtype=np.float32))
        offset += a.size + pad
    payload = np.concatenate(chunks).astype("<f4").tobytes()
    header = {"config": _config_of(model), "steps": int(steps), "dtype": "f32", "floats": offset,
              "crc32": zlib.crc32(payload) & 0xFFFFFFFF, "tensors": tensors}
    blob = json.dumps(header, sort_keys=True).encode("utf-8")
    head = MAGIC + struct.pack("<II", FORMAT_VERSION, len(blob)) + blob
    head += b"\0" * ((-len(head)) % 64)
    with open(path, "wb") as f:
        f.write(head)
        f.write(payload)
    return header


def load_arena(path):
    """-> (header dict, state_dict incl. the ``positional_encoding.pe`` buffer).  Raises ``RuntimeError`` on a foreign
    file, an unknown format version, a truncated payload or a checksum mismatch."""
    from .Modules import _sinusoid_table
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 16 or raw[:8] != MAGIC:
        raise RuntimeError("%s is not a speaker-embedding weight arena" % path)
    version, hlen = struct.unpack("<II", raw[8:16])
    if version != FORMAT_VERSION:
        raise RuntimeError("%s: arena format version %d, this build reads version %d" % (path, version, FORMAT_VERSION))
    header = json.loads(raw[16:16 + hlen].decode("utf-8"))
    start = 16 + hlen
    start += (-start) % 64
    payload = raw[start:]
    if len(payload) != 4 * header["floats"]:
        raise RuntimeError("%s: payload has %d bytes, header says %d" % (path, len(payload), 4 * header["floats"]))
    if (zlib.crc32(payload) & 0xFFFFFFFF) != header["crc32"]:
        raise RuntimeError("%s: checksum mismatch" % path)
    flat = np.frombuffer(payload, dtype="<f4")
    state = {}
    for t in header["tensors"]:
        state[t["name"]] = torch.from_numpy(flat[t["offset"]:t["offset"] + t["numel"]].reshape(t["shape"]).copy())
    cfg = header["config"]
    state["positional_encoding.pe"] = _sinusoid_table(cfg["Max_Position"], cfg["Embedding_Size"])
    return header, state


def model_from_arena(path, device=None):
    """A ``GE2E`` in eval mode built from an exported arena."""
    from .Arg_Parser import Recursive_Parse
    from .Modules import GE2E
    header, state = load_arena(path)
    c = header["config"]
    hp = Recursive_Parse({"Sound": {"Mel_Dim": c["Mel_Dim"]},
                          "GE2E": {"Embedding_Size": c["Embedding_Size"],
                                   "Positional_Encoding": {"Max_Position": c["Max_Position"],
                                                           "Dropout_Rate": c["PE_Dropout_Rate"]},
                                   "Transformer": {"Num_Layers": c["Num_Layers"], "Head": c["Head"],
                                                   "Dropout_Rate": c["Dropout_Rate"]}}})
    model = GE2E(hp)
    model.load_state_dict(state, strict=True)
    model.eval()
    return model.to(device) if device is not None else model
