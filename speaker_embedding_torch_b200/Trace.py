"""TorchScript export of the encoder: drop-in for the reference's ``Trace.py``.

Contract kept (Trace.py:7-44): ``Tracer(hp_path, checkpoint_path)``, ``forward(x, lengths)``, example
inputs ``x = rand(1, Mel_Dim, 400)``, ``lengths = LongTensor([400])``, output ``./traced/ge2e.pts``.
The reference passes ``lengths`` into ``GE2E.forward``'s ``samples`` slot, which cannot run
(SURVEY.md D7); here ``lengths`` is accepted and ignored (the collater crops every slice to one
length, Datasets.py:77-84, so the encoder never sees padding).

The traced graph holds one ``spkemb::encoder_infer`` node; import ``speaker_embedding_torch_b200``
before ``torch.jit.load`` so the op is registered.
"""
import argparse
import logging
import os

import torch
import yaml

from .Arg_Parser import Recursive_Parse
from .Modules import GE2E


class Tracer(torch.nn.Module):
    def __init__(self, hp_path: str, checkpoint_path: str, device="cuda"):
        super().__init__()
        self.hp = Recursive_Parse(yaml.load(open(hp_path, encoding="utf-8"), Loader=yaml.Loader))
        self.model = GE2E(self.hp)
        self.Load_Checkpoint(path=checkpoint_path)
        self.model.eval()
        for param in self.model.parameters():
            param.requires_grad = False
        self.model.to(device)

    def Load_Checkpoint(self, path):
        state_dict = torch.load(path, map_location="cpu")
        self.model.load_state_dict(state_dict["Model"])
        self.steps = state_dict["Steps"]
        logging.info("Checkpoint loaded at {} steps.".format(self.steps))

    def forward(self, x, lengths=None):
        return self.model(x)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-hp", "--hyper_parameters", required=True, type=str)
    parser.add_argument("-c", "--checkpoint_file", required=True, type=str)
    args = parser.parse_args(argv)
    tracer = Tracer(args.hyper_parameters, args.checkpoint_file)
    x = torch.rand(1, tracer.hp.Sound.Mel_Dim, 400, device="cuda")
    lengths = torch.LongTensor([400]).cuda()
    traced_model = torch.jit.trace(tracer, (x, lengths), check_trace=False)
    os.makedirs("traced", exist_ok=True)
    traced_model.save("./traced/ge2e.pts")
    return traced_model


if __name__ == "__main__":
    main()
