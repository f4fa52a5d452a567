"""dict -> nested ``argparse.Namespace`` (same contract as the reference's Arg_Parser.py:3-12)."""
from argparse import Namespace


def Recursive_Parse(args_Dict):
    return Namespace(**{k: Recursive_Parse(v) if isinstance(v, dict) else v for k, v in args_Dict.items()})


def default_hyper_parameters():
    """The encoder block of the reference Hyper_Parameters.yaml:1-18 (only what GE2E.__init__ reads)."""
    return Recursive_Parse({
        "Sound": {"Mel_Dim": 80},
        "GE2E": {
            "Embedding_Size": 256,
            "Positional_Encoding": {"Max_Position": 1024, "Dropout_Rate": 0.1},
            "Transformer": {"Num_Layers": 3, "Head": 4, "Dropout_Rate": 0.1},
        },
    })
