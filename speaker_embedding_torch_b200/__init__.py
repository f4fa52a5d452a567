"""speaker_embedding_torch_b200 -- B200-native (sm_100a) speaker-embedding hot path.

Drop-in modules for CODEJIN/Speaker_Embedding_Torch's ``Modules.py`` / ``distributed.py`` /
``Trace.py`` backed by hand-written CUDA kernels behind the C ABI in ``include/spkemb.h``.
"""
from . import _native  # noqa: F401
from .Modules import GE2E, GE2E_Loss, Conv1d, Positional_Encoding  # noqa: F401
from .Arg_Parser import Recursive_Parse  # noqa: F401

__version__ = "0.1.0"
