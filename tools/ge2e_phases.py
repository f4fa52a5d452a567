"""Launch list target: the small-N GE2E loss (three stream-ordered stages) at N = 64 / 128, re-tiled stage and first stage.
Run under `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ge2e` to see the stages one by one."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speaker_embedding_torch_b200 import GE2E_Loss, _native

crit = GE2E_Loss().cuda()
for n in (64, 128):
    torch.manual_seed(n)
    e = torch.nn.functional.normalize(torch.randn(n * 15, 256, device="cuda"), dim=1).requires_grad_(True)
    for v2 in (1, 0):
        _native.set_option("ge2e_row_tile_v2", v2)
        for _ in range(4):
            e.grad = None
            crit(e, 15).backward()
        torch.cuda.synchronize()
_native.set_option("ge2e_row_tile_v2", 1)
print("done")
