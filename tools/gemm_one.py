"""One GEMM shape repeated a few times (ncu target): python tools/gemm_one.py planes m n k [block_n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaker_embedding_torch_b200 import _native as N  # noqa: E402

planes, m, n, k = [int(v) for v in sys.argv[1:5]]
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
a = torch.randn(planes, m, k, device="cuda").to(torch.float16)
b = torch.randn(planes, n, k, device="cuda").to(torch.float16)
bias = torch.randn(n, device="cuda")
for _ in range(5):
    out = N.gemm(a, b, planes, m, n, k, bias=bias, relu=True, block_n=bn)
torch.cuda.synchronize()
print(float(out.float().sum()))
