"""How does the tensor core round its fp32 accumulator?  One-plane fp16 GEMMs (operands exactly representable, so the
fp64 product is the exact answer), all-positive operands (no cancellation), growing K: prints the mean SIGNED relative
error and the rms relative error of the result.  Round-to-nearest accumulation gives a signed mean near 0 and an rms
that grows like sqrt(K); truncation gives a negative mean that grows like K.  Also: the same data through a 2-plane
product (3 MMAs per k-step).  Diagnostic, GPU box only: python tools/acc_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaker_embedding_torch_b200 import _native as N  # noqa: E402

torch.manual_seed(0)
m, n = 256, 256
for planes in (1, 2):
    for k in (16, 64, 256, 1024, 4096):
        for signed in (0, 1):
            A = torch.rand(m, k, device="cuda") + 0.5
            B = torch.rand(n, k, device="cuda") + 0.5
            if signed:
                A = A * torch.where(torch.rand_like(A) < 0.5, -1.0, 1.0)
            a_s, b_s = N.split_pack(A, planes), N.split_pack(B, planes)
            ref = N.split_unpack(a_s).double() @ N.split_unpack(b_s).double().t()
            out = N.gemm(a_s, b_s, planes, m, n, k, False, False, out_f32=True)
            torch.cuda.synchronize()
            f32 = (N.split_unpack(a_s) @ N.split_unpack(b_s).t()).double()        # cuBLAS fp32 (no TF32) for scale
            den = (N.split_unpack(a_s).double().abs() @ N.split_unpack(b_s).double().abs().t())
            e = (out.double() - ref) / den
            e32 = (f32 - ref) / den
            print("planes %d K %5d %s: tensor core  mean %+.2e rms %.2e   | torch fp32 mean %+.2e rms %.2e  (units of sum|a||b|; "
                  "2^-24 = 6.0e-08)" % (planes, k, "signed  " if signed else "positive", e.mean().item(),
                                        e.pow(2).mean().sqrt().item(), e32.mean().item(), e32.pow(2).mean().sqrt().item()))
