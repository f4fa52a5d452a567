"""CTA-pair (cta_group::2) GEMM vs the single-CTA kernel: bitwise comparison and timing (debug aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaker_embedding_torch_b200 import _native as N  # noqa: E402


def split(x, planes):
    out, r = [], x.float()
    for _ in range(planes):
        h = r.to(torch.float16)
        out.append(h)
        r = r - h.float()
    return torch.stack(out).contiguous()


def run(planes, m, n, k, b_mn=False, bias=False, relu=False, out_f32=False, reps=10):
    a = split(torch.randn(m, k, device="cuda"), planes)
    bw = torch.randn(k, n, device="cuda") if b_mn else torch.randn(n, k, device="cuda")
    b = split(bw, planes)
    bv = torch.randn(n, device="cuda") if bias else None
    res = {}
    for mode in (0, 1):
        N.set_option("gemm_cta_pairs", mode)
        for _ in range(2):
            out = N.gemm(a, b, planes, m, n, k, b_mn=b_mn, bias=bv, relu=relu, out_f32=out_f32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            out = N.gemm(a, b, planes, m, n, k, b_mn=b_mn, bias=bv, relu=relu, out_f32=out_f32)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = (out.float().clone(), e0.elapsed_time(e1) / reps)
    N.set_option("gemm_cta_pairs", 0)
    d = (res[0][0] - res[1][0]).abs().max().item()
    print("P%d %6d x %4d x %4d b_mn=%d bias=%d relu=%d f32=%d  single %.4f ms  pair %.4f ms  max|diff| %.3g" % (
        planes, m, n, k, b_mn, bias, relu, out_f32, res[0][1], res[1][1], d), flush=True)
    return d


if __name__ == "__main__":
    torch.manual_seed(0)
    run(3, 512, 256, 256)
    run(3, 1000, 768, 256, bias=True)
    run(2, 777, 1024, 256, b_mn=True, out_f32=True)
    run(3, 153600, 1024, 256, bias=True, relu=True)
    run(3, 153600, 256, 1024)
    run(3, 153600, 768, 256, bias=True)
    run(3, 153600, 256, 256)
    run(2, 153600, 1024, 256, b_mn=True)
    run(2, 153600, 256, 1024, b_mn=True)
