"""Micro-benchmark of the tcgen05 GEMM through the C ABI (spk_gemm): device time per launch, TFLOP/s, GB/s."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaker_embedding_torch_b200 import _native as N  # noqa: E402


def run(planes, m, n, k, a_mn=False, b_mn=False, out_f32=False, ksplit=1, block_n=0, bias_relu=False, iters=10):
    dev = "cuda"
    a = torch.randn((planes, k, m) if a_mn else (planes, m, k), device=dev).to(torch.float16)
    b = torch.randn((planes, k, n) if b_mn else (planes, n, k), device=dev).to(torch.float16)
    bias = torch.randn(n, device=dev) if bias_relu else None
    atomic = torch.zeros(m, n, device=dev) if ksplit > 1 else None
    for _ in range(3):
        N.gemm(a, b, planes, m, n, k, a_mn, b_mn, bias=bias, relu=bias_relu, out_f32=out_f32, atomic_out=atomic,
               ksplit=ksplit, block_n=block_n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        N.gemm(a, b, planes, m, n, k, a_mn, b_mn, bias=bias, relu=bias_relu, out_f32=out_f32, atomic_out=atomic,
               ksplit=ksplit, block_n=block_n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * m * n * k
    mma = flops * {1: 1, 2: 3, 3: 6}[planes]
    out_b = 4 if (out_f32 or ksplit > 1) else 2 * planes
    nbytes = (m * k + n * k) * 2 * planes + m * n * out_b
    print(json.dumps({"planes": planes, "m": m, "n": n, "k": k, "a_mn": a_mn, "b_mn": b_mn, "out": "f32" if out_b == 4 else "split",
                      "bn": block_n, "ksplit": ksplit, "ms": round(ms, 4), "alg_TFLOPs": round(flops / ms / 1e9, 1),
                      "mma_TFLOPs": round(mma / ms / 1e9, 1), "GBps": round(nbytes / ms / 1e6, 1)}))


if __name__ == "__main__":
    Mt = 153600
    run(1, 8192, 8192, 8192)                      # MMA-bound: what the mainloop can sustain
    run(1, 8192, 8192, 8192, block_n=128)
    run(2, 8192, 4096, 4096)
    run(3, 8192, 4096, 4096)
    for p in (1, 2, 3):
        run(p, Mt, 1024, 256, bias_relu=True)     # FFN1
        run(p, Mt, 1024, 256, out_f32=True)
        run(p, Mt, 256, 1024)                     # FFN2
        run(p, Mt, 768, 256)                      # QKV
        run(p, Mt, 256, 256)                      # out-proj
    run(2, Mt, 1024, 256, b_mn=True)              # ffn2 dgrad shape
    run(2, 1024, 256, Mt, a_mn=True, b_mn=True, ksplit=37)   # wgrad
    run(1, Mt, 1024, 256, block_n=128)
    run(1, Mt, 1024, 256, block_n=64)
