"""BASELINE config 5: fused GE2E forward+backward, N = 64..4096 speakers x 15 utterances x 256-d.
Prints one JSON line per N with the algorithmic roofline (bytes 2*N*M*D*4, FLOPs 6*N*M*N*D)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from speaker_embedding_torch_b200 import GE2E_Loss, _native  # noqa: E402

pk = bench.peaks()
crit = GE2E_Loss().cuda()
M, D = 15, 256
for N in (64, 128, 256, 512, 1024, 2048, 4096):
    torch.manual_seed(N)
    e = torch.nn.functional.normalize(torch.randn(N * M, D, device="cuda"), dim=1).requires_grad_(True)
    for _ in range(3):
        e.grad = None
        crit(e, M).backward()
    torch.cuda.synchronize()
    iters = 20 if N <= 1024 else 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        e.grad = None
        loss = crit(e, M)
        loss.backward()
    e1.record()
    torch.cuda.synchronize()
    wall_us = e0.elapsed_time(e1) / iters * 1e3
    # device time of the library's own launches (CUDA events around every kernel), without Python / autograd overhead
    _native.prof_enable(True)
    for _ in range(iters):
        e.grad = None
        crit(e, M).backward()
    torch.cuda.synchronize()
    rep = _native.prof_report()
    _native.prof_enable(False)
    us = sum(v["ms"] for k, v in rep.items() if k.startswith("ge2e")) / iters * 1e3
    nbytes, flops = 2.0 * N * M * D * 4, 6.0 * N * M * N * D
    roof_us = max(nbytes / (pk["hbm"] * 1e9), flops / (pk["tf_burst"] * 1e12)) * 1e6
    print(json.dumps({"N": N, "M": M, "us_per_fwd_bwd": round(us, 1), "loss": round(loss.item(), 5),
                      "alg_GBps": round(nbytes / us / 1e3, 1), "alg_TFLOPs": round(flops / us / 1e6, 2),
                      "roofline_us": round(roof_us, 2), "bound": "hbm" if nbytes / (pk["hbm"] * 1e9) > flops / (pk["tf_burst"] * 1e12) else "tensor",
                      "path": "fused SIMT kernel" if N < 256 else "tcgen05 GEMM composition",
                      "wall_us_incl_python": round(wall_us, 1), "kernels": sorted(k for k in rep if k.startswith("ge2e"))}))
