"""Write-only / read-only / copy HBM bandwidth probe (torch fill_, sum, copy_), 2 GiB buffers."""
import torch

n = 1 << 29  # fp32 elements = 2 GiB
a = torch.empty(n, device="cuda", dtype=torch.float32)
b = torch.empty(n, device="cuda", dtype=torch.float32)


def timeit(fn, bytes_moved, name):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms  {bytes_moved / ms / 1e6:.0f} GB/s")


timeit(lambda: a.fill_(1.0), 4 * n, "write-only fill")
timeit(lambda: torch.cuda.memset if False else a.zero_(), 4 * n, "write-only zero")
timeit(lambda: a.sum(), 4 * n, "read-only sum")
timeit(lambda: b.copy_(a), 8 * n, "copy")
