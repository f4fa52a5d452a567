import time
import torch
dev = torch.device("cuda", 0)
src = torch.randn(960, 80, 160).pin_memory()
n = src.numel()
dst32 = torch.empty_like(src, device=dev)
buf = torch.empty(n * 4, dtype=torch.uint8, device=dev)
view = buf[: n * 4].view(torch.float32).view(src.shape)
side = torch.cuda.Stream(device=dev)
a = torch.randn(8192, 8192, device=dev)
torch.cuda.synchronize()


def t(name, fn, busy=False):
    torch.cuda.synchronize()
    if busy:
        for _ in range(40):
            a @ a            # ~30 ms of queued compute on the current stream
    t0 = time.perf_counter()
    fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("%-44s cpu %.3f ms  (busy=%s)" % (name, (t1 - t0) * 1e3, busy))


def on_side(fn):
    def g():
        with torch.cuda.stream(side):
            fn()
    return g


for busy in (False, True):
    t("to() current stream", lambda: src.to(dev, non_blocking=True), busy)
    t("copy_ fp32 dst, current stream", lambda: dst32.copy_(src, non_blocking=True), busy)
    t("copy_ uint8-view dst, current stream", lambda: view.copy_(src, non_blocking=True), busy)
    t("copy_ fp32 dst, side stream", on_side(lambda: dst32.copy_(src, non_blocking=True)), busy)
    t("copy_ uint8-view dst, side stream", on_side(lambda: view.copy_(src, non_blocking=True)), busy)
    t("to() side stream", on_side(lambda: src.to(dev, non_blocking=True)), busy)
