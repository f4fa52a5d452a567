"""Device_Prefetcher: host -> device double buffering used by the end-to-end training loop."""
import pytest
import torch

from speaker_embedding_torch_b200.Prefetch import Device_Prefetcher


def test_prefetcher_refuses_cpu():
    with pytest.raises(RuntimeError):
        Device_Prefetcher([torch.zeros(2)], "cpu")


@pytest.mark.gpu
def test_prefetcher_order_and_values():
    dev = torch.device("cuda", 0)
    host = [torch.full((3, 1 << 18), float(i)).pin_memory() for i in range(7)]
    seen = []
    for i, t in enumerate(Device_Prefetcher(host, dev)):
        assert t.is_cuda and t.shape == host[i].shape
        seen.append(float((t * 2).sum().item()) / (2 * t.numel()))   # consume on the compute stream
    assert seen == [float(i) for i in range(7)]


@pytest.mark.gpu
def test_prefetcher_nested_batches():
    dev = torch.device("cuda", 0)
    host = [(torch.ones(4).pin_memory() * i, {"n": torch.tensor([i])}) for i in range(3)]
    out = list(Device_Prefetcher(host, dev, depth=2))
    assert len(out) == 3
    for i, (a, d) in enumerate(out):
        assert a.is_cuda and d["n"].is_cuda
        assert float(a[0].item()) == float(i) and int(d["n"].item()) == i
