"""Input pipeline helpers on the GPU: the ring-buffered prefetcher must hand out every batch exactly once, in order,
with the right contents, while copies run one step ahead on their own stream."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_prefetcher_ring_delivers_every_batch_in_order():
    from spark_deep_learning_b200.utils.data import AsyncScalar, DevicePrefetcher
    dev = torch.device("cuda", 0)
    batches = [(torch.full((64, 1024), float(i)).pin_memory(), {"y": torch.full((64,), i, dtype=torch.int64).pin_memory()})
               for i in range(7)]
    seen, ptrs = [], set()
    reader = AsyncScalar()
    for x, extra in DevicePrefetcher(batches, dev):
        assert x.device == dev and extra["y"].device == dev
        ptrs.add(x.data_ptr())
        # consume on the current stream with a kernel that takes a while, so that the next copy overlaps it
        acc = x
        for _ in range(20):
            acc = acc * 1.0
        reader.push(acc.mean() + extra["y"].float().mean())
        seen.append(None)
    vals = reader.drain()
    torch.cuda.synchronize()
    assert len(seen) == 7 and len(ptrs) == 2            # two ring slots, reused in place
    assert vals[-2:] == [2.0 * 5, 2.0 * 6]


def test_device_prefetcher_handles_a_shape_change():
    from spark_deep_learning_b200.utils.data import DevicePrefetcher
    dev = torch.device("cuda", 0)
    batches = [torch.full((8, 4), 1.0).pin_memory(), torch.full((8, 4), 2.0).pin_memory(), torch.full((3, 4), 3.0).pin_memory()]
    out = [b.sum().item() for b in DevicePrefetcher(batches, dev)]
    assert out == [32.0, 64.0, 36.0]
