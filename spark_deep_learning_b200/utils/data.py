"""Host -> device input pipeline helpers.

``DevicePrefetcher`` overlaps the pinned-host -> HBM copy of batch k+1 with the compute of batch k
(separate copy stream, double buffered), and ``AsyncScalar`` reads a device scalar (the loss) back
through pinned memory without stalling the launch queue.  Used by bench.py's end-to-end loop.
"""
import torch


class DevicePrefetcher(object):
    """Wrap an iterable of (pinned) host batches; yields device batches one step ahead."""

    def __init__(self, loader, device=None):
        self.loader = loader
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)

    def _to_device(self, batch):
        if torch.is_tensor(batch):
            return batch.to(self.device, non_blocking=True)
        if isinstance(batch, (list, tuple)):
            return type(batch)(self._to_device(b) for b in batch)
        if isinstance(batch, dict):
            return {k: self._to_device(v) for k, v in batch.items()}
        return batch

    def _record(self, batch):
        cur = torch.cuda.current_stream(self.device)
        if torch.is_tensor(batch):
            batch.record_stream(cur)
        elif isinstance(batch, (list, tuple)):
            for b in batch:
                self._record(b)
        elif isinstance(batch, dict):
            for b in batch.values():
                self._record(b)

    def __iter__(self):
        it = iter(self.loader)
        nxt = None
        try:
            host = next(it)
        except StopIteration:
            return
        with torch.cuda.stream(self.stream):
            nxt = self._to_device(host)
        while nxt is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            cur = nxt
            self._record(cur)
            try:
                host = next(it)
                with torch.cuda.stream(self.stream):
                    nxt = self._to_device(host)
            except StopIteration:
                nxt = None
            yield cur


class AsyncScalar(object):
    """Deferred device->host read of a scalar: ``push(t)`` starts the copy, ``pop()`` returns the
    value pushed the step before (its copy has long finished), ``drain()`` the remaining ones."""

    def __init__(self, depth=2):
        self._slots = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self._events = [torch.cuda.Event() for _ in range(depth)]
        self._n = 0
        self._read = 0
        self.nbytes = 4

    def push(self, t):
        i = self._n % len(self._slots)
        if self._n - self._read >= len(self._slots):
            self.pop()
        self._slots[i].copy_(t.detach().float().reshape(1), non_blocking=True)
        self._events[i].record()
        self._n += 1

    def pop(self):
        if self._read >= self._n:
            return None
        i = self._read % len(self._slots)
        self._events[i].synchronize()
        self._read += 1
        return float(self._slots[i][0])

    def drain(self):
        vals = []
        while self._read < self._n:
            vals.append(self.pop())
        return vals
