"""Host -> device input pipeline helpers.

``DevicePrefetcher`` overlaps the pinned-host -> HBM copy of batch k+1 with the compute of batch k
(separate copy stream, double buffered), and ``AsyncScalar`` reads a device scalar (the loss) back
through pinned memory without stalling the launch queue.  Used by bench.py's end-to-end loop.
"""
import torch


class DevicePrefetcher(object):
    """Wrap an iterable of (pinned) host batches; yields device batches one step ahead.

    The device side is a ring of ``depth`` preallocated batches that are overwritten in place: no allocation happens
    in steady state.  (A fresh ``tensor.to(device)`` per step makes the caching allocator juggle 150 MB blocks between
    the copy stream and the compute stream; when it runs out of reusable blocks it falls back to ``cudaMalloc`` /
    ``cudaFree``, which synchronise the device - seen as sporadic +5 ms steps in the end-to-end bench.)  A slot is
    rewritten only after the work that was queued on the consumer's stream while it was current has been passed:
    the copy stream waits for an event recorded when the consumer asks for the next batch."""

    def __init__(self, loader, device=None, depth=2):
        self.loader = loader
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = max(2, int(depth))
        self._slots = [None] * self.depth
        self._sig = [None] * self.depth
        self._done = [None] * self.depth          # consumer finished (queued) with the slot
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]

    @staticmethod
    def _signature(batch):
        if torch.is_tensor(batch):
            return (tuple(batch.shape), batch.dtype)
        if isinstance(batch, (list, tuple)):
            return tuple(DevicePrefetcher._signature(b) for b in batch)
        if isinstance(batch, dict):
            return tuple((k, DevicePrefetcher._signature(v)) for k, v in sorted(batch.items()))
        return None

    def _alloc_like(self, batch):
        if torch.is_tensor(batch):
            return torch.empty(batch.shape, dtype=batch.dtype, device=self.device)
        if isinstance(batch, (list, tuple)):
            return type(batch)(self._alloc_like(b) for b in batch)
        if isinstance(batch, dict):
            return {k: self._alloc_like(v) for k, v in batch.items()}
        return batch

    def _copy_into(self, dst, src):
        if torch.is_tensor(src):
            dst.copy_(src, non_blocking=True)
        elif isinstance(src, (list, tuple)):
            for d, s_ in zip(dst, src):
                self._copy_into(d, s_)
        elif isinstance(src, dict):
            for k in src:
                self._copy_into(dst[k], src[k])

    def _prefetch(self, host, k):
        i = k % self.depth
        sig = self._signature(host)
        if self._slots[i] is None or self._sig[i] != sig:
            with torch.cuda.stream(self.stream):
                self._slots[i] = self._alloc_like(host)
            self._sig[i] = sig
        if self._done[i] is not None:
            self.stream.wait_event(self._done[i])
        with torch.cuda.stream(self.stream):
            self._copy_into(self._slots[i], host)
            self._ready[i].record(self.stream)
        return i

    def __iter__(self):
        it = iter(self.loader)
        try:
            host = next(it)
        except StopIteration:
            return
        k = 0
        nxt = self._prefetch(host, k)
        while nxt is not None:
            cur_stream = torch.cuda.current_stream(self.device)
            cur_stream.wait_event(self._ready[nxt])
            cur = nxt
            try:
                host = next(it)
                k += 1
                nxt = self._prefetch(host, k)
            except StopIteration:
                nxt = None
            yield self._slots[cur]
            # the consumer is back for the next batch: everything it queued on its stream so far used `cur`
            ev = self._done[cur] or torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._done[cur] = ev


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
