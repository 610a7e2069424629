"""Host->device input pipeline: the next batch is copied on a side stream while the current one is matched.

The reference moves every batch synchronously right before the forward (``batch = data_to_cuda(batch)`` then
``model(batch)``, evaluate_binary_classifier.py:86-92, src/train/training_loop.py:24-32), so the 0.94 MB of backbone
feature maps per pair (or the images) cross PCIe while the GPU idles.  ``CudaPrefetcher`` wraps any iterable of
collated batches (pinned host tensors) and yields device batches one step ahead; at 256 pairs x 100 keypoints the
262 MB copy (~10 ms) hides behind the ~16 ms of head compute.

The device copies live in two persistent buffer sets that are reused round-robin (no allocator traffic in steady
state); CUDA events order "copy into set j" after "the step that last read set j has finished".
"""
from __future__ import annotations

import torch

from utils.data_to_cuda import data_to_cuda


def _shallow(batch):
    """Containers are copied (data_to_cuda updates them in place); leaves are shared."""
    if isinstance(batch, dict):
        return {k: _shallow(v) for k, v in batch.items()}
    if isinstance(batch, list):
        return [_shallow(v) for v in batch]
    if isinstance(batch, tuple):
        return tuple(_shallow(v) for v in batch)
    return batch


class CudaPrefetcher:
    SLOTS = 2

    def __init__(self, batches, device="cuda"):
        self.batches = batches
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.pools = [[] for _ in range(self.SLOTS)]           # per slot: device buffers in traversal order
        self.released = [None] * self.SLOTS                    # event: the consumer is done with the slot

    def _stage(self, host_batch, slot):
        pool = self.pools[slot]
        cursor = [0]

        def mover(t):
            i = cursor[0]
            cursor[0] += 1
            if i < len(pool) and pool[i].shape == t.shape and pool[i].dtype == t.dtype:
                buf = pool[i]
            else:
                buf = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                if i < len(pool):
                    pool[i] = buf
                else:
                    pool.append(buf)
            buf.copy_(t, non_blocking=True)
            return buf

        with torch.cuda.stream(self.copy_stream):
            if self.released[slot] is not None:
                self.copy_stream.wait_event(self.released[slot])
            dev_batch = data_to_cuda(_shallow(host_batch), device=self.device, mover=mover)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return dev_batch, ready

    def __iter__(self):
        it = iter(self.batches)
        main = torch.cuda.current_stream(self.device)
        slot = 0
        try:
            staged = self._stage(next(it), slot)
        except StopIteration:
            return
        while staged is not None:
            dev_batch, ready = staged
            cur = slot
            slot = (slot + 1) % self.SLOTS
            try:
                staged = self._stage(next(it), slot)           # the NEXT batch starts crossing PCIe now
            except StopIteration:
                staged = None
            main.wait_event(ready)
            yield dev_batch
            done = torch.cuda.Event()                          # everything the consumer enqueued on this batch
            done.record(main)
            self.released[cur] = done


class HostResultRing:
    """Device -> host side of the pipeline: per-step result tensors are copied into pinned host buffers on a side
    stream; ``push`` returns the PREVIOUS step's results (now complete on the host), so the CPU never stalls on the
    step it has just enqueued.  ``flush`` returns the last step's results.  (The reference reads results with
    blocking ``.cpu()`` / ``.item()`` calls right after the forward, e.g. evaluate_binary_classifier.py:93-103.)"""

    def __init__(self, device="cuda", slots: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict(bufs=None, done=None) for _ in range(slots)]
        self.i = 0
        self.pending = None

    def push(self, tensors):
        slot = self.slots[self.i % len(self.slots)]
        self.i += 1
        if slot["bufs"] is None or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(slot["bufs"], tensors)):
            slot["bufs"] = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
        main = torch.cuda.current_stream(self.device)
        produced = torch.cuda.Event()
        produced.record(main)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(produced)
            for b, t in zip(slot["bufs"], tensors):
                b.copy_(t, non_blocking=True)
                t.record_stream(self.stream)
            slot["done"] = torch.cuda.Event()
            slot["done"].record(self.stream)
        prev, self.pending = self.pending, slot
        if prev is not None:
            prev["done"].synchronize()
            return prev["bufs"]
        return None

    def flush(self):
        if self.pending is None:
            return None
        self.pending["done"].synchronize()
        out, self.pending = self.pending["bufs"], None
        return out
