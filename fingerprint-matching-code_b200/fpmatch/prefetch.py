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

import os

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
    SLOTS = 4          # one being copied, up to three being matched (MatchingPipeline keeps 2-3 batches in flight)

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
        """Yields device batches.  The consumer may pull every batch under a different current stream (several batches
        in flight, see ``MatchingPipeline``): the batch is made visible to the stream that is current when it is pulled,
        and its buffers are released by an event recorded on that same stream when the next batch is pulled."""
        it = iter(self.batches)
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
            consumer = torch.cuda.current_stream(self.device)
            consumer.wait_event(ready)
            yield dev_batch
            done = torch.cuda.Event()                          # everything the consumer enqueued on this batch
            done.record(consumer)
            self.released[cur] = done


class HostResultRing:
    """Device -> host side of the pipeline: per-step result tensors are copied into pinned host buffers on a side
    stream; ``push`` returns the results of the step issued ``lag`` pushes earlier (complete on the host by then), so the
    CPU never stalls on a step it has just enqueued and stays ``lag`` steps ahead of the device.  ``flush`` returns the
    oldest outstanding step, or None.  (The reference reads results with blocking ``.cpu()`` / ``.item()`` calls right
    after the forward, e.g. evaluate_binary_classifier.py:93-103.)"""

    def __init__(self, device="cuda", slots: int = 2, lag: int = 1):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.lag = max(1, int(lag))
        self.slots = [dict(bufs=None, done=None) for _ in range(max(slots, self.lag + 1))]
        self.i = 0
        self.queue = []

    def set_lag(self, lag: int):
        assert not self.queue, "change the lag between runs only"
        self.lag = max(1, int(lag))
        while len(self.slots) < self.lag + 1:
            self.slots.append(dict(bufs=None, done=None))

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
        self.queue.append(slot)
        if len(self.queue) > self.lag:
            return self.flush()
        return None

    def flush(self):
        if not self.queue:
            return None
        slot = self.queue.pop(0)
        slot["done"].synchronize()
        return slot["bufs"]


_LANES = {}      # device index -> persistent compute streams shared by all pipelines on that device


def lanes(device, k: int):
    """The first ``k`` compute streams of the device's persistent set (a pipeline of depth 1 reuses the first stream
    of a pipeline of depth 2, so torch's per-stream caching allocator keeps serving the same blocks)."""
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    pool = _LANES.setdefault(key, [])
    while len(pool) < max(1, int(k)):
        pool.append(torch.cuda.Stream(device=device))
    return pool[:max(1, int(k))]


class MatchingPipeline:
    """Streams collated host batches through ``net`` with ``inflight`` batches on the device at once:

        for results in MatchingPipeline(net, batches, keys=("ds_mat", "perm_mat", "k_prob", "cls_prob")):
            ...            # list of pinned host tensors of one batch, in input order

    Batch i is matched on stream ``i % inflight`` (its kernels stay ordered on that stream; consecutive batches are
    independent), its inputs are copied one step ahead on the prefetcher's copy stream and its outputs are copied to
    pinned host memory on the result stream.  With two batches in flight the latency-bound tail of one batch (Sinkhorn,
    exact LAP, AFA-U) runs beside the tensor-bound SplineConv GEMMs of the next one: 8.05 -> 7.5 ms per 256-pair batch
    on B200 (r2k).  Outputs are bit-identical to one-batch-at-a-time execution (tests/test_gpu_head.py)."""

    def __init__(self, net, batches, keys=("ds_mat", "perm_mat", "k_prob", "cls_prob"), device="cuda", inflight=2,
                 feeder=None, ring=None, extra=None):
        self.net, self.keys, self.extra = net, tuple(keys), extra
        self.device = torch.device(device)
        self.feeder = feeder if feeder is not None else CudaPrefetcher(batches, device=self.device)
        if feeder is not None:
            self.feeder.batches = batches
        self.lanes = lanes(self.device, inflight)
        self.ring = ring if ring is not None else HostResultRing(device=self.device, lag=len(self.lanes))
        self.ring.set_lag(len(self.lanes))
        # stagger: batch i + 1's SplineConv front waits for batch i's front (Net.front_gate / front_done), so the tail of
        # one batch runs under the GEMMs of the next instead of both batches marching in lockstep
        self.stagger = os.environ.get("FPMATCH_STAGGER", "1") != "0"

    # of 74 SM pairs: the rest stay free for the other batch's tail kernels (r2n; FPMATCH_INFLIGHT_CLUSTERS overrides)
    GEMM_CLUSTERS_IN_FLIGHT = int(os.environ.get("FPMATCH_INFLIGHT_CLUSTERS", "70"))

    def __iter__(self):
        from . import ops
        cur = torch.cuda.current_stream(self.device)
        for s in self.lanes:
            s.wait_stream(cur)
        ops.set_gemm_max_clusters(self.GEMM_CLUSTERS_IN_FLIGHT if len(self.lanes) > 1 else 0)
        try:
            yield from self._run(cur)
        finally:
            ops.set_gemm_max_clusters(0)

    def _run(self, cur):
        it = iter(self.feeder)
        i = 0
        while True:
            with torch.cuda.stream(self.lanes[i % len(self.lanes)]):
                try:
                    d = next(it)
                except StopIteration:
                    break
                if self.extra:
                    d.update(self.extra)
                if self.stagger and len(self.lanes) > 1:
                    self.net.front_gate = getattr(self.net, "front_done", None) if i else None
                with torch.no_grad():
                    out = self.net(d)
                prev = self.ring.push([out[k] for k in self.keys])
            i += 1
            if prev is not None:
                yield prev
        for s in self.lanes:
            cur.wait_stream(s)
        while True:
            last = self.ring.flush()
            if last is None:
                break
            yield last
