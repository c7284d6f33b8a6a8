"""Double-buffered host -> device feed for full-graph training steps.

The layer's input `x [N, F*d]` fp32 is 5.5 GB at the ogbn-arxiv shape: uploaded serially it costs about as long as
the layer's forward + backward.  `HostFeed` keeps two device copies of every fed tensor and uploads batch i+1 from
pinned host memory on its own copy stream while the caller's stream works on batch i -- the PCIe transfer and the
kernels overlap, nothing is skipped.  (The reference has no such stage: its scripts keep the Cora tensors on one
device for the whole run, `experiments/cora_benchmark_graphsaint.py:30-57`.)

    feed = HostFeed(device)
    feed.submit(x_host, ei_host)              # upload of step 0
    for i in range(steps):
        x, ei = feed.get()                    # the caller's stream now waits for that upload only
        if i + 1 < steps:
            feed.submit(x_host, ei_host)      # step i+1 travels while step i computes
        ...forward / backward on x, ei...
        feed.release()                        # step i no longer reads its buffers

Plain PyTorch plumbing (streams, events, pinned copies); no kernels of its own.
"""
from collections import deque

import torch

__all__ = ["HostFeed"]


class HostFeed:
    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise TypeError("HostFeed uploads to a CUDA device; ampnet_b200 has no CPU path")
        if depth < 2:
            raise ValueError("depth must be at least 2 (one buffer set in use, one in flight)")
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._buffers = [None] * depth      # per slot: list of device tensors
        self._free = [None] * depth         # per slot: event recorded by release() on the consumer's stream
        self._ready = [None] * depth        # per slot: event recorded on the copy stream after the uploads
        self._next = 0                      # slot the next submit() writes
        self._pending = deque()             # slots submitted and not yet handed out
        self._in_use = deque()              # slots handed out and not yet released
        self.bytes_submitted = 0

    def _slot_buffers(self, slot, host_tensors):
        bufs = self._buffers[slot]
        ok = bufs is not None and len(bufs) == len(host_tensors) and all(
            b.shape == h.shape and b.dtype == h.dtype for b, h in zip(bufs, host_tensors))
        if not ok:
            bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors]
            self._buffers[slot] = bufs
            # fresh allocations were made on the caller's stream: order the copy stream behind them
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        return bufs

    def submit(self, *host_tensors):
        """Starts the upload of one batch (pinned host tensors) into the next free buffer set; returns immediately."""
        if len(self._pending) + len(self._in_use) >= self.depth:
            raise RuntimeError("HostFeed: every buffer set is in flight or in use; call get()/release() first")
        for h in host_tensors:
            if h.device.type != "cpu" or not h.is_pinned():
                raise TypeError("HostFeed.submit expects pinned host tensors (tensor.pin_memory())")
        slot = self._next
        self._next = (slot + 1) % self.depth
        bufs = self._slot_buffers(slot, host_tensors)
        if self._free[slot] is not None:
            self.copy_stream.wait_event(self._free[slot])     # the step that last read this set has finished
        with torch.cuda.stream(self.copy_stream):
            for b, h in zip(bufs, host_tensors):
                b.copy_(h, non_blocking=True)
                self.bytes_submitted += h.numel() * h.element_size()
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._ready[slot] = ev
        self._pending.append(slot)

    def get(self):
        """Hands out the oldest submitted batch; the caller's current stream waits for exactly that upload."""
        if not self._pending:
            raise RuntimeError("HostFeed.get() without a submitted batch")
        slot = self._pending.popleft()
        torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        self._in_use.append(slot)
        return tuple(self._buffers[slot])

    def release(self):
        """Marks the oldest handed-out batch as no longer read by the caller's current stream."""
        if not self._in_use:
            raise RuntimeError("HostFeed.release() without a batch in use")
        slot = self._in_use.popleft()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[slot] = ev
