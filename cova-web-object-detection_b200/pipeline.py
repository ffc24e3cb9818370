"""Input staging for the hot path (the four `.to(device)` of
`/root/reference/train.py:48-51`): at native-kernel speed the 19.7 MB/page fp32 image copy over PCIe costs more
than the forward itself, so the copy of batch i+1 must overlap the compute of batch i.

`prefetch(batches, device)` wraps any iterator of collated host batches (tuples of CPU tensors, e.g. what
`custom_collate_fn` yields) and yields the same tuples on the device: tensors are staged through pinned host
buffers and copied on a side stream one batch ahead; the consumer's stream waits on the copy's event, so results
are identical to the synchronous `.to(device)` calls.  Optional - the stock loop keeps working without it.
"""
import os

import torch


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index` (one process per GPU): pinned
    host buffers are then first-touched on the GPU's own NUMA node, so on a two-socket box the host->device copies
    of 8 ranks do not cross the inter-socket link.  Returns the core list, or None when NVML / affinity is unavailable
    (nothing changes then).  Call it before allocating pinned memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        cores = [c for c in cores if c in os.sched_getaffinity(0)]
        if cores:
            os.sched_setaffinity(0, cores)
            return cores
    except Exception:
        pass
    return None


_COPY_STREAMS = {}


class _Slot:
    def __init__(self):
        self.pinned, self.dev, self.event = None, None, None


def _stage(slot, batch, device, stream):
    # The slot's pinned staging buffers are about to be overwritten by the HOST: the asynchronous H2D copy issued from
    # them for an earlier batch must have completed (a stream-side wait_event only orders the GPU work and may leave
    # that copy still queued when the host runs ahead, e.g. an inference loop with no per-batch sync).
    if slot.event is not None:
        slot.event.synchronize()
    with torch.cuda.stream(stream):
        # The device buffers are first written on the COPY stream, so they must come from that stream's pool:
        # a block the caching allocator hands out for the compute stream may still be in use by kernels that are
        # queued there (stream-ordered reuse is only safe on the stream that freed it).
        if slot.dev is None or any(d.shape != t.shape or d.dtype != t.dtype for d, t in zip(slot.dev, batch)):
            slot.dev = [torch.empty(t.shape, dtype=t.dtype, device=device) for t in batch]
            slot.pinned = [None if t.is_pinned() else torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in batch]
        for d, p, t in zip(slot.dev, slot.pinned, batch):
            if p is not None:
                p.copy_(t)
            d.copy_(t if p is None else p, non_blocking=True)
        slot.event = torch.cuda.Event()
        slot.event.record(stream)


def prefetch(batches, device, depth=2):
    """Yield device copies of `batches` (iterable of tuples of CPU tensors), copying one batch ahead on a side
    stream.  A yielded tuple stays valid until `depth` further batches have been requested."""
    device = torch.device(device)
    # ONE copy stream per device for the life of the process: the caching allocator pools blocks per stream, so a
    # fresh stream per call meant ~1 GB of cudaMalloc (and, now and then, a cudaFree + device sync of blocks cached
    # for a dead stream) inside the first steps of every call - seen as 2-3x outliers of a 120 ms measurement.
    key = device.index if device.index is not None else torch.cuda.current_device()
    copy_stream = _COPY_STREAMS.get(key)
    if copy_stream is None:
        copy_stream = _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    slots = [_Slot() for _ in range(depth + 1)]
    done = [None] * (depth + 1)          # consumer-side event: the slot's previous contents are no longer read
    it = iter(batches)
    pending = []
    i = 0

    def issue():
        nonlocal i
        try:
            batch = next(it)
        except StopIteration:
            return False
        s = i % len(slots)
        if done[s] is not None:
            copy_stream.wait_event(done[s])
        tensors = [t for t in batch if isinstance(t, torch.Tensor)]
        _stage(slots[s], tensors, device, copy_stream)
        pending.append((s, batch))
        i += 1
        return True

    for _ in range(depth):
        if not issue():
            break
    while pending:
        s, batch = pending.pop(0)
        compute = torch.cuda.current_stream(device)
        compute.wait_event(slots[s].event)
        for t in slots[s].dev:
            t.record_stream(compute)     # freed blocks must also wait for the consumer's kernels
        dev_iter = iter(slots[s].dev)
        yield tuple(next(dev_iter) if isinstance(t, torch.Tensor) else t for t in batch)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        done[s] = ev
        issue()
