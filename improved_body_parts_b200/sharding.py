"""Image sharding across ranks and the gather of the person lists (BASELINE.json north_star: "images shard
embarrassingly across the 8 GPUs of one box with a gather of the final person lists over NVLink").

The path has no exchange step during compute: images are independent.  Rank r owns the contiguous
block ``shard_range(n, r, world)``; what has to reach the consumer (rank ``dst``) are the wire records of every
image (wire.py: the ``format_results`` payload), in rank order = image order, so an N-GPU run returns exactly what a
1-GPU run returns.  Two transports, same records, same layout on ``dst``:

``PeerWireSink``  the B200 path.  ``dst`` owns one sink buffer, every rank maps it (CUDA IPC over NVLink/NVSwitch) and
                  its assemble kernel stores each image's record STRAIGHT into its slice of that buffer: the gather is
                  fused into the kernel that produces the records -- no collective kernel, nothing that competes with
                  the all-SM persistent kernels, only live person rows cross the links.  Hand-shakes are 64-bit
                  counters: a producer release-stores "step s landed" into ``dst``'s memory after its kernels, ``dst``
                  waits for the counters with stream memory operations (no SM-resident polling) and, once it has
                  consumed a step, acknowledges into each producer's mailbox so the (double-buffered) slice can be reused.
``PackedGather``  the portable path (gloo in the CPU tests, NCCL where peer mapping is unavailable): ONE collective
                  per step on one pre-allocated packed buffer -- no per-step allocation, no concatenation.

``torch.distributed`` is plumbing (rendezvous, the exchange of IPC handles, the fallback collective).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``; earlier ranks take the remainder one image each."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None) -> Tuple[int, int]:
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def exchange_counts(n_local: int, group=None) -> List[int]:
    """Every rank's image count, in rank order (shards may differ by one image, or arbitrarily)."""
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return [int(n_local)]
    out: List[Optional[int]] = [None] * world
    dist.all_gather_object(out, int(n_local), group=group)
    return [int(v) for v in out]


class PackedGather:
    """One collective per step: equally sized packed record buffers gathered into one pre-allocated buffer on ``dst``.

    Uneven shards are padded to the largest shard (sizes are exchanged once, here); ``records()`` trims the padding.
    ``local`` is the ``[max_n, record_bytes]`` uint8 tensor the rank's grouper writes its records into
    (``Grouper.set_wire_output(local.data_ptr())``) -- or that a test fills by hand."""

    def __init__(self, n_local: int, record_bytes: int, device, dst: int = 0, group=None):
        import torch
        self.rank, self.world = _world(group)
        self.group, self.dst, self.record_bytes = group, dst, int(record_bytes)
        self.counts = exchange_counts(n_local, group)
        self.max_n = max(self.counts) if self.counts else 0
        self.local = torch.zeros((self.max_n, self.record_bytes), dtype=torch.uint8, device=device)
        self.out = torch.zeros((self.world, self.max_n, self.record_bytes), dtype=torch.uint8, device=device) \
            if self.rank == dst else None
        self._views = list(self.out.unbind(0)) if self.out is not None else None  # views, not copies: gather writes in place

    def gather(self):
        """Issue the collective on the current stream; returns the ``[world, max_n, record_bytes]`` buffer on ``dst``."""
        import torch.distributed as dist
        if self.world == 1:
            self.out.copy_(self.local[None])
            return self.out
        dist.gather(self.local, self._views, dst=self.dst, group=self.group)
        return self.out

    def records(self):
        """``dst`` only: the ``[sum(counts), record_bytes]`` records in image order (a copy only if shards are uneven)."""
        import torch
        if self.out is None:
            return None
        if all(c == self.max_n for c in self.counts):
            return self.out.reshape(-1, self.record_bytes)
        return torch.cat([self.out[r, :c] for r, c in enumerate(self.counts)], dim=0)


class PeerWireSink:
    """Records travel as the assemble kernel stores them: every rank writes into its slice of ``dst``'s buffer.

    Layout of the sink (device memory of ``dst``): ``slots`` x ``world`` 64-bit "landed" counters at offset 0, then, from
    offset 4096, ``slots`` generations of ``sum(counts)`` records in image order.  Each rank also owns a mailbox word
    (its own device memory, mapped by ``dst``) that carries ``dst``'s acknowledgements.  Step protocol (s = 0, 1, ...):

        every rank   begin(s)    : wait (local mailbox) until generation s % slots has been consumed   [s >= slots]
                                   -> returns the device address for Grouper.set_wire_output
                     ... kernels; the assemble kernel stores the records over NVLink ...
                     publish(s)  : release-store s+1 into dst's counter [s % slots][rank]
        dst          collect(s)  : wait (local counters) for every rank's s+1 -> view of generation s % slots
                     release(s)  : release-store s+1 into every rank's mailbox

    All waits are stream memory operations on LOCAL words (spg_wire_wait); all signals are one-thread kernels.
    Raises ``GroupingError`` when the GPUs cannot map each other's memory (callers fall back to PackedGather)."""

    HEADER = 4096

    def __init__(self, n_local: int, record_bytes: int, device: int, dst: int = 0, group=None, slots: int = 3):
        import torch.distributed as dist

        from . import grouping as G
        self.G = G
        self.rank, self.world = _world(group)
        self.device, self.dst, self.slots, self.record_bytes = int(device), dst, int(slots), int(record_bytes)
        self.counts = exchange_counts(n_local, group)
        self.total = sum(self.counts)
        self.offset = sum(self.counts[:self.rank])
        self.gen_bytes = self.total * self.record_bytes
        self.sink_bytes = self.HEADER + self.slots * self.gen_bytes
        assert self.slots * self.world * 8 <= self.HEADER
        self._sink_local = self._sink = 0
        self._mail_local, mail_handle = G.wire_create(self.device, 256)
        sink_handle = None
        if self.rank == dst:
            self._sink_local, sink_handle = G.wire_create(self.device, self.sink_bytes)
        handles: List[Optional[tuple]] = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, (sink_handle, mail_handle), group=group)
        else:
            handles = [(sink_handle, mail_handle)]
        self._opened: List[int] = []
        self._group = group
        err = None
        try:
            if self.rank == dst:
                self._sink = self._sink_local
                self._mail = []
                for r, (_, mh) in enumerate(handles):
                    if r == self.rank:
                        self._mail.append(self._mail_local)
                    else:
                        p = G.wire_open(self.device, mh)
                        self._opened.append(p)
                        self._mail.append(p)
            else:
                self._sink = G.wire_open(self.device, handles[dst][0])
                self._opened.append(self._sink)
        except G.GroupingError as e:
            err = str(e)
        errs: List[Optional[str]] = [err]
        if self.world > 1:  # all ranks agree on the outcome, so a failed mapping makes every rank fall back together
            errs = [None] * self.world
            dist.all_gather_object(errs, err, group=group)
        if any(errs):
            self.close()
            raise G.GroupingError("peer mapping of the wire sink failed: " + "; ".join(f"rank {r}: {e}" for r, e in enumerate(errs) if e))

    # -- producer side (every rank) -----------------------------------------------------------------
    def begin(self, step: int, stream) -> int:
        """Make ``stream`` wait until the generation ``step`` will overwrite has been consumed (call it right before
        the assemble stage is launched -- earlier kernels of the step do not need the slot); returns this rank's slice address."""
        if step >= self.slots:
            self.G.wire_wait(self.device, self._mail_local, step - self.slots + 1, stream)
        return self._sink + self.HEADER + (step % self.slots) * self.gen_bytes + self.offset * self.record_bytes

    def counter_address(self, step: int) -> int:
        """Address (in ``dst``'s memory) of this rank's "landed" counter for ``step``; the value to store is ``step + 1``."""
        return self._sink + ((step % self.slots) * self.world + self.rank) * 8

    def publish(self, step: int, stream) -> None:
        """Signal with a separate one-thread kernel.  ``Grouper.arm_wire_signal(sink.counter_address(s), s + 1)`` before
        the assemble launch does the same from inside the kernel and needs no launch."""
        self.G.wire_signal(self.device, self.counter_address(step), step + 1, stream)

    # -- consumer side (dst) ----------------------------------------------------------------------------
    def collect(self, step: int, stream):
        """``dst``: make ``stream`` wait for every rank's records of ``step``; returns the uint8 view of that generation."""
        assert self.rank == self.dst
        for r in range(self.world):
            self.G.wire_wait(self.device, self._sink_local + ((step % self.slots) * self.world + r) * 8, step + 1, stream)
        return self.G.device_bytes_view(self._sink_local + self.HEADER + (step % self.slots) * self.gen_bytes, self.gen_bytes,
                                        self.device).view(self.total, self.record_bytes)

    def release(self, step: int, stream) -> None:
        """``dst``: the generation of ``step`` has been consumed; let the producers reuse it."""
        assert self.rank == self.dst
        for lo in range(0, self.world, 32):
            self.G.wire_signal_many(self.device, self._mail[lo:lo + 32], step + 1, stream)

    def close(self) -> None:
        """Collective: unmap the peers' buffers, then (after a barrier: an exporter must outlive its importers) free ours."""
        import torch.distributed as dist
        for p in getattr(self, "_opened", []):
            self.G.wire_close(p)
        self._opened = []
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=getattr(self, "_group", None))
        if getattr(self, "_sink_local", 0):
            self.G.wire_destroy(self.device, self._sink_local)
            self._sink_local = 0
        if getattr(self, "_mail_local", 0):
            self.G.wire_destroy(self.device, self._mail_local)
            self._mail_local = 0
