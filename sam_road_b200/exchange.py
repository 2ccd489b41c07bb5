"""The path's one exchange step -- every rank's per-tile results to every rank -- without
taking SMs away from the kernels it overlaps.

The compute kernels of this library are persistent (one CTA per SM, static work split): a collective
KERNEL that occupies a few SMs while they run delays the CTAs that would have run there for the whole
collective, and with a static split those CTAs then finish last (measured: 5 % per step with NCCL's
all-gather kernels on 8 CTAs beside the encoder).  `TileExchange` therefore moves the data with the COPY
ENGINES over NVLink: the gather buffer of every rank is symmetric memory (torch symmetric memory: cuMem
allocations mapped into every peer's address space), a rank pushes its block into each peer's buffer with
plain device-to-device copies on a side stream, and a signal-pad barrier on that stream tells every rank
when all blocks have landed.  No reduction is involved, so results are bit-identical to an all-gather.

Fallback (no symmetric memory / no P2P): `all_gather_into_tensor` on NCCL, asynchronously, with the
communicator limited to a few CTAs.  `TileExchange.backend` says which one is in use.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist


class TileExchange:
    def __init__(self, rows_per_rank: int, row_shape: Sequence[int], dtype: torch.dtype, device: torch.device,
                 group=None, slots: int = 1, prefer_copy_engine: bool = True):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows = int(rows_per_rank)
        self.device = torch.device(device)
        self.slots = slots
        full = (self.world * self.rows, *row_shape)
        self.backend = "nccl_async"
        self.note = ""
        self._bufs, self._hdl, self._peer = [], [], []
        self._flags = self._flag_peers = None      # symmetric int32 [slots, world] + peer views (memop barrier)
        self._round = [0] * slots
        self._works = [None] * slots
        self._done = [None] * slots
        self.stream = torch.cuda.Stream(self.device)
        if prefer_copy_engine and self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                for _ in range(slots):
                    buf = symm_mem.empty(full, dtype=dtype, device=self.device)
                    hdl = symm_mem.rendezvous(buf, grp)
                    self._bufs.append(buf)
                    self._hdl.append(hdl)
                    self._peer.append([buf if r == self.rank else hdl.get_buffer(r, full, dtype)
                                       for r in range(self.world)])
                flags = symm_mem.empty((slots, self.world), dtype=torch.int32, device=self.device)
                flags.zero_()
                fh = symm_mem.rendezvous(flags, grp)
                self._flags, self._flag_hdl = flags, fh
                self._flag_peers = [flags if r == self.rank else fh.get_buffer(r, (slots, self.world), torch.int32)
                                    for r in range(self.world)]
                torch.cuda.synchronize(self.device)
                dist.barrier(group=group)              # every rank's flags are zero before anybody writes
                self.backend = "peer_copy_engine"
                try:                                   # probe the stream memory operations once
                    self._barrier_memops(0, self.stream)
                    self.stream.synchronize()
                    self.barrier_kind = "stream_memops"
                except Exception as e:
                    self.barrier_kind = f"signal_pad_kernel ({type(e).__name__})"
                dist.barrier(group=group)
            except Exception as e:      # symmetric memory unavailable on this system: NCCL path
                self.note = f"symmetric memory unavailable ({type(e).__name__}: {str(e)[:120]})"
                self._bufs, self._hdl, self._peer = [], [], []
        if not self._bufs:
            self._bufs = [torch.empty(full, dtype=dtype, device=self.device) for _ in range(slots)]

    def _barrier_memops(self, slot: int, stream) -> None:
        """Arrive (write the round number into this rank's word of every peer's flag row) and wait (until
        every word of this rank's own row has reached it) -- stream memory operations, no kernel."""
        from . import _lib
        lib = _lib.load()
        self._round[slot] += 1
        r = self._round[slot]
        sp = stream.cuda_stream
        for k in range(1, self.world):
            peer = (self.rank + k) % self.world
            addr = self._flag_peers[peer].data_ptr() + 4 * (slot * self.world + self.rank)
            _lib.check(lib.samroad_stream_write_value32(addr, r, sp), "samroad_stream_write_value32")
        for k in range(1, self.world):
            peer = (self.rank + k) % self.world
            addr = self._flags.data_ptr() + 4 * (slot * self.world + peer)
            _lib.check(lib.samroad_stream_wait_value32(addr, r, sp), "samroad_stream_wait_value32")

    def _barrier(self, slot: int) -> None:
        if getattr(self, "barrier_kind", "") == "stream_memops":
            self._barrier_memops(slot, self.stream)
        else:
            self._hdl[slot].barrier(channel=slot)

    # the rank's own block inside its gather buffer: producers write their results straight into it
    def local_block(self, slot: int = 0) -> torch.Tensor:
        return self._bufs[slot][self.rank * self.rows: (self.rank + 1) * self.rows]

    def gathered(self, slot: int = 0) -> torch.Tensor:
        return self._bufs[slot]

    def publish(self, slot: int = 0, lo: int = 0, hi: Optional[int] = None, first: bool = True,
                last: bool = True) -> None:
        """Send rows [lo, hi) of the local block (already written on the current stream) to every rank.
        `first=True` opens a round on this slot: no rank starts overwriting a peer's buffer before every
        rank has finished reading the previous round's contents (everything enqueued on its current
        stream before its publish call).  `last=True` closes the round: after `wait(slot)` the gather
        buffer is complete."""
        hi = self.rows if hi is None else hi
        ready = torch.cuda.Event()
        ready.record()
        if self.backend == "peer_copy_engine":
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                if first:
                    self._barrier(slot)                      # every rank is done with the slot's old contents
                if hi > lo:
                    src = self.local_block(slot)[lo:hi]
                    a, b = self.rank * self.rows + lo, self.rank * self.rows + hi
                    for r in range(1, self.world):          # staggered peers: every link busy at once
                        peer = (self.rank + r) % self.world
                        self._peer[slot][peer][a:b].copy_(src, non_blocking=True)
                if last:
                    self._barrier(slot)                      # all ranks' pushes into this slot have landed
                    ev = torch.cuda.Event()
                    ev.record()
                    self._done[slot] = ev
        else:
            if not last:
                return                                        # the collective moves the whole block at once
            self._works[slot] = dist.all_gather_into_tensor(self._bufs[slot], self.local_block(slot).clone(),
                                                            group=self.group, async_op=True)

    def wait(self, slot: int = 0) -> None:
        """Make the current stream wait for the exchange of `slot`."""
        if self.backend == "peer_copy_engine":
            if self._done[slot] is not None:
                torch.cuda.current_stream(self.device).wait_event(self._done[slot])
                self._done[slot] = None
        elif self._works[slot] is not None:
            self._works[slot].wait()
            self._works[slot] = None

    def drain(self) -> None:
        for s in range(self.slots):
            self.wait(s)


def limit_nccl_ctas(world: int) -> None:
    """NCCL kernels that run beside persistent compute kernels should hold as few SMs as the payload needs."""
    import os
    os.environ.setdefault("NCCL_MAX_CTAS", "2" if world <= 2 else "4")
