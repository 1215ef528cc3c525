"""Gather of finished fields over NVLink peer memory instead of an NCCL collective.

The only exchange on the sharded paths is "every rank ends up with every rank's finished fields" (SURVEY 8e).  NCCL's
all_gather does that with kernels that need SM slots -- and the integrator's persistent launch fills every slot of the
machine (two 512-thread CTAs x 64 registers per SM), so a concurrent NCCL kernel pushes integrator CTAs into an extra
wave: measured 88 % weak scaling at 8 GPUs with the gather overlapped "for free".  Here each rank's output buffer lives
in torch symmetric memory (CUDA VMM allocations mapped into every peer), and a finished block is pushed into all peers'
buffers with plain device-to-device copies on a side stream: NVSwitch gives every pair full bandwidth, the copy engines
do the transfer and no SM is taken from the integrator.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class PeerFields:
    """``[world, count, nlat, nlon]`` f64 on every rank; rank r fills slab r of every rank's copy (``dst=None``) or of
    rank ``dst``'s copy only (a gather: what "NCCL over NVLink is used only to gather finished fields" asks for -- the
    other ranks then receive nothing, and rank ``dst`` takes (world-1) slabs through its NVLink ingress).

    ``push(first, fields)`` queues the copies of ``fields`` (this rank's windows ``first ..``) behind the work already
    on the current stream; ``finish()`` joins the side stream and synchronises the group (after it, every slab written
    before the matching ``finish()`` on its owner is readable).  ``counts`` may differ per rank (ragged shards)."""

    def __init__(self, counts, nlat, nlon, group=None, device=None, dst=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.counts = list(counts)
        self.dst = dst                      # None: every rank receives every slab (all-gather); r: only rank r does (gather)
        if len(self.counts) != self.world:
            raise ValueError('one count per rank')
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        shape = (self.world, max(self.counts), nlat, nlon)
        with torch.cuda.device(self.device):
            self.local = symm.empty(shape, dtype=torch.float64, device=self.device)
            self.handle = symm.rendezvous(self.local, self.group)
            self.peers = [self.local if r == self.rank else self.handle.get_buffer(r, shape, torch.float64)
                          for r in range(self.world)]
            self.side = torch.cuda.Stream(self.device)

    def push(self, first, fields):
        main = torch.cuda.current_stream(self.device)
        n = fields.shape[0]
        self.side.wait_event(main.record_event())
        with torch.cuda.stream(self.side):
            # remote ranks first: their copies ride NVLink while the local one only touches HBM
            for k in range(1, self.world + 1):
                r = (self.rank + k) % self.world
                if self.dst is None or r == self.dst:
                    self.peers[r][self.rank, first:first + n].copy_(fields, non_blocking=True)
        fields.record_stream(self.side)

    def finish(self):
        torch.cuda.current_stream(self.device).wait_stream(self.side)
        dist.barrier(self.group)

    def result(self):
        """``[sum(counts), nlat, nlon]`` in rank order (a view when all counts are equal)."""
        if len(set(self.counts)) == 1:
            return self.local.reshape((-1,) + tuple(self.local.shape[2:]))
        return torch.cat([self.local[r, :c] for r, c in enumerate(self.counts)], dim=0)


class ColumnFlagMail:
    """Mailboxes of the cross-rank column-flag exchange (``lcs_xrank``, include/lcs_b200.h): row-band sharding under the
    as-executed outer-product clamp.  One symmetric-memory buffer per rank, mapped into every peer; the integrator of
    each rank stores its band's column flags into all of them after every sub-step and waits for the others inside the
    persistent kernel.  ``ngroups`` (windows in flight) must be the same on every rank."""

    def __init__(self, ncol, ngroups=1, group=None, device=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.ncol, self.ngroups = int(ncol), int(ngroups)
        nbytes = int(_lib.load().lcs_xrank_mailbox_bytes(self.world, self.ngroups, self.ncol))
        with torch.cuda.device(self.device):
            self.local = symm.empty((nbytes,), dtype=torch.uint8, device=self.device)
            self.handle = symm.rendezvous(self.local, self.group)
            self.reset()
        self.struct = _lib.XRank(self.world, self.rank, self.ngroups, 0, C.c_void_p(int(self.handle.buffer_ptrs_dev)), nbytes)

    def reset(self):
        """Zero the mailbox (running exchange numbers included) on every rank; collective."""
        self.local.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
