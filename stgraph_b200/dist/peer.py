"""Peer-memory feature blocks: the halo exchange fused into the gather kernel.

Every rank keeps ITS row block of the feature matrix in a symmetric-memory buffer
(``torch.distributed._symmetric_memory``: same-size allocation on every GPU, mapped into every
process over NVLink/NVSwitch).  The aggregation kernel receives the P base pointers and fetches a
remote neighbour row with plain NVLink loads while it aggregates
(``stg_agg_scaled_sum_parts_f32``): no all-gather is materialised, no send/recv buffers, and the
transfer overlaps the local gathers warp by warp.  Only a device-side barrier separates "every rank
has written its block" from "kernels may read peers".  torch is plumbing here (allocation,
rendezvous, barrier); the data path is the hand-written kernel.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class PeerBlocks:
    def __init__(self, bounds, feat: int, rank: int, world: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.bounds = [int(b) for b in bounds]
        self.rank, self.world, self.feat = rank, world, feat
        max_rows = max(self.bounds[q + 1] - self.bounds[q] for q in range(world))
        group = group if group is not None else dist.group.WORLD
        self._buf = symm_mem.empty((max(max_rows, 1), feat), dtype=torch.float32, device=device)
        self._hdl = symm_mem.rendezvous(self._buf, group)
        self.ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        self.n_own = self.bounds[rank + 1] - self.bounds[rank]
        #: this rank's rows (write the layer input / gradient here, in place)
        self.own = self._buf[:self.n_own]

    def barrier(self):
        """All ranks have finished writing (or reading) their blocks; stream-ordered, no host sync."""
        self._hdl.barrier()
