"""Halo-only feature exchange for the row-partitioned aggregation (NCCL all-to-all over NVLink).

A rank computes the rows ``[row_lo, row_hi)`` of one CSR direction and owns the feature rows
``[own_lo, own_hi)``.  Its edges reference (a) rows it owns and (b) a set of remote rows -- the
*halo*.  Instead of all-gathering the whole feature matrix (857 MB per rank per aggregation at
P=8 for config 5), only the halo rows travel: once, at plan time, every rank tells each owner which
of its rows it needs; per aggregation the owner gathers those rows into a send buffer and one
uneven ``all_to_all_single`` delivers them.  The rank's slice of the CSR is re-indexed once into the
compact id space ``[own rows | halo rows]`` so the unchanged gather kernel reads a
``[n_own + n_halo, F]`` buffer.  With community structure most neighbours are local and the halo is
a small fraction of N; without locality it degenerates gracefully into a full exchange.

Plan construction is torch plumbing (unique / searchsorted, one-time); the per-step path is one
``index_select`` + one NCCL collective + the aggregation kernel.  (The reference is single-GPU.)
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .. import _lib
from ..graph.static.csr import HUB_THRESHOLD


class HaloPlan:
    def __init__(self, csr, row_bounds, own_bounds, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.row_lo, self.row_hi = int(row_bounds[rank]), int(row_bounds[rank + 1])
        self.own_lo, self.own_hi = int(own_bounds[rank]), int(own_bounds[rank + 1])
        self.n_rows = self.row_hi - self.row_lo
        self.n_own = self.own_hi - self.own_lo
        dev = csr.row_offset.device
        ro = csr.row_offset[self.row_lo:self.row_hi + 1].to(torch.int64)
        e0, e1 = (int(ro[0]), int(ro[-1])) if self.n_rows > 0 else (0, 0)
        cols = csr.column_indices[e0:e1].to(torch.int64)
        local = (cols >= self.own_lo) & (cols < self.own_hi)
        self.halo_ids = torch.unique(cols[~local])                      # sorted global ids
        self.n_halo = int(self.halo_ids.numel())
        remapped = torch.where(local, cols - self.own_lo,
                               self.n_own + torch.searchsorted(self.halo_ids, cols.contiguous()))
        self.local_cols = remapped.to(torch.int32).contiguous()
        self.local_row_offset = (ro - e0).to(torch.int32).contiguous()
        if csr.eids is not None and not csr.eids_identity:
            self.local_eids = csr.eids[e0:e1].contiguous()
        else:
            self.local_eids = (torch.arange(e0, e1, device=dev, dtype=torch.int32) + csr.eid_base).contiguous()
        # ---- the same rows with the edge set split in two: sources I own / sources in the halo.
        # Pass 1 (own sources, ~all edges of a community graph) overlaps the halo exchange, pass 2 adds the
        # halo contributions (stg_agg_scaled_sum_accum_f32); order inside a row is preserved.
        deg = (ro[1:] - ro[:-1])
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=dev, dtype=torch.int64), deg)

        def sub_csr(mask, values):
            cnt = torch.bincount(rows[mask], minlength=self.n_rows)
            sub_ro = torch.zeros(self.n_rows + 1, dtype=torch.int32, device=dev)
            sub_ro[1:] = torch.cumsum(cnt, 0).to(torch.int32)
            return sub_ro.contiguous(), values[mask].to(torch.int32).contiguous(), self.local_eids[mask].contiguous()

        self.own_ro, self.own_cols, self.own_eids = sub_csr(local, cols - self.own_lo)
        self.halo_ro, self.halo_cols, self.halo_eids = sub_csr(~local, torch.searchsorted(self.halo_ids, cols.contiguous()))
        # the halo-source edges once more, over ONLY the rows that have any (view row i = local row halo_out_rows[i])
        hdeg = (self.halo_ro[1:] - self.halo_ro[:-1])
        sel = torch.nonzero(hdeg > 0).reshape(-1)
        self.halo_out_rows = sel.to(torch.int32).contiguous()
        cro = torch.zeros(sel.numel() + 1, dtype=torch.int32, device=dev)
        cro[1:] = torch.cumsum(hdeg[sel].to(torch.int64), 0).to(torch.int32)
        self.halo_compact_ro = cro.contiguous()
        self.eid_base = csr.eid_base
        self.num_global_edges = csr.num_edges
        self.num_local_edges = e1 - e0
        # ---- who sends what: split my halo by owner, tell every owner which rows I need
        bt = torch.as_tensor(list(own_bounds), dtype=torch.int64, device=dev)
        split = torch.searchsorted(self.halo_ids, bt)
        recv_counts = (split[1:] - split[:-1]).to(torch.int64)
        send_counts = torch.empty_like(recv_counts)
        dist.all_to_all_single(send_counts, recv_counts, group=group)
        self.out_splits = [int(v) for v in recv_counts.cpu()]            # rows I receive from each owner
        self.in_splits = [int(v) for v in send_counts.cpu()]             # rows I send to each requester
        send_ids = torch.empty(sum(self.in_splits), dtype=torch.int64, device=dev)
        dist.all_to_all_single(send_ids, self.halo_ids, self.in_splits, self.out_splits, group=group)
        self.send_index = (send_ids - self.own_lo).contiguous()
        self._view = None
        self._hub = None
        self._split_views = None
        self._keep = []

    # ------------------------------------------------------- overlapped two-pass form
    @staticmethod
    def hub_threshold_for(n_edges: int, n_rows: int) -> int:
        """Rows longer than this go to the block-per-row kernel.  One warp walks a row at ~8 edges per memory round
        trip, so the longest warp-owned row bounds the launch from below: 1024 edges (~0.1 ms) are invisible in the
        2 ms pass over the whole graph but a third of a 0.3 ms pass over one rank's share of an 8-way partition, and
        more than the whole halo-source pass (ncu, r2: 23 % of the warp samples of that pass sat at the final barrier
        behind one long row).  Measured on rank 0's share of config 5: at P=8 the halo-source pass (3.5 edges per row)
        takes 0.306 / 0.146 / 0.143 / 0.153 / 0.218 ms at no split / 64 / 128 / 256 / 1024 and the own-source pass 0.409 /
        0.281 / 0.260 / 0.266 / 0.295 ms at 64 / 128 / 256 / 512 / 1024; at P=2 (29 M edges) 1.17 / 1.03 / 1.02 / 1.08 ms at
        128 / 256 / 512 / 1024."""
        env = os.environ.get("STG_PART_HUB_THRESHOLD")       # A/B knob
        if env:
            return int(env)
        if n_rows > 0 and n_edges < 8 * n_rows:
            return 128
        return 256 if n_edges < 16_000_000 else 512

    def _make_view(self, ro, cols, eids, n_rows=None):
        dev = ro.device
        n_rows = self.n_rows if n_rows is None else n_rows
        n_e = int(cols.shape[0])
        threshold = self.hub_threshold_for(n_e, n_rows)
        cap = n_e // max(threshold, 1) + 1
        hub_rows = torch.empty(cap, dtype=torch.int32, device=dev)
        hub_count = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("stg_csr_hub_rows", ro.data_ptr(), n_rows, threshold, hub_rows.data_ptr(), cap,
                  hub_count.data_ptr(), _lib.current_stream_ptr())
        has_hubs = int(hub_count.item()) > 0
        self._keep.append((hub_rows, hub_count))
        v = _lib.StgCsrView()
        v.row_offset = ro.data_ptr()
        v.column_indices = cols.data_ptr() if n_e else None
        v.eids = eids.data_ptr() if n_e else None
        v.node_ids = None
        v.num_nodes = n_rows
        v.num_edges = n_e                  # edges of THIS sub-CSR (the kernel picks its row geometry from the mean degree)
        v.eid_base = self.eid_base
        v.eids_identity = 0
        v.hub_rows = hub_rows.data_ptr() if has_hubs else None
        v.hub_count = hub_count.data_ptr() if has_hubs else None
        v.hub_threshold = threshold if has_hubs else 0
        v.hub_capacity = cap if has_hubs else 0
        queue = torch.zeros(2, dtype=torch.int32, device=dev)      # global row queue of the aggregation kernel
        self._keep.append(queue)
        v.work_queue = queue.data_ptr()
        return v

    def split_views(self):
        """(view over edges whose source I own, view over edges whose source is in the halo)."""
        if self._split_views is None:
            self._split_views = (self._make_view(self.own_ro, self.own_cols, self.own_eids),
                                 self._make_view(self.halo_ro, self.halo_cols, self.halo_eids))
        return self._split_views

    def halo_compact_view(self):
        """Halo-source edges over the rows that have any; pair with ``out_rows=self.halo_out_rows``."""
        if getattr(self, "_compact_view", None) is None:
            self._compact_view = self._make_view(self.halo_compact_ro, self.halo_cols, self.halo_eids,
                                                 n_rows=int(self.halo_out_rows.numel()))
        return self._compact_view

    def new_halo_buffer(self, feat: int, like: torch.Tensor) -> torch.Tensor:
        return torch.empty(max(self.n_halo, 1), feat, dtype=like.dtype, device=like.device)[:self.n_halo]

    def start_exchange(self, own: torch.Tensor, halo_buf: torch.Tensor):
        """Gather the rows my peers need and start the all-to-all on NCCL's stream (returns the Work)."""
        send = own.index_select(0, self.send_index) if self.send_index.numel() else own[:0]
        self._send = send                                  # keep alive until the collective has run
        return dist.all_to_all_single(halo_buf, send.contiguous(), self.out_splits, self.in_splits, group=self.group,
                                      async_op=True)

    def halo_vector(self, own_values: torch.Tensor) -> torch.Tensor:
        """Per-node scalars of the halo rows (one-time exchange), e.g. their ``norm``."""
        buf = torch.empty(max(self.n_halo, 1), 1, dtype=own_values.dtype, device=own_values.device)[:self.n_halo]
        self.start_exchange(own_values.reshape(-1, 1).contiguous(), buf).wait()
        return buf.reshape(-1).contiguous()

    def aggregate(self, kernels, own, halo_buf, ns_own, ns_halo, rs, out, es=None):
        """``out = rs * sum(...)`` over this rank's rows: own-source pass overlapped with the halo exchange."""
        v_own, v_halo = self.split_views()
        work = self.start_exchange(own, halo_buf)
        kernels.agg_scaled_sum(v_own, own, ns_own, es, rs, out=out)
        work.wait()
        if self.n_halo > 0:
            kernels.agg_scaled_sum(v_halo, halo_buf, ns_halo, es, rs, out=out, accumulate=True)
        return out

    def aggregate_pull(self, kernels, peer, halo_buf, ns_own, ns_halo, rs, out, es=None, pull_blocks=32):
        """Same two-pass aggregation with the halo fetched by OUR kernel over peer memory (no NCCL).

        ``peer``: :class:`PeerBlocks` whose ``own`` view holds this rank's rows.  A side stream runs
        ``stg_halo_pull_f32`` (few CTAs, NVLink loads from the owners' blocks) while the main stream
        aggregates the edges whose sources are local; the second pass adds the halo contributions.
        """
        import ctypes

        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(priority=-1)
            p = len(peer.ptrs)
            self._pull_ptrs = (ctypes.c_void_p * p)(*[ctypes.c_void_p(a) for a in peer.ptrs])
            self._pull_bounds = (ctypes.c_int32 * (p + 1))(*peer.bounds)
        v_own, v_halo = self.split_views()
        cur = torch.cuda.current_stream()
        prof = getattr(self, "profile", None)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)] if prof is not None else None
        if ev:
            ev[0].record(cur)
        peer.barrier()                                   # every rank's block is written
        if ev:
            ev[1].record(cur)
        if self.n_halo > 0:
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                if ev:
                    ev[6].record(self._side)
                _lib.call("stg_halo_pull_f32", self._pull_ptrs, self._pull_bounds, len(peer.ptrs), self.halo_ids.data_ptr(),
                          self.n_halo, int(halo_buf.shape[1]), halo_buf.data_ptr(), int(pull_blocks),
                          self._side.cuda_stream)
                if ev:
                    ev[7].record(self._side)
        kernels.agg_scaled_sum(v_own, peer.own, ns_own, es, rs, out=out)
        if ev:
            ev[2].record(cur)
        if self.n_halo > 0:
            cur.wait_stream(self._side)
            if ev:
                ev[3].record(cur)
            kernels.agg_scaled_sum(v_halo, halo_buf, ns_halo, es, rs, out=out, accumulate=True)
        if ev:
            ev[4].record(cur)
        peer.barrier()                                   # nobody still reads my block when I overwrite it next
        if ev:
            ev[5].record(cur)
            prof.append(ev)
        return out

    # ------------------------------------------------- push form (posted NVLink stores, fully overlapped)
    def setup_push(self, feat: int):
        """Symmetric double-buffered halo storage + the (row, peer, slot) send list for ``stg_halo_push_f32``."""
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        dev = self.local_row_offset.device
        group = self.group if self.group is not None else dist.group.WORLD
        mx = torch.tensor([self.n_halo], dtype=torch.int64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        self._max_halo = max(int(mx.item()), 1)
        self._halo_symm = symm_mem.empty((2 * self._max_halo, feat), dtype=torch.float32, device=dev)
        self._halo_hdl = symm_mem.rendezvous(self._halo_symm, group)
        base = [int(a) for a in self._halo_hdl.buffer_ptrs]
        stride = self._max_halo * feat * 4
        self._push_ptrs = [(ctypes.c_void_p * self.world)(*[ctypes.c_void_p(b + k * stride) for b in base]) for k in (0, 1)]
        # where my rows land in each requester's halo buffer: requesters order their halo by owner, then id
        recv_off = torch.zeros(self.world, dtype=torch.int64, device=dev)
        recv_off[1:] = torch.cumsum(torch.as_tensor(self.out_splits[:-1], dtype=torch.int64, device=dev), 0)
        dst_off = torch.empty_like(recv_off)
        dist.all_to_all_single(dst_off, recv_off, group=self.group)
        counts = torch.as_tensor(self.in_splits, dtype=torch.int64, device=dev)
        peer = torch.repeat_interleave(torch.arange(self.world, device=dev, dtype=torch.int64), counts)
        seg_start = torch.zeros(self.world, dtype=torch.int64, device=dev)
        seg_start[1:] = torch.cumsum(counts[:-1], 0)
        within = torch.arange(int(counts.sum()), device=dev, dtype=torch.int64) - seg_start[peer]
        self._send_peer = peer.to(torch.int32).contiguous()
        self._send_slot = (dst_off[peer] + within).contiguous()
        self._side = torch.cuda.Stream(priority=-1)
        self._iter = 0
        self._feat = feat
        self.split_views()

    def halo_rows_view(self, k: int) -> torch.Tensor:
        return self._halo_symm[k * self._max_halo: k * self._max_halo + self.n_halo]

    def aggregate_push(self, kernels, own, ns_own, ns_halo, rs, out, es=None, push_blocks=32, flow="serial"):
        """Halo rows are pushed into the peers' symmetric halo buffers by ``stg_halo_push_f32`` (posted NVLink
        stores, side stream) while the own-source pass runs; one device barrier; then the halo-source pass.
        Double-buffered halo storage makes that single barrier sufficient.

        ``flow="serial"`` (default): the own-source pass WRITES every row, the halo-source pass then ``+=`` over
        only the rows that have a remote neighbour (compacted view, ``stg_agg_scaled_sum_rows_f32``) on the same
        stream: no memset, no atomics, fixed order.  ``flow="concurrent"``: both passes ``red.add`` into a
        zero-filled ``out`` on two streams (at most two addends per element -> still deterministic); measured
        slower once the gather kernel became L2-bound, because concurrency no longer hides any work."""
        if flow == "serial":
            return self._aggregate_push_serial(kernels, own, ns_own, ns_halo, rs, out, es, push_blocks)
        k = self._iter & 1
        self._iter += 1
        v_own, v_halo = self.split_views()
        cur = torch.cuda.current_stream()
        side = self._side
        prof = getattr(self, "push_profile", None)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)] if prof is not None else None
        if ev:
            ev[0].record(cur)
        out.zero_()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if ev:
                ev[2].record(side)
            _lib.call("stg_halo_push_f32", own.data_ptr(), self._feat, self.send_index.data_ptr(), self._send_peer.data_ptr(),
                      self._send_slot.data_ptr(), int(self.send_index.numel()), self._push_ptrs[k], self.world,
                      int(push_blocks), side.cuda_stream)
            if ev:
                ev[3].record(side)
            self._halo_hdl.barrier()                      # every rank's pushes for this step have landed
            if ev:
                ev[4].record(side)
            if self.n_halo > 0:
                kernels.agg_scaled_sum(v_halo, self.halo_rows_view(k), ns_halo, es, rs, out=out, accumulate="red",
                                       stream=side.cuda_stream)
            if ev:
                ev[5].record(side)
        if ev:
            ev[1].record(cur)
        kernels.agg_scaled_sum(v_own, own, ns_own, es, rs, out=out, accumulate="red")
        if ev:
            ev[6].record(cur)
            prof.append(ev)
        cur.wait_stream(side)
        return out

    def _aggregate_push_serial(self, kernels, own, ns_own, ns_halo, rs, out, es, push_blocks):
        k = self._iter & 1
        self._iter += 1
        v_own, _ = self.split_views()
        cur = torch.cuda.current_stream()
        side = self._side
        prof = getattr(self, "push_profile", None)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)] if prof is not None else None
        if ev:
            ev[0].record(cur)
            ev[1].record(cur)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if ev:
                ev[2].record(side)
            _lib.call("stg_halo_push_f32", own.data_ptr(), self._feat, self.send_index.data_ptr(), self._send_peer.data_ptr(),
                      self._send_slot.data_ptr(), int(self.send_index.numel()), self._push_ptrs[k], self.world,
                      int(push_blocks), side.cuda_stream)
            if ev:
                ev[3].record(side)
            self._halo_hdl.barrier()                      # every rank's pushes for this step have landed
            if ev:
                ev[4].record(side)
        kernels.agg_scaled_sum(v_own, own, ns_own, es, rs, out=out)
        if ev:
            ev[6].record(cur)
        cur.wait_stream(side)
        if self.n_halo > 0:
            kernels.agg_scaled_sum(self.halo_compact_view(), self.halo_rows_view(k), ns_halo, es, rs, out=out,
                                   accumulate=True, out_rows=self.halo_out_rows)
        if ev:
            ev[5].record(cur)
            prof.append(ev)
        return out

    def push_profile_summary(self):
        """Mean device time (ms) of the segments of aggregate_push (set ``plan.push_profile = []`` to collect)."""
        torch.cuda.synchronize()
        n = max(len(self.push_profile), 1)
        seg = {"zero": (0, 1), "own_pass": (1, 6), "push": (2, 3), "barrier": (3, 4), "halo_pass_or_wait": (4, 5),
               "own_end_to_halo_end": (6, 5), "start_to_halo_end": (0, 5)}
        return {k: sum(e[a].elapsed_time(e[b]) for e in self.push_profile) / n for k, (a, b) in seg.items()}

    def profile_summary(self):
        """Mean device time (ms) of each segment of aggregate_pull (set ``plan.profile = []`` to collect)."""
        torch.cuda.synchronize()
        names = ["barrier_in", "own_pass", "wait_pull", "halo_pass", "barrier_out"]
        acc = {k: 0.0 for k in names + ["pull_kernel", "total"]}
        for ev in self.profile:
            for i, k in enumerate(names):
                acc[k] += ev[i].elapsed_time(ev[i + 1])
            acc["pull_kernel"] += ev[6].elapsed_time(ev[7])
            acc["total"] += ev[0].elapsed_time(ev[5])
        return {k: v / max(len(self.profile), 1) for k, v in acc.items()}

    # ------------------------------------------------------------------ buffers
    def new_buffer(self, feat: int, like: torch.Tensor) -> torch.Tensor:
        """``[n_own + n_halo, feat]``: own rows first (write them in place), halo rows behind."""
        return torch.empty(self.n_own + self.n_halo, feat, dtype=like.dtype, device=like.device)

    def exchange(self, buf: torch.Tensor) -> torch.Tensor:
        """Fill the halo part of ``buf`` from the owners (own rows ``buf[:n_own]`` must be valid)."""
        own = buf[:self.n_own]
        send = own.index_select(0, self.send_index) if self.send_index.numel() else own[:0]
        dist.all_to_all_single(buf[self.n_own:], send.contiguous(), self.out_splits, self.in_splits, group=self.group)
        return buf

    def extend_vector(self, own_values: torch.Tensor) -> torch.Tensor:
        """Per-node scalars (e.g. ``norm``) in the ``[own | halo]`` id space (one-time exchange)."""
        buf = torch.empty(self.n_own + self.n_halo, 1, dtype=own_values.dtype, device=own_values.device)
        buf[:self.n_own, 0] = own_values.reshape(-1)
        return self.exchange(buf).reshape(-1)

    # --------------------------------------------------------------------- view
    def view(self) -> _lib.StgCsrView:
        if self._view is None:
            dev = self.local_row_offset.device
            cap = self.num_local_edges // max(HUB_THRESHOLD, 1) + 1
            hub_rows = torch.empty(cap, dtype=torch.int32, device=dev)
            hub_count = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.call("stg_csr_hub_rows", self.local_row_offset.data_ptr(), self.n_rows, HUB_THRESHOLD,
                      hub_rows.data_ptr(), cap, hub_count.data_ptr(), _lib.current_stream_ptr())
            has_hubs = int(hub_count.item()) > 0
            self._hub = (hub_rows, hub_count)
            v = _lib.StgCsrView()
            v.row_offset = self.local_row_offset.data_ptr()
            v.column_indices = self.local_cols.data_ptr()
            v.eids = self.local_eids.data_ptr()
            v.node_ids = None
            v.num_nodes = self.n_rows
            v.num_edges = self.num_global_edges
            v.eid_base = self.eid_base
            v.eids_identity = 0
            v.hub_rows = hub_rows.data_ptr() if has_hubs else None
            v.hub_count = hub_count.data_ptr() if has_hubs else None
            v.hub_threshold = HUB_THRESHOLD if has_hubs else 0
            v.hub_capacity = cap if has_hubs else 0
            self._view = v
        return self._view
