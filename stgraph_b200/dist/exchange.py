"""Per-aggregation halo exchange + the two aggregation passes of one rank (new: the reference is single-GPU).

One :class:`HaloExchange` serves one CSR direction of a :class:`~stgraph_b200.dist.partition.PartitionedGraph`
at one feature width.  Per aggregation (``aggregate``):

    side stream (high priority)                          my stream
    -----------------------------------------------      -------------------------------------------------
    gather the rows my peers need into a send buffer      own-source pass: every local row, edges whose
    P-1 copy-engine copies into the peers' halo buffers     source row I own (packed {col, scale} metadata,
    P-1 copy-engine writes of the arrival flags             global row queue)  ->  out = ...
                                                          wait for the P-1 arrival flags (one spinning warp)
                                                          halo-source pass: rows with a remote neighbour,
                                                            out += ... (row-subset form)

Only the gather runs on SMs (``stg_halo_exchange_f32``): no SM copies halo bytes over NVLink and nothing on the side
stream needs an SM slot once the persistent aggregation grid has filled the chip (a one-warp signalling kernel waited
0.47 ms for one); no NCCL and no device-wide barrier on the data path: the halo buffers are double-buffered in
symmetric memory and the flags only grow, which orders "peer q has read buffer k of step t-2" before "I overwrite it
at step t" transitively (my step t starts after my halo pass t-1, which waited for q's flag t-1, which q posted after
its step t-1 sends, which q started after its halo pass t-2).  Measured alternatives (8 GPUs, config 5): the round-1 SM
push kernel (``mode="sm"``, kept for A/B runs) 0.5-0.78 ms per exchange and the own-source pass slowed from 0.32 to
0.6 ms; per-peer gather kernels feeding copies on three private streams, gathers ahead of the aggregation pass: no
faster at 8 ranks (1.33-1.35 vs 1.31 ms/step), slower at 4 (2.02 vs 1.74 ms/step) -- removed; the halo-source pass
split by arrival order of the owners (first half of the peers while the second half is still on the wire): 1.41 ms/step
with two groups, 1.63 with three, against 1.26 with one pass -- every extra pass revisits the rows -- removed.  torch is plumbing here:
allocation, symmetric-memory rendezvous, streams.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from .. import _lib

#: SM clocks a rank waits for its peers' flags before it gives up and flags the step as failed (~2 s at 1.9 GHz)
WAIT_TIMEOUT_CYCLES = int(os.environ.get("STG_PEER_WAIT_CYCLES", str(4 << 30)))


class HaloExchange:
    def __init__(self, plan, feat: int, ns_own: torch.Tensor | None, ns_halo: torch.Tensor | None,
                 mode: str | None = None, gather_blocks: int = 0, push_blocks: int = 64,
                 edge_scale: torch.Tensor | None = None):
        import torch.distributed._symmetric_memory as symm_mem

        from .. import kernels

        self.plan, self.feat = plan, int(feat)
        self.mode = mode or os.environ.get("STG_HALO_MODE", "ce")
        self.gather_blocks, self.push_blocks = int(gather_blocks), int(push_blocks)
        world, rank = plan.world, plan.rank
        dev = plan.local_row_offset.device
        group = plan.group if plan.group is not None else dist.group.WORLD
        mx = torch.tensor([plan.n_halo], dtype=torch.int64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=plan.group)
        self.max_halo = max(int(mx.item()), 1)
        # symmetric memory: two halo buffers + two flag rows per rank, mapped into every process
        self._halo = symm_mem.empty((2 * self.max_halo, self.feat), dtype=torch.float32, device=dev)
        self._halo_hdl = symm_mem.rendezvous(self._halo, group)
        self._flags = symm_mem.empty((2 * 16,), dtype=torch.int32, device=dev)
        self._flags.zero_()
        self._flags_hdl = symm_mem.rendezvous(self._flags, group)
        torch.cuda.synchronize(dev)
        dist.barrier(group=plan.group)             # every rank's flags are zero before anybody signals
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._seq = torch.arange(65536, dtype=torch.int32, device=dev)     # source words of the copy-engine flag writes
        n_send = int(plan.send_index.numel())
        self.send_buf = torch.empty(max(n_send, 1), self.feat, dtype=torch.float32, device=dev)
        # where my rows land in each requester's halo buffer (requesters order their halo by owner, then id)
        recv_off = torch.zeros(world, dtype=torch.int64, device=dev)
        recv_off[1:] = torch.cumsum(torch.as_tensor(plan.out_splits[:-1], dtype=torch.int64, device=dev), 0)
        dst_off = torch.empty_like(recv_off)
        dist.all_to_all_single(dst_off, recv_off, group=plan.group)
        dst_off = [int(v) for v in dst_off.cpu()]
        halo_base = [int(a) for a in self._halo_hdl.buffer_ptrs]
        flag_base = [int(a) for a in self._flags_hdl.buffer_ptrs]
        stride = self.max_halo * self.feat * 4
        self._peer_dst, self._peer_flag, self._peer_halo = [], [], []
        for k in (0, 1):
            self._peer_dst.append((ctypes.c_void_p * world)(
                *[ctypes.c_void_p(halo_base[q] + k * stride + dst_off[q] * self.feat * 4) for q in range(world)]))
            self._peer_halo.append((ctypes.c_void_p * world)(*[ctypes.c_void_p(halo_base[q] + k * stride) for q in range(world)]))
            self._peer_flag.append((ctypes.c_void_p * world)(
                *[ctypes.c_void_p(flag_base[q] + (k * 16 + rank) * 4) for q in range(world)]))
        off = [0]
        for c in plan.in_splits:
            off.append(off[-1] + int(c))
        self._send_off = (ctypes.c_int64 * (world + 1))(*off)
        self._zero_off = (ctypes.c_int64 * (world + 1))(*([0] * (world + 1)))
        if self.mode == "sm":       # (row, peer, slot) list of the SM push kernel
            counts = torch.as_tensor(plan.in_splits, dtype=torch.int64, device=dev)
            peer = torch.repeat_interleave(torch.arange(world, device=dev, dtype=torch.int64), counts)
            seg = torch.as_tensor(off[:-1], dtype=torch.int64, device=dev)
            within = torch.arange(n_send, device=dev, dtype=torch.int64) - seg[peer]
            self._send_peer = peer.to(torch.int32).contiguous()
            self._send_slot = (torch.as_tensor(dst_off, dtype=torch.int64, device=dev)[peer] + within).contiguous()
        self.side = torch.cuda.Stream(device=dev, priority=-1)
        self._ev_in = torch.cuda.Event()
        self._ev_sent = torch.cuda.Event()
        self._iter = 0
        # the two sub-CSRs of my rows with their packed {col, scale} metadata (scales are fixed per graph: norm)
        self.v_own, _ = plan.split_views()
        self.v_halo = plan.halo_compact_view()
        # edge_scale (edge weights): one value per GLOBAL edge id, replicated on every rank like the structure; the
        # sub-CSRs carry the global ids of their edges, so the product ns * es is folded into the packed metadata
        self.meta_own = kernels.pack_edge_meta(self.v_own, ns_own, edge_scale, device=dev) if plan.own_cols.numel() else None
        self.meta_halo = (kernels.pack_edge_meta(self.v_halo, ns_halo, edge_scale, device=dev)
                          if plan.halo_cols.numel() else None)
        self.ns_own, self.ns_halo = ns_own, ns_halo
        self.profile = None            # set to [] to collect per-call CUDA events (bench.py `segments`)

    def halo_rows(self, k: int) -> torch.Tensor:
        return self._halo[k * self.max_halo: k * self.max_halo + self.plan.n_halo]

    # ------------------------------------------------------------------ one aggregation
    def aggregate(self, x_own: torch.Tensor, rs: torch.Tensor | None, out: torch.Tensor) -> torch.Tensor:
        """``out[r] = rs[r] * sum_e scale_e * x[col_e]`` over my rows; ``x_own`` = my rows of the source matrix."""
        from .. import kernels

        plan, world, rank = self.plan, self.plan.world, self.plan.rank
        assert x_own.shape == (plan.n_own, self.feat) and x_own.is_contiguous()
        self._iter += 1
        it, k = self._iter, self._iter & 1
        cur = torch.cuda.current_stream()
        side = self.side
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)] if self.profile is not None else None
        if ev:
            ev[0].record(cur)
        self._ev_in.record(cur)
        side.wait_event(self._ev_in)
        n_send = int(plan.send_index.numel())
        with torch.cuda.stream(side):
            if ev:
                ev[4].record(side)
            if self.mode == "sm":
                if n_send:
                    _lib.call("stg_halo_push_f32", x_own.data_ptr(), self.feat, plan.send_index.data_ptr(),
                              self._send_peer.data_ptr(), self._send_slot.data_ptr(), n_send, self._peer_halo[k], world,
                              self.push_blocks, side.cuda_stream)
                if ev:
                    ev[5].record(side)
                    ev[6].record(side)
                _lib.call("stg_peer_signal", self._peer_flag[k], world, rank, it & 0xFFFF, side.cuda_stream)
            elif ev:          # profiling: the three steps one by one, with an event between them
                _lib.call("stg_rows_gather_f32", x_own.data_ptr(), self.feat, plan.send_index.data_ptr(), n_send,
                          self.send_buf.data_ptr(), self.gather_blocks, side.cuda_stream)
                ev[5].record(side)
                _lib.call("stg_halo_send_f32", self.send_buf.data_ptr(), self.feat, world, rank, self._send_off,
                          self._peer_dst[k], side.cuda_stream)
                ev[6].record(side)
                _lib.call("stg_halo_exchange_f32", x_own.data_ptr(), self.feat, plan.send_index.data_ptr(), self._zero_off,
                          self.send_buf.data_ptr(), self._peer_dst[k], self._peer_flag[k], self._seq.data_ptr(), it & 0xFFFF,
                          world, rank, self.gather_blocks, side.cuda_stream)        # no rows: the flag copies only
            else:
                _lib.call("stg_halo_exchange_f32", x_own.data_ptr(), self.feat, plan.send_index.data_ptr(), self._send_off,
                          self.send_buf.data_ptr(), self._peer_dst[k], self._peer_flag[k], self._seq.data_ptr(), it & 0xFFFF,
                          world, rank, self.gather_blocks, side.cuda_stream)
            self._ev_sent.record(side)
            if ev:
                ev[7].record(side)
        # my stream: the edges whose source I own (launched right behind the gather: the two run side by side)
        if self.meta_own is not None:
            kernels.agg_packed_sum_rows(self.v_own, self.meta_own, None, x_own, rs, out, accumulate=False)
        else:
            out.zero_()
        if ev:
            ev[1].record(cur)
        _lib.call("stg_peer_wait", self._flags.data_ptr() + k * 16 * 4, world, rank, it & 0xFFFF, WAIT_TIMEOUT_CYCLES,
                  self.status.data_ptr(), cur.cuda_stream)
        if ev:
            ev[2].record(cur)
        if self.meta_halo is not None:
            kernels.agg_packed_sum_rows(self.v_halo, self.meta_halo, plan.halo_out_rows, self.halo_rows(k), rs, out,
                                        accumulate=True)
        cur.wait_event(self._ev_sent)       # my sends have drained: x_own and the send buffer may be reused from here on
        kernels.launch_count += 3 if self.mode == "sm" else 2       # gather (or push + signal) + wait kernel of this call
        if ev:
            ev[3].record(cur)
            self.profile.append(ev)
        return out

    def check(self):
        """Raise if a wait kernel ever timed out (host sync; call outside timed regions)."""
        s = int(self.status.item())
        if s != 0:
            raise RuntimeError(f"rank {self.plan.rank}: halo flags of peer {s - 1} never arrived (peer lost or deadlock)")

    def profile_summary(self):
        """Mean device time (ms) of the segments of ``aggregate`` (set ``self.profile = []`` to collect)."""
        torch.cuda.synchronize()
        seg = {"own_pass": (0, 1), "wait_flags": (1, 2), "halo_pass": (2, 3), "total": (0, 3), "gather_or_push": (4, 5),
               "send": (5, 6), "flags": (6, 7), "exchange_total": (4, 7)}
        n = max(len(self.profile), 1)
        return {name: sum(e[a].elapsed_time(e[b]) for e in self.profile) / n for name, (a, b) in seg.items()}
