#!/usr/bin/env python
"""Benchmark of the hot path: fused GCN neighbour aggregation (forward + backward) on the
ogbn-products-shaped synthetic graph of BASELINE.json (config 5: 2,449,029 V / 61,859,140 E, F=100).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference ...                     # CPU reference (torch index_add), rank 0 only

One step = one forward aggregation over the in-edge CSR + one backward aggregation over the
out-edge CSR ("primary measurement = the F=100 aggregation fwd+bwd").
``value`` = algorithmic bytes of the whole job / max-over-ranks device time (GB/s); ``e2e`` = the same
through the C-ABI host-buffer entry point with H2D/D2H inside the timed region; ``roofline`` = the
forward kernel against the measured HBM peak; ``cpu_baseline`` = the oracle port on the host cores.
At N>1 the graph is row-partitioned (edge-balanced) and feature rows are exchanged over NCCL each
aggregation (strong scaling: total work fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "gcn_fused_aggregation_fwd_bwd_algorithmic_hbm_gbs"
UNIT = "GB/s"
FEAT = 100


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("STG_BENCH_SCALE", "1.0")),
                    help="shrink the graph (debug only; the judged number is scale=1)")
    ap.add_argument("--locality", type=float, default=0.9)
    ap.add_argument("--window", type=int, default=8192)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / e2e / locality-free legs")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get("agg_rows_kernel_fwd_dram_bytes_per_launch")
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(args, dev):
    from stgraph_b200.utils import synthetic

    d = synthetic.products_shaped(seed=0, device=dev, scale=args.scale, locality=args.locality, window=args.window)
    return d


def cpu_reference_sample(src, dst, n, feat, budget_edges=4_000_000, reps=3, threads=None):
    """torch-CPU index_add of the same vertex program on the first rows holding ~budget_edges edges."""
    import numpy as np

    from oracle import aggregate as A
    from oracle import structure as S

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    src, dst = src.cpu().numpy(), dst.cpu().numpy()
    e_full = src.shape[0]
    indeg = np.bincount(dst, minlength=n)
    ro = np.concatenate([[0], np.cumsum(indeg)])
    rows = int(np.searchsorted(ro, min(budget_edges, e_full), side="left"))
    rows = max(1, min(rows, n))
    keep = dst < rows
    f = S.forward_csr(src[keep], dst[keep], n)
    e_s = int(keep.sum())
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, feat, generator=g)
    norm = torch.rand(n, generator=g) + 0.5
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        A.scaled_sum(f.row_offset, f.column_indices, f.eids, x, norm, None, norm, dtype=torch.float32)
        times.append(time.perf_counter() - t0)
    return min(times), e_s, e_full, rows, threads


def run_reference(args):
    """--impl reference: the reference's GPU-only path restated as torch-CPU index_add, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from stgraph_b200.utils import synthetic

    d = synthetic.products_shaped(seed=0, device="cpu", scale=min(args.scale, 0.25), locality=args.locality,
                                  window=args.window)
    n_full, e_full = int(2449029 * args.scale), int(61859140 * args.scale) // 2 * 2
    n = d["num_nodes"]
    b_full = 2 * synthetic.gcn_algorithmic_bytes(n_full, e_full, FEAT)
    ts = []
    e_s = rows = threads = None
    for i in range(args.warmup + args.steps):
        t, e_s, e_samp_full, rows, threads = cpu_reference_sample(d["src"], d["dst"], n, FEAT, budget_edges=3_000_000,
                                                                  reps=1)
        if i >= args.warmup:
            ts.append(t)
    t_step = sum(ts) / len(ts)
    # one sample = forward aggregation of e_s edges; a full step is fwd+bwd over e_full edges
    value = b_full * (e_s / (2.0 * e_full)) / t_step / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config5: GCN aggregation fwd+bwd, ogbn-products-shaped synthetic graph",
                       "num_nodes": n_full, "num_edges": e_full, "feat": FEAT, "locality": args.locality,
                       "window": args.window, "scale": args.scale,
                       "sample": {"edges": e_s, "rows": rows, "graph_nodes": n,
                                  "note": "each step = forward aggregation of a bounded row sample; value scaled by edge share"}},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"forward aggregation of the first {rows} destination rows ({e_s} edges, F={FEAT}) "
                                       f"of a {n}-node graph from the same generator; throughput scaled by edge share"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from stgraph_b200 import kernels
    from stgraph_b200.dist import PartitionedGraph
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    d = make_workload(args, dev)
    if world > 1:   # every rank must hold the same graph: rank 0's
        dist.broadcast(d["src"], src=0)
        dist.broadcast(d["dst"], src=0)
    n, e = d["num_nodes"], int(d["src"].shape[0])
    graph = StaticGraph(torch.stack([d["src"], d["dst"]], 1), None, n)
    norm = graph.degree_norm().reshape(-1).contiguous()
    gen = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, FEAT, device=dev, generator=gen)
    gout = torch.randn(n, FEAT, device=dev, generator=gen)
    b_alg_one = synthetic.gcn_algorithmic_bytes(n, e, FEAT)
    b_alg_step = 2 * b_alg_one

    halo_info = None
    parity = None
    dist_mode = os.environ.get("STG_DIST_MODE", "push")
    if world > 1:
        pg = PartitionedGraph(graph, rank, world)
        # single-GPU result of this rank's rows, to check the partitioned path after the first step
        f_lo_, f_hi_ = pg.local_rows("fwd")
        ref_rows = kernels.agg_scaled_sum(graph.fwd_view(), x, norm, None, norm)[f_lo_:f_hi_].clone()
        mag_rows = kernels.agg_scaled_sum(graph.fwd_view(), x.abs(), norm, None, norm)[f_lo_:f_hi_].clone()
        peer_ok = False
        if dist_mode == "pull":
            try:
                from stgraph_b200.dist import PeerBlocks

                # feature and gradient rows are owned by the forward (destination) partition
                px = PeerBlocks(pg.fwd_bounds, FEAT, rank, world, dev)
                pgout = PeerBlocks(pg.fwd_bounds, FEAT, rank, world, dev)
                peer_ok = True
            except Exception as ex:            # symmetric memory unavailable: fall back to the NCCL halo exchange
                if rank == 0:
                    print(f"[bench] peer-memory path unavailable ({ex!r}); using NCCL halo all-to-all", file=sys.stderr)
        if dist_mode == "push":
            try:
                hf, hb = pg.halo_plans()
                hf.setup_push(FEAT)
                hb.setup_push(FEAT)
                peer_ok = True
            except Exception as ex:
                peer_ok = False
                if rank == 0:
                    print(f"[bench] peer-memory push path unavailable ({ex!r}); using NCCL halo all-to-all", file=sys.stderr)
        if peer_ok and dist_mode == "push":
            own_lo, own_hi = hf.own_lo, hf.own_hi
            x_own = x[own_lo:own_hi].clone()
            g_own = gout[own_lo:own_hi].clone()
            ns_own = norm[own_lo:own_hi].contiguous()
            nsh_f, nsh_b = norm[hf.halo_ids].contiguous(), norm[hb.halo_ids].contiguous()   # norm is replicated (10 MB)
            rs_f = norm[hf.row_lo:hf.row_hi].contiguous()
            rs_b = norm[hb.row_lo:hb.row_hi].contiguous()
            out_f = torch.empty(hf.n_rows, FEAT, device=dev)
            out_b = torch.empty(hb.n_rows, FEAT, device=dev)
            push_blocks = int(os.environ.get("STG_PUSH_BLOCKS", "32"))
            push_flow = os.environ.get("STG_PUSH_FLOW", "serial")
            halo_info = {"mode": "halo rows pushed over NVLink by our kernel (posted stores into symmetric memory) during the "
                                 "own-source pass; " + ("halo-source pass += over the rows with remote neighbours"
                                                        if push_flow == "serial" else
                                                        "halo pass concurrent with the own-source pass (vector red.add)")
                                 + "; no NCCL on the data path", "flow": push_flow,
                         "halo_rows_fwd": hf.n_halo, "halo_rows_bwd": hb.n_halo, "own_rows": hf.n_own,
                         "full_allgather_rows": n - hf.n_own, "halo_edges_fwd": int(hf.halo_cols.shape[0]),
                         "own_edges_fwd": int(hf.own_cols.shape[0]), "push_blocks": push_blocks}
            del x, gout

            def step(ev=None):
                if ev:
                    ev[0].record()
                hf.aggregate_push(kernels, x_own, ns_own, nsh_f, rs_f, out_f, push_blocks=push_blocks, flow=push_flow)
                if ev:
                    ev[1].record()
                hb.aggregate_push(kernels, g_own, ns_own, nsh_b, rs_b, out_b, push_blocks=push_blocks, flow=push_flow)
        elif peer_ok and dist_mode == "pull":
            hf, hb = pg.halo_plans()
            own_lo, own_hi = hf.own_lo, hf.own_hi
            px.own.copy_(x[own_lo:own_hi])
            pgout.own.copy_(gout[own_lo:own_hi])
            halo_f = hf.new_halo_buffer(FEAT, x)
            halo_b = hb.new_halo_buffer(FEAT, gout)
            ns_own = norm[own_lo:own_hi].contiguous()
            nsh_f, nsh_b = norm[hf.halo_ids].contiguous(), norm[hb.halo_ids].contiguous()   # norm is replicated (10 MB)
            rs_f = norm[hf.row_lo:hf.row_hi].contiguous()
            rs_b = norm[hb.row_lo:hb.row_hi].contiguous()
            out_f = torch.empty(hf.n_rows, FEAT, device=dev)
            out_b = torch.empty(hb.n_rows, FEAT, device=dev)
            hf.split_views()
            hb.split_views()
            pull_blocks = int(os.environ.get("STG_PULL_BLOCKS", "64"))
            halo_info = {"mode": "halo rows pulled over NVLink by our kernel (symmetric memory), overlapped with the "
                                 "own-source pass; no NCCL on the data path",
                         "halo_rows_fwd": hf.n_halo, "halo_rows_bwd": hb.n_halo, "own_rows": hf.n_own,
                         "full_allgather_rows": n - hf.n_own, "halo_edges_fwd": int(hf.halo_cols.shape[0]),
                         "own_edges_fwd": int(hf.own_cols.shape[0]), "pull_blocks": pull_blocks}
            del x, gout

            def step(ev=None):
                if ev:
                    ev[0].record()
                hf.aggregate_pull(kernels, px, halo_f, ns_own, nsh_f, rs_f, out_f, pull_blocks=pull_blocks)
                if ev:
                    ev[1].record()
                hb.aggregate_pull(kernels, pgout, halo_b, ns_own, nsh_b, rs_b, out_b, pull_blocks=pull_blocks)
        else:
            dist_mode = "halo"
            hf, hb = pg.halo_plans()                      # halo-only exchange plans (dist/halo.py)
            own_lo, own_hi = hf.own_lo, hf.own_hi
            x_own = x[own_lo:own_hi].clone()
            g_own = gout[own_lo:own_hi].clone()
            halo_f = hf.new_halo_buffer(FEAT, x)
            halo_b = hb.new_halo_buffer(FEAT, gout)
            ns_own = norm[own_lo:own_hi].contiguous()
            nsh_f, nsh_b = hf.halo_vector(ns_own), hb.halo_vector(ns_own)
            rs_f = norm[hf.row_lo:hf.row_hi].contiguous()
            rs_b = norm[hb.row_lo:hb.row_hi].contiguous()
            out_f = torch.empty(hf.n_rows, FEAT, device=dev)
            out_b = torch.empty(hb.n_rows, FEAT, device=dev)
            hf.split_views()
            hb.split_views()
            halo_info = {"mode": "NCCL halo all-to-all overlapped with the own-source pass",
                         "halo_rows_fwd": hf.n_halo, "halo_rows_bwd": hb.n_halo, "own_rows": hf.n_own,
                         "full_allgather_rows": n - hf.n_own,
                         "halo_edges_fwd": int(hf.halo_cols.shape[0]), "own_edges_fwd": int(hf.own_cols.shape[0])}
            del x, gout                                   # only the owned blocks + halos stay resident

            def step(ev=None):
                if ev:
                    ev[0].record()
                hf.aggregate(kernels, x_own, halo_f, ns_own, nsh_f, rs_f, out_f)
                if ev:
                    ev[1].record()
                hb.aggregate(kernels, g_own, halo_b, ns_own, nsh_b, rs_b, out_b)
    else:
        out_f = torch.empty_like(x)
        out_b = torch.empty_like(x)
        vf, vb = graph.fwd_view(), graph.bwd_view()
        # the public path of a static graph (GCNConv -> executor -> kernels.agg_scaled_sum_graph): the {col, scale}
        # array of this norm is packed on the first (warm-up) call and reused, like the CSR arrays themselves
        fcsr, bcsr = graph._forward_graph, graph._backward_graph

        def step(ev=None):
            if ev:
                ev[0].record()
            kernels.agg_scaled_sum_graph(fcsr, x, norm, None, norm, out=out_f)
            if ev:
                ev[1].record()
            kernels.agg_scaled_sum_graph(bcsr, gout, norm, None, norm, out=out_b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        torch.cuda.synchronize()
        ok = bool(((out_f - ref_rows).abs() <= 1e-5 * mag_rows + 1e-30).all())
        t_ok = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        parity = bool(int(t_ok.item()))
        assert parity, "partitioned aggregation disagrees with the single-GPU result"
        del ref_rows, mag_rows
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = kernels.launch_count
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_beg.record()
    for i in range(args.steps):
        step(kev[i])
    t_end.record()
    barrier()
    launches = kernels.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1 and os.environ.get("STG_DIST_PROFILE") and dist_mode == "pull":
        hf.profile = []
        for _ in range(5):
            hf.aggregate_pull(kernels, px, halo_f, ns_own, nsh_f, rs_f, out_f, pull_blocks=pull_blocks)
        summ = hf.profile_summary()
        summ.update({"rows": hf.n_rows, "own_edges": int(hf.own_cols.shape[0]), "halo_edges": int(hf.halo_cols.shape[0]),
                     "halo_rows": hf.n_halo})
        hf.profile = None
        allsum = [None] * world
        dist.all_gather_object(allsum, summ)
        if rank == 0:
            for r_, s_ in enumerate(allsum):
                print(f"[bench] rank {r_} aggregate_pull segments (ms): " + json.dumps({k: round(v, 3) for k, v in s_.items()}),
                      file=sys.stderr)
    if world > 1 and os.environ.get("STG_DIST_PROFILE") and dist_mode == "push":
        hf.push_profile = []
        for _ in range(6):
            hf.aggregate_push(kernels, x_own, ns_own, nsh_f, rs_f, out_f, push_blocks=push_blocks, flow=push_flow)
        summ = hf.push_profile_summary()
        hf.push_profile = None
        allsum = [None] * world
        dist.all_gather_object(allsum, summ)
        if rank == 0:
            for r_, s_ in enumerate(allsum):
                print(f"[bench] rank {r_} aggregate_push segments (ms): " + json.dumps({k: round(v, 3) for k, v in s_.items()}),
                      file=sys.stderr)
    ms_total = t_beg.elapsed_time(t_end)
    ms_fwd_kernel = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    if world > 1:
        t = torch.tensor([ms_total, ms_fwd_kernel], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_fwd_kernel = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = b_alg_step / (ms_step * 1e-3) / 1e9

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        # ---- end to end through the C-ABI host-buffer entry point (pinned host memory) ----
        xh = x.cpu().pin_memory()
        gh = gout.cpu().pin_memory()
        nh = norm.cpu().pin_memory()
        oh = torch.empty(n, FEAT).pin_memory()
        oh2 = torch.empty(n, FEAT).pin_memory()
        scratch = torch.empty(kernels.host_scratch_bytes(n, e, FEAT), dtype=torch.uint8, device=dev)
        scratch2 = torch.empty_like(scratch)
        for _ in range(2):
            kernels.agg_scaled_sum_host(vf, xh, oh, scratch, nh, None, nh)
        torch.cuda.synchronize()
        k = max(3, min(args.steps, 8))
        # (a) one blocking call after the other: H2D, kernel, D2H strictly in sequence
        t0 = time.perf_counter()
        for _ in range(k):
            kernels.agg_scaled_sum_host(vf, xh, oh, scratch, nh, None, nh)
            kernels.agg_scaled_sum_host(vb, gh, oh2, scratch, nh, None, nh)
        torch.cuda.synchronize()
        e2e_serial_ms = (time.perf_counter() - t0) / k * 1e3
        # (b) the forward and the backward call enqueued on two streams (two scratch buffers): one call's H2D
        # overlaps the other's D2H; every step still copies all of its inputs in and all of its results out
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        for _ in range(2):
            kernels.agg_scaled_sum_host(vf, xh, oh, scratch, nh, None, nh, stream=sa)
            kernels.agg_scaled_sum_host(vb, gh, oh2, scratch2, nh, None, nh, stream=sb)
        torch.cuda.synchronize()
        oh.zero_()
        t0 = time.perf_counter()
        for _ in range(k):
            kernels.agg_scaled_sum_host(vf, xh, oh, scratch, nh, None, nh, stream=sa)
            kernels.agg_scaled_sum_host(vb, gh, oh2, scratch2, nh, None, nh, stream=sb)
        sa.synchronize()
        sb.synchronize()
        e2e_ms = (time.perf_counter() - t0) / k * 1e3
        h2d = 2 * (n * FEAT * 4 + 2 * n * 4)
        d2h = 2 * n * FEAT * 4
        extras["e2e"] = {"value": b_alg_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
                         "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "api": "stg_agg_scaled_sum_f32_host_async: forward and backward calls on two streams (pinned host "
                                "buffers; H2D + kernels + D2H per call, one call's H2D overlapping the other's D2H)",
                         "blocking_calls_ms_per_step": e2e_serial_ms,
                         "blocking_calls_value": b_alg_step / (e2e_serial_ms * 1e-3) / 1e9}
        assert torch.equal(oh, out_f.cpu()), "host-buffer path and device path disagree"
        del scratch, scratch2
        # ---- the plain kernel (column load + dependent norm gather per edge) on the same inputs, for the record ----
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            kernels.agg_scaled_sum(vf, x, norm, None, norm, out=out_f)
        a.record()
        for _ in range(5):
            kernels.agg_scaled_sum(vf, x, norm, None, norm, out=out_f)
            kernels.agg_scaled_sum(vb, gout, norm, None, norm, out=out_b)
        b.record()
        torch.cuda.synchronize()
        ms_plain = a.elapsed_time(b) / 5
        extras["plain_kernel"] = {"ms_per_step": ms_plain, "value": b_alg_step / (ms_plain * 1e-3) / 1e9, "unit": UNIT,
                                  "note": "stg_agg_scaled_sum_f32 (no packed edge metadata)"}
        # ---- locality-free variant of the same shape (secondary figure) ----
        try:
            d0 = synthetic.products_shaped(seed=0, device=dev, scale=args.scale, locality=0.0)
            g0 = StaticGraph(torch.stack([d0["src"], d0["dst"]], 1), None, n)
            nm0 = g0.degree_norm().reshape(-1).contiguous()
            v0f, v0b = g0.fwd_view(), g0.bwd_view()
            for _ in range(3):
                kernels.agg_scaled_sum(v0f, x, nm0, None, nm0, out=out_f)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                kernels.agg_scaled_sum(v0f, x, nm0, None, nm0, out=out_f)
                kernels.agg_scaled_sum(v0b, gout, nm0, None, nm0, out=out_b)
            b.record()
            torch.cuda.synchronize()
            ms0 = a.elapsed_time(b) / 5
            extras["locality_free"] = {"ms_per_step": ms0, "value": b_alg_step / (ms0 * 1e-3) / 1e9, "unit": UNIT}
            del g0, d0
        except Exception as ex:  # secondary figure only
            extras["locality_free"] = {"error": str(ex)[:200]}
        # ---- CPU baseline: the oracle port on the host cores, bounded sample ----
        t_cpu, e_s, e_f, rows, threads = cpu_reference_sample(d["src"], d["dst"], n, FEAT)
        cpu_val = b_alg_one * (e_s / e_f) / t_cpu / 1e9
        extras["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                                  "sample": f"torch-CPU index_add (fp32) forward aggregation of the first {rows} "
                                            f"destination rows = {e_s} of {e_f} edges, F={FEAT}, best of 3; "
                                            f"throughput scaled by edge share", "seconds": t_cpu}

    if rank == 0 and world == 1 and not args.no_extras and os.environ.get("STG_BENCH_EPOCHS", "1") != "0":
        # ---- the other half of BASELINE.json's metric: GCN / TGCN epoch ms on configs 1 and 2 (scripts/bench_configs.py:
        # the reference's training loops on the synthetic Cora- and WikiMaths-shaped inputs); secondary figures
        # (a separate process with a time limit: nothing that happens there can cost the main measurement)
        try:
            import tempfile

            with tempfile.TemporaryDirectory() as tmp:
                path = os.path.join(tmp, "configs.json")
                env = dict(os.environ, STG_CONFIGS_OUT=path, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank)))
                subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bench_configs.py"), "1", "2"], env=env, cwd=ROOT,
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=240, check=True)
                res = json.load(open(path))
            c2 = res.get("config2_tgcn_wikimaths_epoch_ms", {})
            extras["epoch_ms"] = {
                "config1_gcn_cora_2layer": res.get("config1_gcn_cora_epoch_ms"),
                "config2_tgcn_wikimaths_723_steps": {"drop_in_layers": c2.get("dropin"), "fused_cell": c2.get("fused"),
                                                     "fused_cell_one_cuda_graph": c2.get("fused_cudagraph")},
                "note": "fwd + bwd + Adam per epoch, synthetic data of the reference datasets' shapes (BASELINE.json configs 1-2)"}
            for k_, v_ in res.items():
                if k_.endswith("_error"):
                    extras["epoch_ms"][k_] = str(v_)[:200]
        except Exception as ex:      # secondary figures only
            extras["epoch_ms"] = {"error": repr(ex)[:300]}

    if rank == 0:
        peak, peak_src = peaks()
        packed = world == 1 and bool(graph._forward_graph._meta_cache)
        achieved = b_alg_one / (ms_fwd_kernel * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config5: GCN aggregation fwd+bwd, ogbn-products-shaped synthetic graph",
                       "num_nodes": n, "num_edges": e, "feat": FEAT, "locality": args.locality, "window": args.window,
                       "l2_policy": "inputs (980 MB features + 500 MB structure) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"edge-balanced row partition x{world}, {dist_mode}" if world > 1 else "single GPU",
                       "halo": halo_info, "partitioned_result_matches_single_gpu": parity,
                       "edge_metadata": ("{col, norm[col]} packed per CSR slot once per graph (8 B/edge, read coalesced); "
                                         "algorithmic bytes still count 4 B/edge") if packed else "column_indices + norm gather",
                       "scale": args.scale},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": UNIT, "frac": achieved / peak,
                         "traffic": ncu_traffic(),
                         "kernel": ("agg_rows_pipe_kernel<4,32,1,4,8,kPacked> + agg_hub_kernel<4,32,1,kPacked>" if packed else
                                    "agg_rows_pipe_kernel<4,32,1,4,8,kPlain> + agg_hub_kernel<4,32,1,kPlain>")
                                   + " overlapped (forward, in-edge CSR)",
                         "kernel_ms": ms_fwd_kernel, "algorithmic_bytes": b_alg_one, "peak_source": peak_src,
                         "gather_model_gbs": 4.0 * (e * FEAT + n * FEAT + e) / (ms_fwd_kernel * 1e-3) / 1e9},
            "gpu_launches": launches, "clocks": clocks,
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
