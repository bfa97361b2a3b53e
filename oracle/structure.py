"""Structure oracle (numpy): CSR/CSC, degrees, snapshot diffs, PCSR/GPMA views.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every array here is a pure function of the snapshot's edge *set*, so plain
numpy ``lexsort`` / ``bincount`` / ``cumsum`` reproduce the reference's arrays
bit for bit.  Citations are relative to ``/root/reference``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class CsrArrays:
    """One direction of a graph in CSR form (all int32, like the reference)."""

    row_offset: np.ndarray      # [N+1]
    column_indices: np.ndarray  # [E]
    eids: np.ndarray            # [E]
    node_ids: np.ndarray        # [N] rows in non-increasing row-length order
    row_degrees: np.ndarray     # [N] length of each row  (CSR::out_degrees)
    col_degrees: np.ndarray     # [N] occurrences as a column (CSR::in_degrees)


def _as_edge_arrays(edges):
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    return e[:, 0].copy(), e[:, 1].copy()


def degree_sorted_node_ids(row_degrees: np.ndarray) -> np.ndarray:
    """Rows ordered by non-increasing length, ascending id inside a tie.

    The reference uses ``std::sort`` with ``lhs.first > rhs.first``
    (``stgraph/graph/static/csr.cu:143-154``): the tie order is unspecified,
    so only "non-increasing degree" is contractual; we pick the stable order.
    """
    n = row_degrees.shape[0]
    return np.lexsort((np.arange(n), -row_degrees.astype(np.int64))).astype(np.int32)


def forward_csr(src, dst, num_nodes: int) -> CsrArrays:
    """In-edge CSR: rows = destination, columns = source, eid = rank in (dst,src).

    ``stgraph/graph/static/static_graph.py:65-72`` sorts by ``(x[1], x[0])`` and
    numbers the edges by position; ``csr.cu:96-140`` (with
    ``is_edge_reverse=True``) then makes one row per destination.  Rows with no
    edge inherit the previous offset (``csr.cu:131-140``), i.e. a plain
    exclusive scan of the row lengths.
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    order = np.lexsort((src, dst))          # primary dst, secondary src
    s, d = src[order], dst[order]
    in_deg = np.bincount(d, minlength=num_nodes).astype(np.int32)
    out_deg = np.bincount(s, minlength=num_nodes).astype(np.int32)
    row_offset = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(in_deg, out=row_offset[1:])
    return CsrArrays(
        row_offset=row_offset,
        column_indices=s.astype(np.int32),
        eids=np.arange(s.shape[0], dtype=np.int32),
        node_ids=degree_sorted_node_ids(in_deg),
        row_degrees=in_deg,
        col_degrees=out_deg,
    )


def backward_csr(src, dst, num_nodes: int) -> CsrArrays:
    """Out-edge CSR carrying the *forward* eids.

    ``static_graph.py:75-78`` sorts the ``(src, dst, eid)`` triples
    lexicographically; ``csr.cu`` (``is_edge_reverse=False``) makes one row per
    source with ``column = dst`` and ``eids`` = the forward edge id.
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    forder = np.lexsort((src, dst))
    s, d = src[forder], dst[forder]
    feid = np.arange(s.shape[0], dtype=np.int64)
    border = np.lexsort((feid, d, s))       # primary src, then dst, then eid
    out_deg = np.bincount(s, minlength=num_nodes).astype(np.int32)
    in_deg = np.bincount(d, minlength=num_nodes).astype(np.int32)
    row_offset = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(out_deg, out=row_offset[1:])
    return CsrArrays(
        row_offset=row_offset,
        column_indices=d[border].astype(np.int32),
        eids=feid[border].astype(np.int32),
        node_ids=degree_sorted_node_ids(out_deg),
        row_degrees=out_deg,
        col_degrees=in_deg,
    )


def weighted_in_degrees(src, dst, weights_by_eid, num_nodes: int) -> np.ndarray:
    """``StaticGraph.weighted_in_degrees`` (``static_graph.py:124-126``).

    ``csr.cu:126`` accumulates ``edge_weight[eid]`` in fp32, in eid order, per
    destination row; the Python side truncates to int32.  ``weights_by_eid`` is
    indexed by the post-sort eid (trap T8).
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    order = np.lexsort((src, dst))
    d = dst[order]
    w = np.asarray(weights_by_eid, dtype=np.float32)
    acc = np.zeros(num_nodes, dtype=np.float32)
    # sequential fp32 accumulation in eid order, exactly like the host loop
    for i in range(d.shape[0]):
        acc[d[i]] = np.float32(acc[d[i]] + w[i])
    return acc.astype(np.int32)


# --------------------------------------------------------------------------
# dynamic graphs
# --------------------------------------------------------------------------
def snapshot_edge_sets(edge_lists) -> list[np.ndarray]:
    """Per-timestamp de-duplicated edge sets as sorted int64 keys ``(dst<<32)|src``.

    ``stgraph/graph/dynamic/dynamic_graph.py:58-63`` collapses duplicates with a
    Python ``set``.
    """
    out = []
    for edges in edge_lists:
        s, d = _as_edge_arrays(edges)
        out.append(np.unique((d << 32) | s))
    return out


def snapshot_updates(edge_lists):
    """``graph_updates[t] = {"add": S_t - S_{t-1}, "delete": S_{t-1} - S_t}``.

    Each list is sorted by (dst, src) (``dynamic_graph.py:66-79``); returned as
    ``(src, dst)`` int32 array pairs.
    """
    keys = snapshot_edge_sets(edge_lists)
    ups = []
    prev = np.zeros(0, dtype=np.int64)
    for k in keys:
        add = np.setdiff1d(k, prev, assume_unique=True)
        dele = np.setdiff1d(prev, k, assume_unique=True)
        ups.append({"add": _split_keys(add), "delete": _split_keys(dele)})
        prev = k
    return ups


def _split_keys(keys: np.ndarray):
    return (keys & 0xFFFFFFFF).astype(np.int32), (keys >> 32).astype(np.int32)


def labelled_forward_view(keys: np.ndarray, num_nodes: int, descending_rows: bool = False):
    """Compacted forward view of a PCSR/GPMA snapshot.

    ``keys``: sorted unique ``(dst<<32)|src``.  Labels are 1-based ranks among
    the live keys in key order (``gpma.cu:1121-1146``, ``pcsr.cu:748-760``);
    the kernels use ``label - 1`` as the edge id (``tpl_fa_gpma.jinja:32-34``,
    ``tpl_fa_pcsr.jinja:32-34``).  PCSR emits each row back to front
    (``pcsr.cu:842-855``) -> ``descending_rows=True``.
    """
    src, dst = _split_keys(keys)
    labels = np.arange(1, keys.shape[0] + 1, dtype=np.int32)
    in_deg = np.bincount(dst, minlength=num_nodes).astype(np.int32)
    out_deg = np.bincount(src, minlength=num_nodes).astype(np.int32)
    row_offset = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(in_deg, out=row_offset[1:])
    col, lab = src.copy(), labels.copy()
    if descending_rows:
        for r in range(num_nodes):
            b, e = row_offset[r], row_offset[r + 1]
            col[b:e] = col[b:e][::-1]
            lab[b:e] = lab[b:e][::-1]
    return CsrArrays(row_offset, col, lab, degree_sorted_node_ids(in_deg), in_deg, out_deg)


def labelled_backward_view(keys: np.ndarray, num_nodes: int, descending_rows: bool = False):
    """Dense transpose of the snapshot carrying the forward labels.

    ``gpma.cu:1165-1231`` (intra-row order nondeterministic in the reference:
    compare rows as sorted (col, label) pairs) and ``pcsr.cu:794-809``.
    """
    src, dst = _split_keys(keys)
    labels = np.arange(1, keys.shape[0] + 1, dtype=np.int64)
    order = np.lexsort((dst, src))
    out_deg = np.bincount(src, minlength=num_nodes).astype(np.int32)
    in_deg = np.bincount(dst, minlength=num_nodes).astype(np.int32)
    row_offset = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(out_deg, out=row_offset[1:])
    col = dst[order].astype(np.int32)
    lab = labels[order].astype(np.int32)
    if descending_rows:
        for r in range(num_nodes):
            b, e = row_offset[r], row_offset[r + 1]
            col[b:e] = col[b:e][::-1]
            lab[b:e] = lab[b:e][::-1]
    return CsrArrays(row_offset, col, lab, degree_sorted_node_ids(out_deg), out_deg, in_deg)


def rows_as_sorted_pairs(row_offset, cols, eids):
    """Canonical form for order-insensitive row comparison: per row, sorted (col,eid)."""
    row_offset = np.asarray(row_offset)
    cols = np.asarray(cols).astype(np.int64)
    eids = np.asarray(eids).astype(np.int64)
    n = row_offset.shape[0] - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(row_offset))
    order = np.lexsort((eids, cols, rows))
    return rows[order], cols[order], eids[order]
