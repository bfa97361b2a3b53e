"""Live pinning of the structure oracle against the reference's OWN code, on random inputs.

``oracle/build_ref.py`` compiles the reference's ``csr.cu`` and ``pcsr.cu`` from where they lie into
``oracle/_ref/*.so`` (host code; no GPU needed to run it).  Where those libraries exist (the build container, and
the GPU box, to which ``oracle/_ref`` travels) these tests drive them on seeded random graphs / streams and demand
bit-identical arrays from ``oracle/structure.py`` -- the committed fixtures in ``tests/golden`` are two such runs.
"""
import os

import numpy as np
import pytest

from oracle import ref_emulate as RE
from oracle import structure as S

HAVE_CSR = os.path.exists(os.path.join(RE.REF_DIR, "ref_csr.so"))
HAVE_PCSR = os.path.exists(os.path.join(RE.REF_DIR, "ref_pcsr.so"))


def _graph(n, e, seed):
    rng = np.random.default_rng(seed)
    key = rng.choice(n * n, size=min(e, n * n), replace=False)
    return (key // n).astype(np.int32), (key % n).astype(np.int32)


@pytest.mark.skipif(not HAVE_CSR, reason="oracle/_ref/ref_csr.so not built (needs /root/reference)")
@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (2, 1, 1), (7, 20, 2), (50, 400, 3), (300, 5000, 4), (1000, 3000, 5)])
def test_static_structure_oracle_equals_reference_csr_builder(n, e, seed):
    src, dst = _graph(n, e, seed)
    if e == 0:
        src, dst = np.zeros(0, np.int32), np.zeros(0, np.int32)
    w = np.random.default_rng(seed + 100).uniform(0.1, 1.0, size=src.shape[0]).astype(np.float32)
    fwd, bwd = RE.reference_static_graph(src, dst, w, n)
    f, b = S.forward_csr(src, dst, n), S.backward_csr(src, dst, n)
    for name in ("row_offset", "column_indices", "eids"):
        np.testing.assert_array_equal(getattr(f, name), getattr(fwd, name), err_msg=f"fwd {name}")
        np.testing.assert_array_equal(getattr(b, name), getattr(bwd, name), err_msg=f"bwd {name}")
    np.testing.assert_array_equal(f.row_degrees, fwd.out_degrees)         # CSR::out_degrees = row lengths
    np.testing.assert_array_equal(f.col_degrees, fwd.in_degrees)
    np.testing.assert_array_equal(b.row_degrees, bwd.out_degrees)
    for ids, deg in ((fwd.node_ids, f.row_degrees), (bwd.node_ids, b.row_degrees)):
        assert sorted(ids.tolist()) == list(range(n)) and np.all(np.diff(deg[ids]) <= 0)
    np.testing.assert_array_equal(S.weighted_in_degrees(src, dst, w, n), fwd.weighted_out_degrees.astype(np.int32))


def _stream(n, t_count, base, churn, seed):
    rng = np.random.default_rng(seed)
    cur = set()
    while len(cur) < base:
        a, b = rng.integers(0, n, 2)
        if a != b:
            cur.add((int(a), int(b)))
    snaps = []
    for _ in range(t_count):
        snaps.append(sorted(cur))
        curl = sorted(cur)
        for i in rng.choice(len(curl), size=min(churn, len(curl)), replace=False):
            cur.discard(curl[i])
        while len(cur) < base:
            a, b = rng.integers(0, n, 2)
            if a != b:
                cur.add((int(a), int(b)))
    return snaps


@pytest.mark.skipif(not HAVE_PCSR, reason="oracle/_ref/ref_pcsr.so not built (needs /root/reference)")
@pytest.mark.parametrize("n,T,base,churn,seed", [(5, 4, 6, 3, 0), (16, 6, 60, 30, 1), (40, 5, 200, 200, 2), (120, 8, 900, 150, 3)])
def test_pcsr_oracle_equals_reference_pcsr_on_random_streams(n, T, base, churn, seed):
    """Forward roll, then backward roll, like PCSRGraph (pcsr_graph.py:45-166); `churn == base` replaces every edge."""
    snaps = _stream(n, T, base, churn, seed)
    keys = S.snapshot_edge_sets(snaps)
    ups = S.snapshot_updates(snaps)
    ref = RE.RefPcsr(n, len({e for s in snaps for e in s}))
    try:
        def same(t, reverse):
            exp = (S.labelled_backward_view if reverse else S.labelled_forward_view)(keys[t], n, descending_rows=True)
            ro, col, eid, nid, ind, outd = ref.build(reverse)
            np.testing.assert_array_equal(ro.astype(np.int32), exp.row_offset, err_msg=f"t={t} rev={reverse} row_offset")
            np.testing.assert_array_equal(col.astype(np.int32), exp.column_indices, err_msg=f"t={t} rev={reverse} cols")
            np.testing.assert_array_equal(eid.astype(np.int32), exp.eids, err_msg=f"t={t} rev={reverse} labels")
            if not reverse:
                np.testing.assert_array_equal(outd.astype(np.int32), exp.row_degrees)
                np.testing.assert_array_equal(ind.astype(np.int32), exp.col_degrees)

        for t in range(T):
            ref.step(ups[t]["add"], ups[t]["delete"])
            same(t, False)
            same(t, True)
        for t in range(T - 1, 0, -1):
            ref.step(ups[t]["delete"], ups[t]["add"])
            same(t - 1, True)
            same(t - 1, False)
    finally:
        ref.close()
