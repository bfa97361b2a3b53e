"""CPU oracle for the STGraph vertex-centric aggregation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as something the CUDA path falls back to.

What it restates (citations are relative to ``/root/reference``):

* graph structure   -> ``oracle/structure.py``
    ``stgraph/graph/static/static_graph.py:65-78`` (edge ordering, eids),
    ``stgraph/graph/static/csr.cu:68-157`` (row offsets, degrees, node_ids),
    ``stgraph/graph/dynamic/dynamic_graph.py:56-79`` (snapshot diffs),
    ``stgraph/graph/dynamic/gpma/gpma.cu:1121-1231`` (labels, backward CSR),
    ``stgraph/graph/dynamic/pcsr/pcsr.cu:748-876`` (labels, descending rows).
* aggregation math  -> ``oracle/aggregate.py``
    ``stgraph/nn/pytorch/static/gcn_conv.py:162-182`` (GCN vertex programs),
    ``stgraph/nn/pytorch/static/gat_conv.py:48-56`` (GAT as traced, trap T2),
    ``stgraph/compiler/registry.py:195-406`` (per-op forward + gradient rules),
    ``stgraph/compiler/code_gen/templates/fa/tpl_fa_csr_unsorted.jinja:1-57``.
* the reference kernels themselves, emitted by the reference's own code
  generator and executed on the CPU through a SIMT-emulation shim ->
  ``oracle/ref_emulate.py`` + ``oracle/build_ref.py`` (outputs in
  ``oracle/_ref/``, fixtures in ``tests/golden/``).

Parity pinning status: the reference ships no numeric tests for this path
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference
itself run in the build container*: its CSR builder (``csr.cu`` compiled from
where it lies), its host-side PCSR (``pcsr.cu`` compiled from where it lies; fixtures
``tests/golden/ref_pcsr.npz``) and its generated CUDA kernels executed by the emulation
shim.  Still unpinned: GPMA (``gpma.cu`` needs CUDA dynamic parallelism v1, which does not
exist for sm_100 -- SURVEY.md trap T4); its view contract is the same labelled view as PCSR's
with ascending rows.
The committed fixtures in ``tests/golden/`` were produced by
``oracle/make_golden.py``; see DESIGN.md "Oracle".
"""
