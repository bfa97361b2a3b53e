"""GPMAGraph (mirror of ``stgraph/graph/dynamic/gpma/gpma_graph.py:28-152``).

Same constructor / ``graph_type() == "gpma"`` / roll-forward, cache and rewind behaviour.  The
reference's GPMA (``gpma.cu``) is a gapped packed-memory array updated level by level with
device-side launches (CDP1 -- it does not even compile for sm_100, SURVEY.md trap T4).  Here the
live snapshot is the gap-free sorted key array and one update is a merge-path insert/delete
(``csrc/snapshot.cu``); the compacted view -- rows sorted by source id, labels = 1 + rank among live
keys (``gpma.cu:1121-1146``), dense transposed backward CSR with the same labels
(``gpma.cu:1165-1231``, deterministic here), in/out degrees -- is what the contract pins.
Parity of this class is against the structure oracle (``oracle/structure.py``); the reference module
cannot be built for this GPU, so it is "parity unpinned" by reference outputs.
"""
from .dynamic_graph import KeyedDynamicGraph


class GPMAGraph(KeyedDynamicGraph):
    _descending_rows = False
    _label_base = 1

    def graph_type(self) -> str:
        return "gpma"
