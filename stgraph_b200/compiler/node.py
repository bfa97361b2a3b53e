"""The symbolic centre vertex handed to a vertex program (``stgraph/compiler/node.py:8-26``).

``innbs`` / ``inedges`` (and the out- variants) are ONE-element lists: a list comprehension over
them runs once and Python's builtin ``sum`` turns into an aggregation.  A consequence the
reference ships with (trap T2): builtin ``max`` over such a list returns its only element.
"""
from .utils import EdgeDirection


class NbNode:
    def __init__(self, center, direction):
        self._central_node = center
        self._direction = direction


class NbEdge:
    def __init__(self, center, direction, nbnode):
        self._direction = direction
        if direction == EdgeDirection.IN:
            self.src, self.dst = nbnode, center
        else:
            self.src, self.dst = center, nbnode


class CentralNode:
    def __init__(self):
        self.innbs = [NbNode(self, EdgeDirection.IN)]
        self.outnbs = [NbNode(self, EdgeDirection.OUT)]
        self.inedges = [NbEdge(self, EdgeDirection.IN, nb) for nb in self.innbs]
        self.outedges = [NbEdge(self, EdgeDirection.OUT, nb) for nb in self.outnbs]
